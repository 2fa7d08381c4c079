"""CPU: the off-box reference replay (csharp/ReferenceReplay -> tests/tools/replay_reference_dump.py) checked end to end on
SYNTHETIC dumps written in the dumper's CSV layout -- so that once a machine with a .NET SDK has produced the real dump,
pinning the oracles to reference-executed vectors is one command.  If tests/golden/cartpole_reference.npz (the converted real
dump) is present, the oracle is held to it here; it is absent in this image (no dotnet), and the test says so."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import replay_reference_dump as RR  # noqa: E402


def test_cartpole_dump_layout_and_check(tmp_path):
    import make_golden as MG   # the independent pure-Python restatement of CartPoleEnv.cs:137-186 stands in for the C# run
    rng = np.random.default_rng(3)
    lines = ["x,x_dot,theta,theta_dot,action,sbd,nx,nx_dot,ntheta,ntheta_dot,reward,done,next_sbd"]
    for i in range(4000):
        s = rng.uniform([-2.6, -3, -0.25, -3.5], [2.6, 3, 0.25, 3.5]).astype(np.float32).astype(np.float64)
        if i % 3 == 1:
            s[0] = np.float32(rng.choice([-1.0, 1.0]) * 2.4 - 0.02 * s[1])
        a, sbd = int(rng.integers(0, 2)), int(rng.integers(-1, 3))
        ns, rew, done, nsbd = MG.cartpole_step(tuple(float(v) for v in s), a, sbd)
        lines.append(",".join([repr(float(v)) for v in s] + [str(a), str(sbd)] + [repr(float(v)) for v in ns] + [repr(float(rew)), str(int(done)), str(nsbd)]))
    path = tmp_path / "cartpole_reference.csv"
    path.write_text("\n".join(lines) + "\n")
    fx = RR.load_cartpole(str(path))
    out = RR.check_cartpole(fx)
    assert out["transitions"] == 4000 and out["oracle_f64_max_rel_err"] <= 1e-12 and out["engine_f32_max_rel_err"] <= 1e-5
    assert 0 < fx["done"].sum() < 4000
    fx["done"][7] ^= 1   # a reference that disagreed would be caught
    with pytest.raises(AssertionError):
        RR.check_cartpole(fx)


def test_lunar_dump_layout_and_replay(tmp_path):
    """An episode of the generic engine (PID policy, engine draws) written in the dumper's layout replays without divergence
    under the options it was made with and diverges at the first touch-down under another BeginContact semantic."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
    import world2d_lib as W
    from lunar_pid_study import pid
    seed, gid = 1000, 3
    wi, ti = W.ctor_draws(seed, gid)
    w = W.LunarWorld(wind_idx=wi, torque_idx=ti)
    rd, zd = W.reset_draws(seed, gid, 0), W.step_draws(seed, gid, 0)
    obs = w.reset(rd, zd)
    f = lambda v: repr(float(v))   # noqa: E731
    lines = ["# seed=%d policy=pid wind_idx=%d torque_idx=%d" % (seed, wi, ti), "kind,action,d0,d1,o0,o1,o2,o3,o4,o5,o6,o7,reward,done,extra...",
             ",".join(["reset", "0", f(zd[0]), f(zd[1])] + [f(v) for v in obs] + ["0", "0"] + [f(v) for v in rd])]
    total, touched = np.float32(0), False
    for k in range(3000):
        a = pid(obs)
        sd = W.step_draws(seed, gid, k + 1)
        obs, r, d = w.step(a, sd)
        touched = touched or obs[6] > 0 or obs[7] > 0
        total = np.float32(total + r)
        lines.append(",".join(["step", str(a), f(sd[0]), f(sd[1])] + [f(v) for v in obs] + [f(r), str(d)]))
        if d:
            break
    lines.append("# total_reward=%s" % f(total))
    w.close()
    assert touched
    path = tmp_path / "lunar_pid_seed1000.csv"
    path.write_text("\n".join(lines) + "\n")
    ep = RR.load_lunar(str(path))
    assert ep["wind_idx"] == wi and len(ep["action"]) == k + 1 and ep["total_reward"] == pytest.approx(float(total))
    first, steps, ret, worst = RR.replay_lunar(ep)
    assert first == -1 and steps == k + 1 and worst == 0.0 and ret == pytest.approx(float(total))
    first2, _, _, _ = RR.replay_lunar(ep, begin_contact_false=2)
    assert first2 > 0   # contacts that never resolve: the trajectories part at touch-down
    rep = RR.lunar_report(ep)
    assert len(rep) == 12 and sum(1 for _, fd, _, _, _ in rep if fd == -1) >= 1


def test_oracle_against_reference_executed_cartpole_vectors_if_present():
    path = os.path.join(ROOT, "tests", "golden", "cartpole_reference.npz")
    if not os.path.exists(path):
        pytest.skip("no reference-executed vectors in this image (no dotnet): run csharp/ReferenceReplay off-box, then "
                    "tests/tools/replay_reference_dump.py <dir> --write-golden")
    fx = dict(np.load(path))
    RR.check_cartpole(fx)
