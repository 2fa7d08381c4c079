"""GPU parity of LunarLander (discrete + continuous) against the CPU oracle.  Both sides run the same
float32 operation sequence, so every output -- observations, rewards, done flags, the full rigid-body
state, contact flags, warm-start impulses -- is compared BIT FOR BIT.  Parity with the reference
itself is unpinned (Aether.Physics2D and NumSharp are absent; see oracle/README.md)."""
import numpy as np
import pytest

import oracle_lib as O
import gymnet_b200 as G

pytestmark = pytest.mark.gpu


def pid(s):
    """The reference test's heuristic (tests/Gym.Tests/Envs/Aether/LunarLanderEnvironment.cs:102-150), discrete."""
    angle_targ = np.clip(s[:, 0] * 0.5 + s[:, 2] * 1.0, -0.4, 0.4)
    hover_targ = 0.55 * np.abs(s[:, 0])
    angle_todo = (angle_targ - s[:, 4]) * 0.5 - s[:, 5] * 1.0
    hover_todo = (hover_targ - s[:, 1]) * 0.5 - s[:, 3] * 0.5
    legs = (s[:, 6] > 0) | (s[:, 7] > 0)
    angle_todo = np.where(legs, 0.0, angle_todo)
    hover_todo = np.where(legs, -s[:, 3] * 0.5, hover_todo)
    a = np.zeros(len(s), np.int32)
    a[angle_todo > 0.05] = 1
    a[angle_todo < -0.05] = 3
    a[(hover_todo > np.abs(angle_todo)) & (hover_todo > 0.05)] = 2
    return a


def assert_same(tag, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.dtype.kind == "f" or b.dtype.kind == "f":
        a = a.astype(np.float32); b = b.astype(np.float32)
        same = (a == b) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    if not same.all():
        bad = np.argwhere(~same)
        raise AssertionError("%s: %d of %d values differ; first at %s gpu=%r oracle=%r; max |diff| %.3e" % (
            tag, len(bad), same.size, tuple(bad[0]), a[tuple(bad[0])], b[tuple(bad[0])],
            float(np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.dtype.kind == "f" else 0.0))


@pytest.mark.parametrize("continuous", [False, True])
def test_reset_and_free_running_policy_bit_exact(continuous):
    n, k = 512, 400
    kind = O.LUNARLANDER_CONT if continuous else O.LUNARLANDER
    env = G.LunarLanderVecEnv(n, continuous=continuous, seed=1000, env_id_offset=5)
    ora = O.OracleEnv(kind, n, seed=1000, env_id_offset=5, mode=O.MODE_F32)
    obs = env.ResetBatch(); oobs = ora.reset()
    assert_same("reset obs", obs, oobs)
    rng = np.random.default_rng(3)
    alive = np.ones(n, bool)
    touched = 0
    for t in range(k):
        if continuous:
            a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
        else:
            a = np.where(rng.random(n) < 0.5, pid(obs), rng.integers(0, 4, n)).astype(np.int32)
        obs, rew, done = env.StepBatch(a)
        oobs, orew, odone = ora.step(a)
        assert_same("obs t=%d" % t, obs, oobs); assert_same("reward t=%d" % t, rew, orew); assert_same("done t=%d" % t, done, odone)
        touched += int((obs[:, 6:] > 0).any(axis=1).sum())
        if t % 50 == 49 or t == k - 1:
            st, ax, tt = env.GetState(); ost, oax, ott = ora.get_state()
            assert tt == ott
            assert_same("state t=%d" % t, st, ost); assert_same("aux t=%d" % t, ax, oax)
    assert touched > 0          # legs did reach the ground somewhere: the contact solver was exercised
    env.Close()


def test_three_lanes_per_lander_variant_bit_exact():
    """GYMCUDA_LUNAR_TRIO=1 (the contact class stepped by three lanes per lander, lunar_core.cuh "TRIO"; measured: no gain, so it
    is off by default): the same free-running comparison with the oracle, in a process of its own because the switch is read once."""
    import os, subprocess, sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle_lib as O, gymnet_b200 as G
from test_gpu_lunar import pid
n, k = 2048, 330
env = G.LunarLanderVecEnv(n, seed=1000, env_id_offset=5, auto_reset=True, time_limit=300)
ora = O.OracleEnv(O.LUNARLANDER, n, seed=1000, env_id_offset=5, auto_reset=True, time_limit=300, mode=O.MODE_F32)
obs = env.ResetBatch(); assert np.array_equal(obs, ora.reset())
rng = np.random.default_rng(3); touched = 0; episodes = 0
for t in range(k):
    a = np.where(rng.random(n) < 0.5, pid(obs), rng.integers(0, 4, n)).astype(np.int32)
    obs, rew, done = env.StepBatch(a)
    oobs, orew, odone = ora.step(a)
    assert np.array_equal(obs, oobs) and np.array_equal(rew, orew) and np.array_equal(done, odone), t
    touched += int((obs[:, 6:] > 0).any(axis=1).sum()); episodes += int((done != 0).sum())
st, ax, tt = env.GetState(); ost, oax, ott = ora.get_state()
assert tt == ott and np.array_equal(st, ost.astype(np.float32)) and np.array_equal(ax, oax)
assert touched > 0 and episodes > 0
print("TRIO-OK", touched, episodes)
""" % (os.path.dirname(os.path.abspath(__file__)), os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, GYMCUDA_LUNAR_TRIO="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "TRIO-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


def test_rollout_random_with_auto_reset_bit_exact():
    n, k = 1024, 300
    env = G.LunarLanderVecEnv(n, seed=7, auto_reset=True, time_limit=200)
    ora = O.OracleEnv(O.LUNARLANDER, n, seed=7, auto_reset=True, time_limit=200, mode=O.MODE_F32)
    assert_same("reset", env.ResetBatch(), ora.reset())
    obs, rew, done, act = env.RolloutRandom(k)
    oo, orr, od, oa = ora.rollout_random(k)
    assert_same("actions", act, oa); assert_same("done", done, od); assert_same("reward", rew, orr); assert_same("obs", obs, oo)
    st, ax, t = env.GetState(); ost, oax, ot = ora.get_state()
    assert_same("state", st, ost); assert_same("aux", ax, oax)
    assert done.sum() > n            # every env finished at least one episode (crash, landing or the limit)
    assert env.Stats()["episodes"] == int(done.sum())
    env.Close()


def test_teacher_forced_from_contact_states():
    """Snapshot oracle states during touchdown, force them into the GPU, one step, compare everything."""
    n = 256
    ora = O.OracleEnv(O.LUNARLANDER, n, seed=11, mode=O.MODE_F32)
    obs = ora.reset()
    snaps = []
    for t in range(260):
        obs, _, _ = ora.step(pid(obs))
        if t >= 60 and t % 10 == 0:
            snaps.append(ora.get_state())
    env = G.LunarLanderVecEnv(n, seed=11)
    env.ResetBatch()
    chk = O.OracleEnv(O.LUNARLANDER, n, seed=11, mode=O.MODE_F32); chk.reset()
    rng = np.random.default_rng(0)
    contacts = 0
    for st, ax, t in snaps:
        a = rng.integers(0, 4, n).astype(np.int32)
        env.SetState(st, ax, t); chk.set_state(st, ax, t)
        g = env.StepBatch(a); o = chk.step(a)
        for name, x, y in zip(("obs", "reward", "done"), g, o):
            assert_same(name, x, y)
        gs, ga, _ = env.GetState(); os_, oa, _ = chk.get_state()
        assert_same("state", gs, os_); assert_same("aux", ga, oa)
        contacts += int((ga[:, :3] != 0).any(axis=1).sum())
    assert contacts > 0
    env.Close()


def test_reference_pid_loop_and_quirks():
    """T3: the reference test's PID closed loop (LunarLanderEnvironment.cs:38-77) on 64 landers.  The
    reference's golden for ITS stream (seed 1000: 1547 steps, 184.01764) is not reproducible here; we
    check the loop terminates below MAX_STEPS like the reference asserts, and the env quirks."""
    n = 64
    env = G.LunarLanderVecEnv(n, seed=1000)
    obs = env.ResetBatch()
    total = np.zeros(n); steps = np.zeros(n, int); alive = np.ones(n, bool); last = np.zeros(n)
    for t in range(5000):
        obs, rew, done = env.StepBatch(pid(obs))
        total += np.where(alive, rew, 0.0); steps += alive
        last = np.where(alive & (done > 0), rew, last)
        alive &= done == 0
        if not alive.any():
            break
    # Assert.IsTrue(steps < MAX_STEPS) (:72) holds for the reference's single seeded lander; a batch also holds the
    # ones that drift out to the LEFT, which the reference never terminates (`pos.X > 1` only, :762)
    assert (steps < 5000).mean() >= 0.9
    finished = steps < 5000
    assert set(np.unique(last[finished])) <= {-100.0, 100.0}    # crash / out of view, or asleep after landing (:762-771)
    assert (last == 100.0).any()                      # some landers land and fall asleep (+100, :767-771)
    print("PID loop: mean steps %.0f, mean return %.1f, landed %d/%d (reference golden for its own stream: 1547 steps, 184.01764)"
          % (steps.mean(), total.mean(), int((last == 100.0).sum()), n))
    env.Close()


def test_invalid_action_is_rejected():
    env = G.LunarLanderVecEnv(8, seed=1)
    before = env.ResetBatch()
    with pytest.raises(G.InvalidActionError):         # LunarLanderEnv.cs:604-607
        env.StepBatch(np.array([0, 1, 2, 3, 4, -1, 0, 0], np.int32))
    after = env.Observe()
    assert np.array_equal(after[4], before[4]) and np.array_equal(after[5], before[5])
    assert not np.array_equal(after[0], before[0])
    env.Close()


def test_ctor_range_checks():
    with pytest.raises(ValueError):
        G.LunarLanderVecEnv(4, gravity=-13.0)         # LunarLanderEnv.cs:396-399
    with pytest.raises(ValueError):
        G.LunarLanderVecEnv(4, wind_power=25.0)       # :400-403


def test_wind_and_sharding():
    n = 256
    full = G.LunarLanderVecEnv(n, seed=3, auto_reset=True, time_limit=150, enable_wind=1)
    ora = O.OracleEnv(O.LUNARLANDER, n, seed=3, auto_reset=True, time_limit=150, mode=O.MODE_F32)
    O.lib().oracle_set_lunar_params(ora.h, -10.0, 1, 15.0, 1.5)
    assert_same("reset", full.ResetBatch(), ora.reset())
    fo, fr, fd, fa = full.RolloutRandom(200)
    oo, orr, od, oa = ora.rollout_random(200)
    assert_same("done", fd, od)
    assert np.abs(fo - oo).max() <= 1e-4              # wind goes through double tanh/sin: libm vs CUDA may differ in an ulp
    half = G.LunarLanderVecEnv(n // 2, seed=3, auto_reset=True, time_limit=150, enable_wind=1, env_id_offset=n // 2)
    half.ResetBatch()
    ho, hr, hd, ha = half.RolloutRandom(200)
    assert_same("shard obs", ho, fo[:, n // 2:]); assert_same("shard done", hd, fd[:, n // 2:])
    full.Close(); half.Close()
