#!/usr/bin/env python
"""Replays a dump of the REAL reference (csharp/ReferenceReplay, run off-box where a .NET SDK exists) through the oracles
of this repo -- the step that turns "parity against our restatement" into "parity against the reference" (SURVEY 8f rank 1).
TEST / ANALYSIS TOOL: it drives oracle/ (and, with --gpu, libgymcuda through the C ABI); nothing in the product uses it.

    python tests/tools/replay_reference_dump.py <dump_dir> [--gpu] [--write-golden]

  cartpole_reference.csv   every teacher-forced transition against oracle F64 (the reference's arithmetic: done / reward /
                           steps_beyond_done exact, next state <= 1e-12 relative -- .NET's Math.Sin / Cos and libm may differ
                           in the last bit) and against the engine arithmetic (done exact, state <= 1e-5); --write-golden turns
                           the dump into tests/golden/cartpole_reference.npz, which tests/test_reference_replay_cpu.py and the
                           GPU fixture test then hold the oracle and the kernels to
  lunar_*_seed*.csv        the recorded draws + actions are fed to the generic engine (oracle/world2d) under every setting of
                           the open semantics (BeginContact returning false, TOI sub-stepping, ApplyForce at the origin): per
                           setting the first step whose observation leaves 1e-4, the step count and the return -- the setting
                           that reproduces the reference (golden: seed 1000 -> 1547 steps, 184.01764) is the one to freeze
"""
import csv
import glob
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def load_cartpole(path):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[0], np.array(rows[1:], dtype=np.float64)
    col = {h: i for i, h in enumerate(hdr)}
    pick = lambda *names: data[:, [col[n] for n in names]]   # noqa: E731
    return {"state": pick("x", "x_dot", "theta", "theta_dot"), "action": data[:, col["action"]].astype(np.int32),
            "sbd": data[:, col["sbd"]].astype(np.int32), "next_state": pick("nx", "nx_dot", "ntheta", "ntheta_dot"),
            "reward": data[:, col["reward"]].astype(np.float32), "done": data[:, col["done"]].astype(np.uint8),
            "next_sbd": data[:, col["next_sbd"]].astype(np.int32)}


def check_cartpole(fx, gpu=False):
    """Returns a dict of findings; raises AssertionError on a parity failure."""
    import oracle_lib as O
    n = len(fx["state"])
    aux = np.zeros((n, 3), np.int32); aux[:, 0] = fx["sbd"]
    out = {"transitions": n}
    for mode, tol, name in ((O.MODE_F64, 1e-12, "oracle_f64"), (O.MODE_F32, 1e-5, "engine_f32")):
        e = O.OracleEnv(O.CARTPOLE, n, seed=0, time_limit=-1, mode=mode)
        e.reset(); e.set_state(fx["state"], aux, 0)
        obs, rew, done = e.step(fx["action"])
        st, ax, _ = e.get_state()
        assert np.array_equal(done, fx["done"]), "%s: %d done flags differ from the reference" % (name, int((done != fx["done"]).sum()))
        assert np.array_equal(rew, fx["reward"]) and np.array_equal(ax[:, 0], fx["next_sbd"]), name + ": reward / steps_beyond_done differ"
        den = np.maximum(np.maximum(np.abs(st), np.abs(fx["next_state"])), [2.4, 1.0, 0.21, 1.0])
        err = float((np.abs(st - fx["next_state"]) / den).max())
        assert err <= tol, "%s: next state off by %.3g (tolerance %.1g)" % (name, err, tol)
        out[name + "_max_rel_err"] = err
        e.close()
    if gpu:
        import gymnet_b200 as G
        env = G.CartPoleVecEnv(n, seed=0, time_limit=-1)
        env.ResetBatch(); env.SetState(fx["state"].astype(np.float32), aux, 0)
        obs, rew, done = env.StepBatch(fx["action"])
        st, ax, _ = env.GetState()
        f32_in = np.array_equal(fx["state"].astype(np.float32).astype(np.float64), fx["state"])
        assert f32_in, "the dump's states are not float32-representable"
        assert np.array_equal(done, fx["done"]) and np.array_equal(rew, fx["reward"]) and np.array_equal(ax[:, 0], fx["next_sbd"])
        den = np.maximum(np.maximum(np.abs(st), np.abs(fx["next_state"])), [2.4, 1.0, 0.21, 1.0])
        out["gpu_max_rel_err"] = float((np.abs(st - fx["next_state"]) / den).max())
        assert out["gpu_max_rel_err"] <= 1e-5
        env.Close()
    return out


def load_lunar(path):
    meta, rows, total = {}, [], None
    for line in open(path):
        line = line.strip()
        if line.startswith("#"):
            for tok in line[1:].split():
                k, _, v = tok.partition("=")
                meta[k] = v
            continue
        if not line or line.startswith("kind,"):
            continue
        rows.append(line.split(","))
    reset = rows[0]
    assert reset[0] == "reset"
    ep = {"wind_idx": int(meta["wind_idx"]), "torque_idx": int(meta["torque_idx"]), "policy": meta.get("policy", "?"),
          "seed": int(meta.get("seed", -1)), "total_reward": float(meta["total_reward"]) if "total_reward" in meta else None,
          "zero_draws": np.array(reset[2:4], np.float32), "reset_obs": np.array(reset[4:12], np.float32),
          "reset_draws": np.array(reset[14:28], np.float32)}
    steps = rows[1:]
    ep["action"] = np.array([r[1] for r in steps], np.int32)
    ep["draws"] = np.array([r[2:4] for r in steps], np.float32)
    ep["obs"] = np.array([r[4:12] for r in steps], np.float32)
    ep["reward"] = np.array([r[12] for r in steps], np.float32)
    ep["done"] = np.array([r[13] for r in steps], np.int32)
    return ep


SWITCHES = ("begin_contact_false", "continuous_physics", "force_at_origin")


def replay_lunar(ep, atol=1e-4, **opts):
    """Feeds the recorded draws and actions to the generic engine; returns (first step whose observation differs by more than
    atol or -1, steps run, return, max |obs error| before the divergence)."""
    import world2d_lib as W
    w = W.LunarWorld(wind_idx=ep["wind_idx"], torque_idx=ep["torque_idx"], **opts)
    obs = w.reset(ep["reset_draws"], ep["zero_draws"])
    worst = float(np.abs(obs - ep["reset_obs"]).max())
    first = 0 if worst > atol else -1
    total = np.float32(0)
    steps = 0
    for k in range(len(ep["action"])):
        obs, r, d = w.step(int(ep["action"][k]), ep["draws"][k])
        total = np.float32(total + r)
        steps += 1
        err = float(np.abs(obs - ep["obs"][k]).max())
        if first < 0:
            if err > atol or d != int(ep["done"][k]) or abs(float(r) - float(ep["reward"][k])) > 1e-3 * max(1.0, abs(float(ep["reward"][k]))):
                first = k + 1
            else:
                worst = max(worst, err)
        if d:
            break
    w.close()
    return first, steps, float(total), worst


def lunar_report(ep):
    lines = []
    for vals in itertools.product((0, 1, 2), (0, 1), (0, 1)):
        opts = dict(zip(SWITCHES, vals))
        first, steps, total, worst = replay_lunar(ep, **opts)
        lines.append((opts, first, steps, total, worst))
    return lines


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if not args:
        print(__doc__); return 2
    d = args[0]
    gpu, write = "--gpu" in sys.argv, "--write-golden" in sys.argv
    cp = os.path.join(d, "cartpole_reference.csv")
    if os.path.exists(cp):
        fx = load_cartpole(cp)
        print("CartPole:", check_cartpole(fx, gpu=gpu))
        if write:
            np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cartpole_reference.npz"), **fx)
            print("wrote tests/golden/cartpole_reference.npz (%d reference-executed transitions)" % len(fx["state"]))
    for path in sorted(glob.glob(os.path.join(d, "lunar_*_seed*.csv"))):
        ep = load_lunar(path)
        print("%s: reference ran %d steps, return %s" % (os.path.basename(path), len(ep["action"]), ep["total_reward"]))
        for opts, first, steps, total, worst in lunar_report(ep):
            print("   %-70s first divergence %-6s steps %-5d return %10.4f  max |obs err| before it %.2e" % (
                opts, "none" if first < 0 else first, steps, total, worst))
    return 0


if __name__ == "__main__":
    sys.exit(main())
