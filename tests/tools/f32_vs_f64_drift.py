#!/usr/bin/env python
"""SURVEY 8(c) T2, the free-running half: how far the engine arithmetic (float32 state, oracle F32 mode == the CUDA
kernels) drifts from the reference arithmetic (float64, oracle F64 mode) when both run FREE from the same initial
states under the same action sequences -- no teacher forcing.  Per env family: the maximum relative state error while
the two still agree on every `done`, and the histogram of the first step at which a `done` flag differs (after that
the episodes are different episodes and the comparison ends for that env).  CPU only (the oracle is the checker).

    python tests/tools/f32_vs_f64_drift.py [n_envs] [steps]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

SCALE = {"CartPole-v1": [2.4, 1.0, 0.21, 1.0], "Pendulum-v1": [3.14, 8.0], "MountainCar-v0": [1.2, 0.07],
         "MountainCarContinuous-v0": [1.2, 0.07], "Acrobot-v1": [3.14, 3.14, 12.0, 28.0]}
KINDS = {"CartPole-v1": O.CARTPOLE, "Pendulum-v1": O.PENDULUM, "MountainCar-v0": O.MOUNTAINCAR,
         "MountainCarContinuous-v0": O.MOUNTAINCAR_CONT, "Acrobot-v1": O.ACROBOT}


def drift(name, n, steps, seed=0):
    kind = KINDS[name]
    a32 = O.OracleEnv(kind, n, seed=seed, auto_reset=True, mode=O.MODE_F32)
    a64 = O.OracleEnv(kind, n, seed=seed, auto_reset=True, mode=O.MODE_F64)
    a32.reset(); a64.reset()
    scale = np.array(SCALE[name])
    alive = np.ones(n, bool)                 # still the same sequence of episodes
    first = np.full(n, -1)
    worst = 0.0
    for t in range(steps):
        a = a32.sample_actions(); a64.sample_actions()
        _, _, d32 = a32.step(a)
        _, _, d64 = a64.step(a)
        differ = alive & (d32 != d64)
        first[differ] = t
        alive &= ~differ
        if t % 8 == 0 or t == steps - 1:
            s32, _, _ = a32.get_state(); s64, _, _ = a64.get_state()
            diff = s32 - s64
            if name == "Acrobot-v1":
                diff[:, :2] = (diff[:, :2] + np.pi) % (2 * np.pi) - np.pi
            den = np.maximum(np.maximum(np.abs(s32), np.abs(s64)), scale)
            err = (np.abs(diff) / den).max(1)
            if alive.any():
                worst = max(worst, float(err[alive].max()))
    div = first[first >= 0]
    hist = np.histogram(div, bins=[0, 10, 30, 100, 300, 1000, 3000, 10 ** 9])[0].tolist() if div.size else [0] * 7
    return {"env": name, "envs": n, "steps": steps, "diverged": int(div.size), "frac_diverged": div.size / n,
            "first_divergence_hist": dict(zip(["<10", "<30", "<100", "<300", "<1000", "<3000", ">=3000"], hist)),
            "max_rel_state_err_while_in_step": worst}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    for name in KINDS:
        print(json.dumps(drift(name, n, steps)), flush=True)


if __name__ == "__main__":
    main()
