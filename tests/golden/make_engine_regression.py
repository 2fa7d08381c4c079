"""Freezes the ENGINE arithmetic (oracle F32 mode == the CUDA kernels, bit for bit) as a regression fixture.

Unlike make_golden.py (an independent restatement of the reference's float64 arithmetic) these vectors are produced BY
the oracle: they do not prove anything about the reference.  They are a tripwire -- a change of detmath, of an env's
float32 formulation, of the RNG stream layout or of the LunarLander solver changes these bits, and has to be made on
purpose (re-run this script, say so in DESIGN.md section 5).

    python tests/golden/make_engine_regression.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

CASES = [("cartpole", O.CARTPOLE, 16, 200), ("pendulum", O.PENDULUM, 16, 250), ("mountaincar", O.MOUNTAINCAR, 16, 250),
         ("mountaincar_cont", O.MOUNTAINCAR_CONT, 16, 250), ("acrobot", O.ACROBOT, 16, 520),
         ("lunarlander", O.LUNARLANDER, 24, 320), ("lunarlander_cont", O.LUNARLANDER_CONT, 12, 200)]


def run(kind, n, k, seed=11):
    o = O.OracleEnv(kind, n, seed=seed, auto_reset=True, mode=O.MODE_F32)
    first = o.reset()
    obs, rew, done, act = o.rollout_random(k)
    st, aux, t = o.get_state()
    return {"first_obs": first, "obs_every_10": obs[::10].copy(), "obs_last": obs[-1].copy(), "reward_sum": rew.sum(0, dtype=np.float64),
            "done": np.packbits(done.astype(bool), axis=0), "actions_head": act[:16].copy(), "state": st.astype(np.float32), "aux": aux,
            "t": np.int64(t)}


def main():
    out = {}
    for name, kind, n, k in CASES:
        for key, v in run(kind, n, k).items():
            out["%s/%s" % (name, key)] = v
    np.savez_compressed(os.path.join(HERE, "engine_regression.npz"), **out)
    print("wrote engine_regression.npz:", sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main()
