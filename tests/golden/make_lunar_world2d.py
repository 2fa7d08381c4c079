"""Golden single-step transitions of LunarLander produced by the structurally independent oracle (oracle/world2d: a generic
Box2D-2.3-lineage engine with LunarLanderEnv.cs built on top), in the kernel's state layout.

Two sets:  `ref/*`  the generic oracle on the reference's own rotation arithmetic, (float)Math.Sin((double)a) -- the engine
                    must agree on every integer word (contact flags, touching masks, limit states, contact ids, pair lists),
                    on done, and on the floats within tests/world2d_lib.py TOLERANCES;
           `det/*`  the generic oracle with the engine's deterministic float32 sincos plugged in -- the engine must agree
                    BIT FOR BIT.
The NumSharp stream of the reference cannot be reproduced, so the draws come from the engine's Philox stream; the step draws
of transition i are those of (seed 2024, env id i, step index 1000).

    python tests/golden/make_lunar_world2d.py     ->  tests/golden/lunar_world2d.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import world2d_lib as W  # noqa: E402

SEED, T = 2024, 1000


def main():
    out = {}
    for name, det, count in (("ref", 0, 1536), ("det", 1, 1024)):
        tr = W.generate_transitions(count, seed=SEED, T=T, landers=6, det_sincos=det, rng_seed=17 + det)
        cat = W.categories(tr["aux0"], tr["state0"])
        print(name, {k: int((cat == k).sum()) for k in sorted(set(cat))})
        for k, v in tr.items():
            out["%s/%s" % (name, k)] = v
    path = os.path.join(HERE, "lunar_world2d.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
