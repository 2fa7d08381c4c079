"""CPU: the oracle's F32 mode (= the CUDA kernels' arithmetic, bit for bit) still produces the frozen vectors of
tests/golden/engine_regression.npz.  A tripwire against accidental changes of detmath, of an env's float32 formulation,
of the RNG stream layout or of the LunarLander solver; deliberate changes regenerate the fixture
(tests/golden/make_engine_regression.py).  The GPU side of the same statement is the live kernel-vs-oracle comparison of
tests/test_gpu_*.py."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_engine_regression as M  # noqa: E402

FIX = np.load(os.path.join(HERE, "golden", "engine_regression.npz"))


@pytest.mark.parametrize("name,kind,n,k", M.CASES, ids=[c[0] for c in M.CASES])
def test_engine_arithmetic_is_frozen(name, kind, n, k):
    got = M.run(kind, n, k)
    for key, v in got.items():
        want = FIX["%s/%s" % (name, key)]
        assert np.array_equal(np.asarray(v), want, equal_nan=True), "%s/%s differs from the frozen engine arithmetic" % (name, key)
    assert int(np.unpackbits(FIX["%s/done" % name], axis=0).sum()) > 0 or name in ("mountaincar_cont",)
