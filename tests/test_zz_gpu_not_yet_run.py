"""GPU tests written after this round's last GPU session (the GPU budget was spent): verified on the host SIMT executor
or by construction, never yet run on a B200.  They live in the LAST test file of the suite and are non-strict xfail, so
that an unexpected failure -- or a CUDA error that poisons the process -- cannot hide or break the verified tests; a pass
shows up as XPASS: move the test to test_gpu_classic.py and drop the marker then."""
import numpy as np
import pytest

import gymnet_b200 as G

pytestmark = pytest.mark.gpu

NOT_YET_RUN_ON_GPU = pytest.mark.xfail(reason="added after the round's last GPU session; not yet run on a B200", strict=False)


@NOT_YET_RUN_ON_GPU
def test_registered_buffers_aligned_like_managed_arrays():
    """A GCHandle-pinned managed float[] starts 8 mod 16: the kernel's float4 observation stores cannot address it in
    place, so that one buffer is staged while the others stay zero-copy -- same results either way."""
    import ctypes as C
    from gymnet_b200 import _native as N
    n = 2500
    L = N.lib()
    a_env = G.CartPoleVecEnv(n, seed=11, auto_reset=True); a_env.ResetBatch()
    b_env = G.CartPoleVecEnv(n, seed=11, auto_reset=True); b_env.ResetBatch()
    def arr(shape, dtype, rem):          # data pointer == rem mod 16, inside whole pages no other buffer shares
        count = int(np.prod(shape)); item = np.dtype(dtype).itemsize
        raw = np.empty(count * item + 3 * 4096, np.uint8)
        page = (-raw.ctypes.data) % 4096
        view = raw[page + rem: page + rem + count * item].view(dtype).reshape(shape)
        assert view.ctypes.data % 16 == rem
        span = (rem + count * item + 4095) // 4096 * 4096          # the pages the view lives in: what gets page-locked
        return view, (raw.ctypes.data + page, span), raw
    act, r0, k0 = arr((n,), np.int32, 4); obs, r1, k1 = arr((n, 4), np.float32, 8)
    rew, r2, k2 = arr((n,), np.float32, 12); done, r3, k3 = arr((n,), np.uint8, 1)
    regions = [r0, r1, r2, r3]
    rng = np.random.default_rng(6)
    for phase in range(2):               # pageable, then registered
        if phase == 1:
            for base, span in regions:
                N.check(L.gymcuda_host_register(C.c_void_p(base), span))
        for _ in range(25):
            a = rng.integers(0, 2, n).astype(np.int32)
            act[:] = a
            N.check(L.gymcuda_step(a_env._h, C.c_void_p(act.ctypes.data), C.c_void_p(obs.ctypes.data),
                                   C.c_void_p(rew.ctypes.data), C.c_void_p(done.ctypes.data)))
            o, r, d = b_env.StepBatch(a)
            assert np.array_equal(obs, o) and np.array_equal(rew, r) and np.array_equal(done, d)
    for base, span in regions:
        N.check(L.gymcuda_host_unregister(C.c_void_p(base)))
    assert np.array_equal(a_env.Observe(), b_env.Observe())
    a_env.Close(); b_env.Close()


@NOT_YET_RUN_ON_GPU
def test_observation_and_reward_normalisation():
    """gymcuda_normalize against its numpy restatement: running statistics over many steps, in-place
    normalisation with clipping, an evaluation-mode call with frozen statistics, and the statistics read back."""
    from hostsim_lib import NormalizeModel
    n = 3000
    env = G.PendulumVecEnv(n, seed=2, auto_reset=True)
    env.ResetBatch()
    env.NormalizeConfig(0.95, 1e-6, 1.5, 5.0)
    model = NormalizeModel(n, 3, 0.95, 1e-6, 1.5, 5.0)
    rng = np.random.default_rng(4)
    for t in range(230):                                  # past Pendulum's 200-step limit: returns are reset at `done`
        a = rng.uniform(-2, 2, (n, 1)).astype(np.float32)
        obs, rew, done = env.StepBatch(a)
        update = t % 7 != 6
        want_o, want_r = model(obs, rew, done, update)
        got_o, got_r = obs.copy(), rew.copy()
        env.Normalize(got_o, got_r, done, update=update)
        assert np.allclose(got_o, want_o, rtol=1e-5, atol=1e-5), "observations at step %d" % t
        assert np.allclose(got_r, want_r, rtol=1e-5, atol=1e-5), "rewards at step %d" % t
    assert (np.abs(got_o) == 1.5).any()                   # the clip is active
    st = env.NormalizeStats()
    assert st["count"] == model.count
    mean = model.s / model.count
    assert np.allclose(st["obs_mean"], mean, rtol=1e-9, atol=1e-12)
    assert np.allclose(st["obs_var"], model.q / model.count - mean * mean, rtol=1e-7)
    mr = model.sr / model.count
    assert np.isclose(st["return_var"], model.qr / model.count - mr * mr, rtol=1e-7)
    env.NormalizeReset()
    assert env.NormalizeStats()["count"] == 0
    env.Close()
