"""GPU tests of the round-2 entry points, all through the C ABI against the CPU oracle:
terminal observations under auto-reset, k-step launches with caller-supplied actions, Box.Sample on device, rejected
actions of asynchronous steps, two handles driven from two host threads."""
import ctypes as C
import threading

import numpy as np
import pytest

import oracle_lib as O
import gymnet_b200 as G
from gymnet_b200 import _native as N
from helpers import KINDS, random_actions

pytestmark = pytest.mark.gpu

MAKE = {"CartPole-v1": G.CartPoleVecEnv, "Pendulum-v1": G.PendulumVecEnv, "MountainCar-v0": G.MountainCarVecEnv,
        "MountainCarContinuous-v0": G.MountainCarContinuousVecEnv, "Acrobot-v1": G.AcrobotVecEnv, "LunarLander-v2": G.LunarLanderVecEnv}


def oracle_pair(name, n, seed, off, limit):
    a = O.OracleEnv(KINDS[name], n, seed=seed, env_id_offset=off, auto_reset=True, mode=O.MODE_F32, time_limit=limit)
    b = O.OracleEnv(KINDS[name], n, seed=seed, env_id_offset=off, auto_reset=False, mode=O.MODE_F32, time_limit=limit)
    return a, b


@pytest.mark.parametrize("name,n,limit,k,where", [("CartPole-v1", 5000, 0, 80, "pageable"), ("CartPole-v1", 4096, 0, 60, "pinned"),
                                                  ("MountainCar-v0", 1000, 30, 70, "pageable"), ("Acrobot-v1", 1500, 25, 60, "pinned"),
                                                  ("Pendulum-v1", 999, 10, 25, "pageable"), ("LunarLander-v2", 600, 120, 260, "pageable")])
def test_terminal_observations_under_auto_reset(name, n, limit, k, where):
    """SURVEY 8b auto-reset row: rows of the side buffer of the envs whose step returned done hold the observation of the state
    the episode ended in -- the oracle stepped WITHOUT auto-reset from the same state -- bit for bit; other rows are untouched;
    the step's own outputs (post-reset observation, terminal reward, done) are unchanged by the side buffer."""
    rng = np.random.default_rng(3)
    a, b = oracle_pair(name, n, 9, 123, limit)
    env = MAKE[name](n, seed=9, env_id_offset=123, auto_reset=True, time_limit=limit)
    assert np.array_equal(env.ResetBatch(), a.reset()); b.reset()
    L = N.lib()
    pinned = None
    if where == "pinned":
        pinned = C.c_void_p()
        N.check(L.gymcuda_host_alloc(C.byref(pinned), n * env.obs_dim * 4))
        term = np.ctypeslib.as_array(C.cast(pinned, C.POINTER(C.c_float)), shape=(n * env.obs_dim,)).reshape(n, env.obs_dim)
    else:
        term = np.empty((n, env.obs_dim), np.float32)
    term[:] = -7.0
    env.SetTerminalObs(term)
    seen = 0
    for _ in range(k):
        act = random_actions(env, rng, n)
        st, aux, t = a.get_state()
        b.set_state(st, aux, t)
        wo, wr, wd = a.step(act)
        to, _, _ = b.step(act)
        before = term.copy()
        go, gr, gd = env.StepBatch(act)
        assert np.array_equal(go, wo) and np.array_equal(gr, wr) and np.array_equal(gd, wd)
        d = wd != 0
        assert np.array_equal(term[d], to[d]) and np.array_equal(term[~d], before[~d])
        seen += int(d.sum())
    assert seen > n // 50
    env.SetTerminalObs(None)
    before = term.copy()
    env.StepBatch(random_actions(env, rng, n))
    assert np.array_equal(term, before)
    env.Close()
    if pinned is not None:
        N.check(L.gymcuda_host_free(pinned))


def test_terminal_observations_need_auto_reset_and_device_buffers_work():
    import torch
    env = G.CartPoleVecEnv(64, seed=1)
    with pytest.raises(ValueError):
        env.SetTerminalObs(np.zeros((64, 4), np.float32))
    env.Close()
    n = 3000
    a, b = oracle_pair("CartPole-v1", n, 4, 0, 0)
    env = G.CartPoleVecEnv(n, seed=4, auto_reset=True); env.ResetBatch(); a.reset(); b.reset()
    dev = torch.device("cuda", 0)
    term = torch.full((n, 4), -7.0, device=dev)
    env.SetTerminalObs(term.data_ptr())
    obs = torch.empty((n, 4), device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(8)
    for _ in range(40):
        act = rng.integers(0, 2, n).astype(np.int32)
        st, aux, t = a.get_state(); b.set_state(st, aux, t)
        _, _, wd = a.step(act); to, _, _ = b.step(act)
        d_act = torch.from_numpy(act).to(dev)
        env.StepDevice(d_act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr()); env.Sync()
        d = wd != 0
        assert np.array_equal(done.cpu().numpy(), wd)
        assert np.array_equal(term.cpu().numpy()[d], to[d])
    env.Close()


@pytest.mark.parametrize("name,n,k", [("CartPole-v1", 4100, 150), ("Pendulum-v1", 1000, 230), ("MountainCar-v0", 777, 250),
                                      ("MountainCarContinuous-v0", 640, 64), ("Acrobot-v1", 1000, 90), ("LunarLander-v2", 300, 140)])
def test_step_many_equals_the_oracle_stepped_k_times(name, n, k):
    """gymcuda_step_many: one launch of k steps with the caller's actions == k oracle steps (auto-reset, default time limits,
    LunarLander with 100); then again split in two calls (state carried across launches)."""
    rng = np.random.default_rng(31)
    limit = 100 if name == "LunarLander-v2" else 0
    o = O.OracleEnv(KINDS[name], n, seed=2, env_id_offset=50, auto_reset=True, mode=O.MODE_F32, time_limit=limit)
    env = MAKE[name](n, seed=2, env_id_offset=50, auto_reset=True, time_limit=limit)
    assert np.array_equal(env.ResetBatch(), o.reset())
    for kk in (k, k // 3, k - k // 3):
        acts = np.stack([random_actions(env, rng, n) for _ in range(kk)])
        obs, rew, done = env.StepMany(acts)
        for j in range(kk):
            wo, wr, wd = o.step(acts[j])
            assert np.array_equal(obs[j], wo) and np.array_equal(rew[j], wr) and np.array_equal(done[j], wd), (name, j)
        st, aux, t = o.get_state()
        gs, ga, gt = env.GetState()
        assert gt == t and np.array_equal(gs, st.astype(np.float32))
    env.Close()


@pytest.mark.parametrize("name,n", [("CartPole-v1", 65536), ("Pendulum-v1", 60000), ("MountainCar-v0", 70000)])
def test_step_many_one_wave_shape_and_partial_outputs(name, n):
    """gymcuda_step_many at a batch that takes the one-wave 512-thread shape of the chunked kernel (launched at an unaligned step
    index: head, chunks, tail), and the same steps with only some outputs requested (the generic variant): both == the oracle."""
    rng = np.random.default_rng(77)
    o = O.OracleEnv(KINDS[name], n, seed=9, env_id_offset=11, auto_reset=True, mode=O.MODE_F32)
    env = MAKE[name](n, seed=9, env_id_offset=11, auto_reset=True)
    assert np.array_equal(env.ResetBatch(), o.reset())
    L = N.lib()
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    od = env.obs_dim
    for kk, partial in ((3, False), (29, False), (13, True), (16, False)):
        acts = np.stack([random_actions(env, rng, n) for _ in range(kk)])
        obs = np.empty((kk, n, od), np.float32); rew = np.empty((kk, n), np.float32); done = np.empty((kk, n), np.uint8)
        rc = L.gymcuda_step_many(env._h, kk, p(acts), None if partial else p(obs), p(rew), p(done))
        assert rc == 0, L.gymcuda_last_error()
        for j in range(kk):
            wo, wr, wd = o.step(acts[j])
            assert np.array_equal(rew[j], wr) and np.array_equal(done[j], wd), (name, kk, j)
            if not partial:
                assert np.array_equal(obs[j], wo), (name, kk, j)
        st, aux, t = o.get_state()
        gs, ga, gt = env.GetState()
        assert gt == t and np.array_equal(gs, st.astype(np.float32))
    env.Close()


@pytest.mark.parametrize("name", ["CartPole-v1", "MountainCar-v0", "MountainCarContinuous-v0"])
def test_step_device_without_observation_copy_and_obs_view(name):
    """gymcuda_step_device(d_obs = GYMCUDA_NO_OBS): no observation is written; gymcuda_obs_view_device points at the state
    array, which IS the current observation of these env kinds -- equal to the oracle's observations after every step,
    auto-resets included; Pendulum has no such view (EINVAL) and gymcuda_observe still recomputes."""
    import torch
    n = 3000
    rng = np.random.default_rng(8)
    limit = 50 if name == "MountainCarContinuous-v0" else 0      # (its episodes last 999 steps otherwise)
    o = O.OracleEnv(KINDS[name], n, seed=4, auto_reset=True, mode=O.MODE_F32, time_limit=limit)
    env = MAKE[name](n, seed=4, auto_reset=True, time_limit=limit)
    assert np.array_equal(env.ResetBatch(), o.reset())
    dev = torch.device("cuda", 0)
    od = env.obs_dim
    view = env.ObsViewDevice()
    rt = C.CDLL("/usr/local/cuda/lib64/libcudart.so")
    rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
    canary = torch.full((n, od), 7.0, device=dev)
    seen = 0
    for t in range(260):
        a = random_actions(env, rng, n)
        wo, wr, wd = o.step(a)
        d_a = torch.from_numpy(a).to(dev)
        env.StepDevice(d_a.data_ptr(), env.NO_OBS, rew.data_ptr(), done.data_ptr()); env.Sync()
        got = np.empty((n, od), np.float32)
        assert rt.cudaMemcpy(C.c_void_p(got.ctypes.data), C.c_void_p(view), C.c_size_t(got.nbytes), 2) == 0   # cudaMemcpyDeviceToHost
        assert np.array_equal(got, wo) and np.array_equal(rew.cpu().numpy(), wr) and np.array_equal(done.cpu().numpy(), wd), t
        seen += int(wd.sum())
    assert seen > 0 and bool((canary == 7.0).all())
    assert np.array_equal(env.Observe(), wo)
    env.Close()
    pend = G.PendulumVecEnv(8, seed=1, auto_reset=True); pend.ResetBatch()
    with pytest.raises(ValueError):          # GYMCUDA_EINVAL (C#: ArgumentException)
        pend.ObsViewDevice()
    pend.Close()


@pytest.mark.parametrize("name,limit", [("CartPole-v1", 0), ("MountainCar-v0", 9), ("Pendulum-v1", 7)])
def test_million_env_batch(name, limit):
    """Batches of 2^20 envs and more (their own launch shape, STEP_BLOCK_BIG): ragged size, == the oracle incl. auto-resets, the time limit firing for every env at once, the done list, and the
    no-observation-copy mode."""
    import torch
    n = (1 << 20) + 77
    rng = np.random.default_rng(12)
    o = O.OracleEnv(KINDS[name], n, seed=6, env_id_offset=3, auto_reset=True, mode=O.MODE_F32, time_limit=limit)
    env = MAKE[name](n, seed=6, env_id_offset=3, auto_reset=True, time_limit=limit)
    assert np.array_equal(env.ResetBatch(), o.reset())
    episodes = 0
    for t in range(24):
        a = random_actions(env, rng, n)
        wo, wr, wd = o.step(a)
        go, gr, gd = env.StepBatch(a)
        assert np.array_equal(gd, wd) and np.array_equal(gr, wr) and np.array_equal(go, wo), (name, t)
        episodes += int(wd.sum())
        if t in (8, 20):
            assert np.array_equal(np.sort(env.DoneIndices()), np.nonzero(wd)[0])
    assert episodes > 0 and env.Stats()["episodes"] == episodes
    if name != "Pendulum-v1":
        dev = torch.device("cuda", 0)
        rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
        for t in range(6):
            a = random_actions(env, rng, n)
            wo, wr, wd = o.step(a)
            d_a = torch.from_numpy(a).to(dev)
            env.StepDevice(d_a.data_ptr(), env.NO_OBS, rew.data_ptr(), done.data_ptr()); env.Sync()
            assert np.array_equal(done.cpu().numpy(), wd) and np.array_equal(rew.cpu().numpy(), wr)
        assert np.array_equal(env.Observe(), wo)
    st, aux, t = o.get_state()
    gs, ga, gt = env.GetState()
    assert gt == t and np.array_equal(gs, st.astype(np.float32))
    env.Close()


def test_step_many_rejects_invalid_actions_per_step():
    n, k = 500, 20
    rng = np.random.default_rng(5)
    o = O.OracleEnv(O.ACROBOT, n, seed=3, auto_reset=True, mode=O.MODE_F32)
    env = G.AcrobotVecEnv(n, seed=3, auto_reset=True)
    assert np.array_equal(env.ResetBatch(), o.reset())
    acts = rng.integers(0, 3, (k, n)).astype(np.int32)
    bad = rng.random((k, n)) < 0.02
    acts[bad] = -1
    L = N.lib()
    obs = np.empty((k, n, 6), np.float32); rew = np.empty((k, n), np.float32); done = np.empty((k, n), np.uint8)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    rc = L.gymcuda_step_many(env._h, k, p(acts), p(obs), p(rew), p(done))
    assert rc == N.EACTION and str(int(bad.sum())).encode() in L.gymcuda_last_error()
    for j in range(k):
        wo, wr, wd = o.step(acts[j])
        ok = ~bad[j]
        assert np.array_equal(obs[j][ok], wo[ok]) and np.array_equal(rew[j][ok], wr[ok]) and np.array_equal(done[j][ok], wd[ok])
        assert np.array_equal(obs[j][bad[j]], wo[bad[j]]) and not rew[j][bad[j]].any() and not done[j][bad[j]].any()
    assert env.Stats()["invalid_actions"] == int(bad.sum())
    env.Close()


def test_sync_reports_actions_rejected_by_asynchronous_steps_and_host_steps_do_not_inherit_them():
    """an invalid action in gymcuda_step_device used to surface as GYMCUDA_EACTION of the NEXT host-buffer step."""
    import torch
    n = 256
    env = G.MountainCarVecEnv(n, seed=1, auto_reset=True); env.ResetBatch()
    dev = torch.device("cuda", 0)
    act = torch.ones(n, dtype=torch.int32, device=dev); act[7] = 9
    obs = torch.empty((n, 2), device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
    env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
    with pytest.raises(G.InvalidActionError):
        env.Sync()
    env.Sync()                                      # reported once
    env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
    env.StepBatch(np.ones(n, np.int32))             # valid host step right after an unreported bad device step: no error
    with pytest.raises(G.InvalidActionError):
        env.StepBatch(np.full(n, 5, np.int32))
    assert env.Stats()["invalid_actions"] == 2 + n
    env.Close()


def box_sample_reference(low, high, count, seed, index):
    """float64 restatement of box_sample.cuh (the reference's four-way split, Box.cs:81-84) on the same Philox words."""
    dim = low.size
    out = np.empty((count, dim))
    for j in range(dim):
        for c in range(count):
            w = O.draw(seed, j, index + c, 4)
            lo, hi = float(low[j]), float(high[j])
            bl, bh = np.isfinite(lo), np.isfinite(hi)
            u = float(w[0] >> 8) * 2.0 ** -24
            if bl and bh:
                out[c, j] = lo + float(np.float32(w[0] >> 8) * np.float32(np.float32(hi - lo) * np.float32(2.0 ** -24)))
            elif bl or bh:
                out[c, j] = (lo if bl else hi) - np.log1p(-u)
            else:
                u1 = (float(w[1] >> 8) + 1.0) * 2.0 ** -24
                u2 = float(w[2] >> 8) * 2.0 ** -24
                out[c, j] = 0.5 + np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return out


def test_box_sample_on_device_four_way_split():
    """Box.cs:69-90 on device: uniform / low + exp / high + exp (the reference adds to High) / normal(0.5, 1); values against a
    float64 restatement on the same Philox words, distributions by Kolmogorov-Smirnov, integer dtype floors."""
    from scipy import stats
    inf = np.inf
    low = np.array([-2.0, 1.5, -inf, -inf, 0.0], np.float32)
    high = np.array([3.0, inf, -4.0, inf, 1.0], np.float32)
    box = G.Box(low, high)
    small = box.SampleBatch(400, seed=77, index=5)
    want = box_sample_reference(low, high, 400, 77, 5)
    assert small.shape == (400, 5) and np.allclose(small, want, rtol=2e-6, atol=2e-6)
    assert np.array_equal(box.SampleBatch(100, seed=77, index=105), box.SampleBatch(400, seed=77, index=5)[100:200])   # pure function of (seed, index + c, j)
    big = box.SampleBatch(200000, seed=1)
    assert (big[:, 0] >= -2).all() and (big[:, 0] < 3).all() and (big[:, 1] >= 1.5).all() and (big[:, 2] >= -4.0).all()
    for col, dist in ((big[:, 0], stats.uniform(-2, 5)), (big[:, 1], stats.expon(1.5)), (big[:, 2], stats.expon(-4.0)),
                      (big[:, 3], stats.norm(0.5, 1.0)), (big[:, 4], stats.uniform(0, 1))):
        assert stats.kstest(col.astype(np.float64), dist.cdf).pvalue > 1e-3
    ib = G.Box(np.array([0, -5]), np.array([10, 5]), dtype=np.int32)           # integer dtype: floor (Box.cs:85-88)
    si = ib.SampleBatch(1000, seed=3)
    assert si.dtype == np.int32 and si[:, 0].min() >= 0 and si[:, 0].max() <= 9 and si[:, 1].min() >= -5 and si[:, 1].max() <= 4
    fl = G.Box(np.array([0.0, -5.0], np.float32), np.array([10.0, 5.0], np.float32)).SampleBatch(1000, seed=3)
    assert np.array_equal(si, np.floor(fl).astype(np.int32))
    L = N.lib()
    assert L.gymcuda_box_sample(0, 0, 0, None, None, 2, 2, 0, None) == N.EINVAL


def test_two_handles_from_two_host_threads():
    """The reference's tests run two env instances on two tasks (tests/Gym.Tests/Envs/Aether/LunarLanderEnvironment.cs:185-190):
    distinct handles are independent -- two threads stepping their own handle concurrently (thread-local error text, own stream,
    own device buffers) get exactly what each gets alone."""
    n, k = 2048, 120
    results, errors = {}, []

    def work(tag, name, seed):
        try:
            rng = np.random.default_rng(seed)
            env = MAKE[name](n, seed=seed, auto_reset=True)
            out = [env.ResetBatch()]
            for j in range(k):
                o, r, d = env.StepBatch(random_actions(env, rng, n))
                out.append(o); out.append(r); out.append(d)
                if j == k // 2:   # an error raised in this thread must carry THIS thread's text
                    try:
                        env.StepBatch(np.full(n, 99, np.int32) if env.act_n > 0 else np.full((n, env.act_dim), np.nan, np.float32))
                        if name != "CartPole-v1":
                            errors.append("%s: invalid actions not reported" % tag)
                    except G.InvalidActionError as ex:
                        if str(n) not in str(ex):
                            errors.append("%s: wrong error text %s" % (tag, ex))
            env.Close()
            results[tag] = out
        except Exception as ex:   # noqa: BLE001
            errors.append("%s: %r" % (tag, ex))

    specs = [("a", "LunarLander-v2", 11), ("b", "Acrobot-v1", 12)]
    for tag, name, seed in specs:       # alone
        work(tag + "_alone", name, seed)
    threads = [threading.Thread(target=work, args=(tag, name, seed)) for tag, name, seed in specs]
    for t in threads: t.start()
    for t in threads: t.join()
    assert not errors, errors
    for tag, _, _ in specs:
        assert len(results[tag]) == len(results[tag + "_alone"])
        assert all(np.array_equal(x, y) for x, y in zip(results[tag], results[tag + "_alone"]))


@pytest.mark.parametrize("name,n", [("CartPole-v1", 65536), ("Acrobot-v1", 4096), ("LunarLander-v2", 2048)])
def test_step_device_under_cuda_graph_capture_and_replay(name, n):
    """gymcuda_set_device_clock: a StepDevice captured into a CUDA graph and replayed k times == k steps of the oracle with the
    same actions (LunarLander's per-step dispersion draws are keyed by the step index, which the replay must advance; the
    done list and the episode counter must follow the replays too); then ordinary launches continue from the replayed state."""
    import torch
    k = 60
    limit = 40 if name == "LunarLander-v2" else 0
    o = O.OracleEnv(KINDS[name], n, seed=6, env_id_offset=10, auto_reset=True, mode=O.MODE_F32, time_limit=limit)
    env = MAKE[name](n, seed=6, env_id_offset=10, auto_reset=True, time_limit=limit)
    assert np.array_equal(env.ResetBatch(), o.reset())
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(device=dev)
    env.SetStream(s.cuda_stream)
    env.SetDeviceClock(True)
    act = torch.zeros(n, dtype=torch.int32, device=dev)
    obs = torch.empty((n, env.obs_dim), device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(12)
    a0 = rng.integers(0, env.act_n, n).astype(np.int32)
    with torch.cuda.stream(s):               # one ordinary launch first (also warms the lazily created resources)
        act.copy_(torch.from_numpy(a0).to(dev))
        env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
    s.synchronize()
    wo, wr, wd = o.step(a0)
    assert np.array_equal(obs.cpu().numpy(), wo)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
    episodes = int(wd.sum())
    for j in range(k):
        a = rng.integers(0, env.act_n, n).astype(np.int32)
        act.copy_(torch.from_numpy(a).to(dev))
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        wo, wr, wd = o.step(a)
        assert np.array_equal(done.cpu().numpy(), wd), (name, j)
        assert np.array_equal(obs.cpu().numpy(), wo) and np.array_equal(rew.cpu().numpy(), wr), (name, j)
        episodes += int(wd.sum())
        if j % 20 == 7:
            assert np.array_equal(np.sort(env.DoneIndices()), np.nonzero(wd)[0])
    assert env.Stats()["episodes"] == episodes
    st, aux, t = o.get_state()
    gs, ga, gt = env.GetState()
    assert gt == t and np.array_equal(gs, st.astype(np.float32))
    env.SetDeviceClock(False)               # back to host counters: ordinary steps continue the same trajectory
    for _ in range(5):
        a = rng.integers(0, env.act_n, n).astype(np.int32)
        go, gr, gd = env.StepBatch(a)
        wo, wr, wd = o.step(a)
        assert np.array_equal(go, wo) and np.array_equal(gd, wd)
    env.Close()


def test_render_cartpole_and_lunarlander_headless_batched():
    """Env.Render rasterised on the device for a subset of the batch, against the numpy restatement of the same
    primitives / colours / draw order (tests/render_ref.py), at the reference's 600 x 400 and at a down-sampled size."""
    import render_ref as RR
    env = G.CartPoleVecEnv(64, seed=3, auto_reset=True); env.ResetBatch()
    rng = np.random.default_rng(0)
    for _ in range(12):
        env.StepBatch(rng.integers(0, 2, 64).astype(np.int32))
    st, _, _ = env.GetState()
    ids = np.array([0, 17, 63], np.int32)
    for (w, h) in ((600, 400), (150, 100)):
        got = env.Render(ids, w, h)
        assert got.shape == (3, h, w, 3) and got.dtype == np.uint8
        for k, e in enumerate(ids):
            want = RR.cartpole(st[e], w, h)
            assert (got[k] != want).any(axis=2).mean() < 1e-3
        colours = {tuple(c) for c in got[0].reshape(-1, 3)}
        assert colours == {(255, 255, 255), (0, 0, 0), (204, 153, 102)}
    first = env.Render(None, 600, 400, count=2)
    assert np.array_equal(first[1], env.Render(np.array([1], np.int32))[0])
    env.Close()
    ll = G.LunarLanderVecEnv(48, seed=5, auto_reset=True); ll.ResetBatch()
    for _ in range(30):
        ll.StepBatch(rng.integers(0, 4, 48).astype(np.int32))
    st, _, _ = ll.GetState()
    ids = np.array([3, 40], np.int32)
    for (w, h) in ((600, 400), (84, 84)):
        got = ll.Render(ids, w, h)
        for k, e in enumerate(ids):
            want = RR.lunar(st[e], w, h)
            assert (got[k] != want).any(axis=2).mean() < 1e-3
    full = ll.Render(ids[:1], 600, 400)[0]
    colours = {tuple(c) for c in full.reshape(-1, 3)}
    assert {(0, 0, 0), (255, 255, 255), (255, 0, 0), (128, 102, 230), (204, 204, 0)} <= colours
    assert (full[396] == 255).all(axis=1).mean() > 0.9                                   # ground near the bottom (the last row is the red base edge)
    odd = ll.Render(ids[:1], 83, 61)[0]                                                  # pixel count not a multiple of 4: byte stores
    assert (odd != RR.lunar(st[ids[0]], 83, 61)).any(axis=2).mean() < 2e-3
    ll.Close()
    pend = G.PendulumVecEnv(4); pend.ResetBatch()
    with pytest.raises(ValueError):
        pend.Render(None, 64, 64, count=1)               # the reference has no Render for it
    pend.Close()


@pytest.mark.parametrize("n", [1, 33, 129, 1000])
def test_round2_entry_points_on_small_and_ragged_batches(n):
    """Batches of 1, 33, 129 and 1000 envs (one lane, a ragged second warp, a ragged second CTA, eight CTAs): the done list built
    on demand, the deferred CTA-level reset pass, the terminal-observation buffer and step_many against the oracle."""
    rng = np.random.default_rng(n)
    a, b = oracle_pair("CartPole-v1", n, 7, 3, 12)
    env = G.CartPoleVecEnv(n, seed=7, env_id_offset=3, auto_reset=True, time_limit=12)
    assert np.array_equal(env.ResetBatch(), a.reset()); b.reset()
    term = np.full((n, 4), -1.0, np.float32)
    env.SetTerminalObs(term)
    for _ in range(40):
        act = rng.integers(0, 2, n).astype(np.int32)
        st, aux, t = a.get_state(); b.set_state(st, aux, t)
        wo, wr, wd = a.step(act); to, _, _ = b.step(act)
        go, gr, gd = env.StepBatch(act)
        assert np.array_equal(go, wo) and np.array_equal(gr, wr) and np.array_equal(gd, wd)
        d = wd != 0
        assert np.array_equal(term[d], to[d])
        assert np.array_equal(np.sort(env.DoneIndices()), np.nonzero(wd)[0])
        assert np.array_equal(np.sort(env.DoneIndices()), np.nonzero(wd)[0])     # asked twice: built once, same list
    env.SetTerminalObs(None)
    acts = rng.integers(0, 2, (25, n)).astype(np.int32)
    obs, rew, done = env.StepMany(acts)
    for j in range(25):
        wo, wr, wd = a.step(acts[j])
        assert np.array_equal(obs[j], wo) and np.array_equal(done[j], wd)
    st, _, t = a.get_state(); gs, _, gt = env.GetState()
    assert gt == t and np.array_equal(gs, st.astype(np.float32))
    assert env.Stats()["episodes"] > 0
    env.Close()
    ll = G.LunarLanderVecEnv(n, seed=2, auto_reset=True, time_limit=15)      # n < 512: LunarLander's single-launch path
    lt = O.OracleEnv(O.LUNARLANDER, n, seed=2, auto_reset=True, mode=O.MODE_F32, time_limit=15)
    assert np.array_equal(ll.ResetBatch(), lt.reset())
    for _ in range(20):
        act = rng.integers(0, 4, n).astype(np.int32)
        go, gr, gd = ll.StepBatch(act); wo, wr, wd = lt.step(act)
        assert np.array_equal(go, wo) and np.array_equal(gd, wd)
        assert np.array_equal(np.sort(ll.DoneIndices()), np.nonzero(wd)[0])
    ll.Close()
