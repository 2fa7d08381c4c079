"""CPU: the engine's KERNELS (kernels.cuh), compiled for the host and run as CTAs of cooperating fibers, against the
oracle -- bit for bit.

tests/hostsim/simt.hpp executes the threads of a CTA as ucontext fibers and makes __ballot_sync / __all_sync /
__shfl_*_sync / __syncwarp / __activemask / __syncthreads real rendezvous between them, so the warp-cooperative parts
of the kernels run as written on a machine without a GPU: the done compaction and the packed done bytes of
step_kernel, the statistics reductions, the staged observation stores and the warp votes of rollout_kernel in its 64-
and 512-thread shapes, ragged last warps, and the three kernels of the LunarLander contact partition.  (Test
infrastructure: nothing here is part of the product, and the GPU tests remain the parity tests proper.)"""
import numpy as np
import pytest

import oracle_lib as O

from hostsim_lib import HostSim, build

F32 = np.float32


@pytest.fixture(scope="module")
def lib():
    return build()


# Every test runs under three thread schedules of the executor: runnable lanes proceed in ascending order, in descending
# order, and in a seeded pseudo-random order.  Correct kernels give identical results under all of them; a kernel that
# relies on an ordering only a missing __syncwarp / __syncthreads would provide breaks under at least one.
@pytest.fixture(autouse=True, params=["ascending", "descending", "random"])
def hs(lib, request):
    lib.hostsim_set_simt(1)
    lib.hostsim_set_schedule(["ascending", "descending", "random"].index(request.param), 12345)
    yield lib
    lib.hostsim_set_schedule(0, 0)
    lib.hostsim_set_simt(0)


def pair(hs, kind, n, seed=5, off=77, done_bits=False, time_limit=0):
    o = O.OracleEnv(kind, n, seed=seed, env_id_offset=off, auto_reset=True, mode=O.MODE_F32, done_bits=done_bits, time_limit=time_limit)
    sim = HostSim(hs, kind, n, seed, off)
    if time_limit:
        sim.limit = time_limit
    assert np.array_equal(sim.reset_kernel(), o.reset())
    return o, sim


def same_state(o, sim):
    st, _, ot = o.get_state()
    return ot == sim.t and np.array_equal(sim.abi_state(), st.astype(F32))


# n chosen so that 3- and 6-float observations take the staged path (n % 4 == 0) in the full warps and the per-lane
# path in a ragged last warp (Acrobot 100 = 3 full warps + 4 lanes), and so that the 512-thread shape has idle warps
ROLLOUT_CASES = [("CartPole-v1", O.CARTPOLE, 200), ("Pendulum-v1", O.PENDULUM, 128), ("Pendulum-v1 ragged", O.PENDULUM, 76),
                 ("MountainCar-v0", O.MOUNTAINCAR, 77), ("MountainCarContinuous-v0", O.MOUNTAINCAR_CONT, 64),
                 ("Acrobot-v1", O.ACROBOT, 100), ("LunarLander-v2", O.LUNARLANDER, 70)]


@pytest.mark.parametrize("name,kind,n", ROLLOUT_CASES, ids=[c[0] for c in ROLLOUT_CASES])
@pytest.mark.parametrize("block", [64, 512])
def test_rollout_kernel_all_outputs_with_real_warps(hs, name, kind, n, block):
    if block == 512 and kind in (O.ACROBOT, O.LUNARLANDER):
        pytest.skip("no one-wave shape for envs whose step is not unrolled")
    o, sim = pair(hs, kind, n)
    launches = (3, 37, 100) if kind == O.LUNARLANDER else (3, 37, 8, 210, 260)
    for k in launches:
        tr = sim.rollout(k, True, block)
        tw = o.rollout_random(k)
        for j, what in enumerate(("obs", "reward", "done", "actions")):
            assert np.array_equal(tr[j], tw[j]), "%s differs in the launch of %d steps" % (what, k)
        assert tr[4] == int(tw[2].sum())          # the warp-reduced episode counter
        assert same_state(o, sim)


@pytest.mark.parametrize("name,kind,n", [("CartPole-v1", O.CARTPOLE, 100), ("Pendulum-v1", O.PENDULUM, 96), ("MountainCar-v0", O.MOUNTAINCAR, 70),
                                         ("Acrobot-v1", O.ACROBOT, 64)], ids=["CartPole-v1", "Pendulum-v1", "MountainCar-v0", "Acrobot-v1"])
def test_rollout_kernel_generic_shape_statistics_and_truncation_bits(hs, name, kind, n):
    """The generic variant: episode return / length sums (per-thread double accumulators, warp shuffles, one atomic pair
    per warp), the running return carried in ep_ret across launches, and done bytes that tell truncation (2) from
    termination (1)."""
    o, sim = pair(hs, kind, n, done_bits=True)
    if sim.limit == 0:
        sim.limit = 0x7fffffff            # what gymcuda_create does for episode statistics without a time limit
    ep_ret = np.zeros(n, F32); sums = np.zeros(2, np.float64)
    ret = np.zeros(n, F32); length = np.zeros(n, np.int64); want_ret = 0.0; want_len = 0; episodes = 0
    for k in (5, 130, 260, 300):
        tr = sim.rollout(k, False, 64, ep_ret=ep_ret, sums=sums, done_bits=1)
        tw = o.rollout_random(k)
        for j, what in enumerate(("obs", "reward", "done", "actions")):
            assert np.array_equal(tr[j], tw[j]), "%s differs in the launch of %d steps" % (what, k)
        for t in range(k):
            ret = (ret + tw[1][t]).astype(F32); length += 1
            d = tw[2][t] != 0
            want_ret += float(ret[d].astype(np.float64).sum()); want_len += int(length[d].sum()); episodes += int(d.sum())
            ret[d] = 0; length[d] = 0
        assert np.array_equal(ep_ret, ret)
        assert np.isclose(sums[0], want_ret, rtol=1e-12, atol=1e-9) and sums[1] == want_len
    assert episodes > 0
    if kind == O.PENDULUM:
        assert set(np.unique(tw[2])) == {0, 2}        # Pendulum never terminates by itself


STEP_CASES = [("CartPole-v1", O.CARTPOLE, 333, 120), ("Pendulum-v1", O.PENDULUM, 130, 210), ("MountainCar-v0", O.MOUNTAINCAR, 260, 210),
              ("Acrobot-v1 limit 25", O.ACROBOT, 129, 60), ("LunarLander-v2", O.LUNARLANDER, 70, 150), ("LunarLanderContinuous-v2", O.LUNARLANDER_CONT, 40, 120)]


@pytest.mark.parametrize("name,kind,n,k", STEP_CASES, ids=[c[0] for c in STEP_CASES])
def test_step_kernel_compaction_statistics_and_packed_done(hs, name, kind, n, k):
    """step_kernel as gymcuda_step_device launches it, ragged last CTA included: outputs against the oracle, the compacted
    list of finished envs (ballot + block scan + one atomic per CTA) against nonzero(done), the double-buffered counter,
    the statistics, done bytes through the packed 4-byte stores and (misaligned buffer) the per-lane stores, and for
    LunarLander the thread -> env permutation of the contact partition."""
    o, sim = pair(hs, kind, n, done_bits=True, time_limit=25 if kind == O.ACROBOT else 0)
    if sim.limit == 0:
        sim.limit = 0x7fffffff
    ep_ret = np.zeros(n, F32); sums = np.zeros(2, np.float64)
    ret = np.zeros(n, F32); length = np.zeros(n, np.int64); want_ret = 0.0; want_len = 0; episodes = 0
    for t in range(k):
        a = o.sample_actions()
        assert np.array_equal(sim.sample_kernel(), a)
        oo, orr, od = o.step(a)
        perm = sim.partition()[0] if (sim.lunar and t % 2 == 0) else None
        so, sr, sd, idx, flag = sim.step_kernel(a, perm=perm, ep_ret=ep_ret, sums=sums, done_bits=1, done_offset=(t % 3 == 2) * 1)
        assert flag == 0
        assert np.array_equal(sd, od), "done differs at step %d" % t
        assert np.array_equal(sr, orr), "reward differs at step %d" % t
        assert np.array_equal(so, oo), "observation differs at step %d" % t
        assert np.array_equal(np.sort(idx), np.nonzero(od)[0]), "compacted done list differs at step %d" % t
        ret = (ret + orr).astype(F32); length += 1
        d = od != 0
        want_ret += float(ret[d].astype(np.float64).sum()); want_len += int(length[d].sum()); episodes += int(d.sum())
        ret[d] = 0; length[d] = 0
        # finished episodes: stats[0] + the count of the latest launch, which the NEXT launch folds in (kernels.cuh)
        assert int(sim.stats[0]) + int(sim.done_count[(sim.seq - 1) & 1]) == episodes and int(sim.stats[1]) == 0
        assert np.array_equal(ep_ret, ret)
        assert np.isclose(sums[0], want_ret, rtol=1e-12, atol=1e-9) and sums[1] == want_len
    assert same_state(o, sim)
    assert episodes > 0


def _pid(s):
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


@pytest.mark.parametrize("kind,n,k,time_limit", [(O.LUNARLANDER, 47, 420, 0), (O.LUNARLANDER_CONT, 13, 150, 0)], ids=["LunarLander-v2", "LunarLanderContinuous-v2"])
def test_lunar_step_kernel_three_lanes_per_lander(hs, request, kind, n, k, time_limit):
    """step_kernel<LunarLanderT<C, true, TRIO>>: ten landers per warp, three lanes each (one body per lane inside the velocity /
    position iterations, joints and exits exchanged with trio-masked shuffles and votes), lanes 30-31 idle, ragged last warp --
    bit-identical to the oracle through free flight, touch-down on one and two legs, belly contact, rest, sleep and the fused
    auto-reset (PID heuristic mixed with random actions so that every one of them occurs), with and without the partition's
    thread -> env permutation."""
    full = "ascending" in request.node.callspec.id     # the long run under one schedule, a short one (free flight) under the others
    if not full:
        if kind != O.LUNARLANDER:
            pytest.skip("the continuous variant runs under one schedule")
        k = 12
    o, sim = pair(hs, kind, n, done_bits=True, time_limit=time_limit)
    if sim.limit == 0:
        sim.limit = 0x7fffffff
    rng = np.random.default_rng(3)
    obs = sim.reset_kernel(mask=np.zeros(n, np.uint8))           # observe only
    hs.hostsim_set_trio(1)
    try:
        episodes = 0; legs = 0; asleep = 0
        for t in range(k):
            if kind == O.LUNARLANDER:
                a = np.where(rng.random(n) < 0.8, _pid(obs), rng.integers(0, 4, n)).astype(np.int32)
            else:
                a = rng.uniform(-1, 1, (n, 2)).astype(F32)
            oo, orr, od = o.step(a)
            # the long run steps two windows with the TRIO kernel -- the first touch-downs and crashes (fused auto-resets) and, 300
            # steps later, landers at rest up to the one that falls asleep -- and the rest with the per-thread body of the plain
            # kernel (fast), in lockstep with the oracle
            trio_now = (75 <= t < 125 or t >= 380) if kind == O.LUNARLANDER else t >= 100
            if full and not trio_now:
                obs = sim.step(a)[0]
                assert np.array_equal(obs, oo)
                continue
            perm = sim.partition()[0] if t % 2 == 0 else None
            so, sr, sd, idx, flag = sim.step_kernel(a, perm=perm, done_bits=1)
            assert flag == 0
            assert np.array_equal(sd, od), "done differs at step %d" % t
            assert np.array_equal(sr, orr), "reward differs at step %d" % t
            assert np.array_equal(so, oo), "observation differs at step %d" % t
            assert np.array_equal(np.sort(idx), np.nonzero(od)[0]), "compacted done list differs at step %d" % t
            assert same_state(o, sim), "state differs at step %d" % t
            episodes += int((od != 0).sum()); legs += int((oo[:, 6:] > 0).any(1).sum()); asleep += int((orr == 100.0).sum())
            obs = so
        if full:
            assert episodes > 0 and legs > 0
            if kind == O.LUNARLANDER:
                assert asleep > 0          # a lander came to rest and fell asleep (+100)
    finally:
        hs.hostsim_set_trio(0)


def test_step_kernel_rejects_invalid_actions_and_broadcasts(hs):
    n = 200
    o, sim = pair(hs, O.MOUNTAINCAR, n)
    before = sim.abi_state()
    a = o.sample_actions()
    bad = np.array([3, 64, 65, 199])
    a[bad] = [3, -1, 7, 1 << 20]
    obs, rew, done, idx, flag = sim.step_kernel(a)
    assert flag == 1 and int(sim.stats[1]) == len(bad)           # mapped host flag raised, invalid actions counted per warp
    after = sim.abi_state()
    assert np.array_equal(after[bad], before[bad])               # those envs were not stepped
    ok = np.setdiff1d(np.arange(n), bad)
    assert not np.array_equal(after[ok], before[ok]) and not done.any() and (rew[bad] == 0).all()
    # IVecEnv.Step(int action): one action for every env (VecEnvWrapper.cs:22-24)
    o2, sim2 = pair(hs, O.CARTPOLE, 150)
    for t in range(40):
        oo, orr, od = o2.step(np.full(150, t % 2, np.int32))       # the reference test's i % 2 loop (CartpoleEnvironment.cs:19-26)
        so, sr, sd, idx, flag = sim2.step_kernel(None, bcast=t % 2)
        assert np.array_equal(so, oo) and np.array_equal(sr, orr) and np.array_equal(sd, od)
        assert np.array_equal(np.sort(idx), np.nonzero(od)[0])


@pytest.mark.parametrize("n", [1, 31, 255, 256, 257, 5000, 270000])
def test_contact_partition_is_a_stable_partition(hs, n):
    """partition_count / partition_scan / partition_scatter (gymcuda.cu: lunar_step): free-flight landers first, landers
    whose broad phase holds a contact pair after them, each class in ascending env order; 270 000 envs make the single-CTA
    scan walk more than one 1024-wide tile of block counts."""
    rng = np.random.default_rng(n)
    sim = HostSim(hs, O.LUNARLANDER, n, 1)
    sim.aux[:] = 0
    touching = rng.random(n) < (0.3 if n < 100000 else 0.02)
    pairs_word = 26                                   # lunar::Lander::pairs[0]: 0xffffffff = no contact exists
    sim.aux[pairs_word, :] = -1
    sim.aux[pairs_word, np.nonzero(touching)[0]] = rng.integers(0, 0x2a, int(touching.sum())) | ~np.int32(0xff)
    perm, block_free = sim.partition()
    want = np.concatenate([np.nonzero(~touching)[0], np.nonzero(touching)[0]]).astype(np.int32)
    assert np.array_equal(perm, want)
    assert block_free[-1] == int((~touching).sum())


@pytest.mark.parametrize("kind", [O.CARTPOLE, O.MOUNTAINCAR, O.ACROBOT, O.LUNARLANDER, O.PENDULUM, O.MOUNTAINCAR_CONT, O.LUNARLANDER_CONT])
def test_sample_and_masked_reset_kernels(hs, kind):
    n = 150
    o, sim = pair(hs, kind, n)
    rng = np.random.default_rng(kind)
    for t in range(6):
        a = o.sample_actions()
        assert np.array_equal(sim.sample_kernel(), a)
        if sim.actn > 0:        # Discrete.Sample(mask): uniform over the entries equal to 1, Start when there is none (Discrete.cs:19-25)
            mask = (rng.random((n, sim.actn)) < 0.5).astype(np.uint8)
            mask[:5] = 0; mask[5:10] = 1; mask[10:15, 1:] = 2
            assert np.array_equal(sim.sample_kernel(mask), o.sample_actions(mask))
        oo, orr, od = o.step(a)
        so, sr, sd, idx, flag = sim.step_kernel(a)
        assert np.array_equal(so, oo) and np.array_equal(sd, od)
        m = (rng.random(n) < 0.3).astype(np.uint8)
        assert np.array_equal(sim.reset_kernel(m), o.reset(m))          # reset_masked: obs of every env, new episode for the masked ones
        assert same_state(o, sim)


def test_fused_step_gather_and_wait_kernels(hs):
    """gymcuda_step_gather_device on the host: four "ranks" (shards of one batch) run step_kernel one after the other with
    every rank's gather buffer standing in for the cudaIpc mapping; each rank's buffer must end up holding the
    observations of the whole batch in the slot of this step's parity, its arrival flags must carry the gather sequence
    number (published by the LAST CTA, which also re-arms the block counter), and gather_wait_kernel must accept exactly
    that -- and name the missing rank when one never publishes."""
    import ctypes as C
    world, n, kind = 4, 300, O.ACROBOT
    od = O.dims(kind)["obs_dim"]
    full = O.OracleEnv(kind, world * n, seed=9, env_id_offset=0, auto_reset=True, mode=O.MODE_F32)
    full.reset()
    sims = [HostSim(hs, kind, n, 9, r * n) for r in range(world)]
    for s in sims:
        s.reset_kernel()
    bufs = [np.full((2, world, n, od), np.nan, F32) for _ in range(world)]
    flags = [np.zeros(world, np.uint32) for _ in range(world)]
    counters = [np.zeros(1, np.uint32) for _ in range(world)]
    PF = (C.c_void_p * world)(*[b.ctypes.data for b in bufs])
    PG = (C.c_void_p * world)(*[f.ctypes.data for f in flags])
    for gseq in range(1, 6):
        a = full.sample_actions()
        oo, orr, od_ = full.step(a)
        for r, s in enumerate(sims):
            if gseq == 5 and r == 2:
                continue                                  # rank 2 "dies" before its fifth step
            hs.hostsim_set_gather(world, r, gseq, PF, PG, counters[r].ctypes.data_as(C.c_void_p))
            so, sr, sd, idx, flag = s.step_kernel(a[r * n:(r + 1) * n])
            assert np.array_equal(sr, orr[r * n:(r + 1) * n]) and np.array_equal(sd, od_[r * n:(r + 1) * n])
            assert counters[r][0] == 0                    # re-armed by the last CTA
        for r in range(world):
            if gseq < 5:
                assert hs.hostsim_gather_wait(flags[r].ctypes.data_as(C.c_void_p), world, gseq) == 0
                assert np.array_equal(bufs[r][gseq & 1].reshape(world * n, od), oo), "rank %d, gather %d" % (r, gseq)
                assert (flags[r] == gseq).all()
            else:
                assert hs.hostsim_gather_wait(flags[r].ctypes.data_as(C.c_void_p), world, gseq) == 1 + 2


@pytest.mark.parametrize("kind", [O.CARTPOLE, O.ACROBOT, O.LUNARLANDER], ids=["CartPole-v1", "Acrobot-v1", "LunarLander-v2"])
def test_per_env_seeds(hs, kind):
    """VecEnv.Seed(int[]) (VecEnv.cs:48-53): every env draws from the stream of ITS seed -- reset, random policy,
    LunarLander's constructor and per-step dispersion draws -- in every kernel."""
    n = 130
    o, sim = pair(hs, kind, n)
    seeds = np.random.default_rng(3).integers(-2**31, 2**31, n).astype(np.int32)
    seeds[:4] = [0, 5, 5, -1]
    o.seed_each(seeds)
    sim.seed_each(seeds)
    try:
        obs = sim.reset_kernel()
        assert np.array_equal(obs, o.reset())
        assert not np.array_equal(obs[1], obs[2])                    # same seed, different env id: different streams
        for t in range(12):
            a = o.sample_actions()
            assert np.array_equal(sim.sample_kernel(), a)
            oo, orr, od = o.step(a)
            so, sr, sd, idx, flag = sim.step_kernel(a)
            assert np.array_equal(so, oo) and np.array_equal(sr, orr) and np.array_equal(sd, od)
        tr = sim.rollout(70, True, 64)
        tw = o.rollout_random(70)
        for j, what in enumerate(("obs", "reward", "done", "actions")):
            assert np.array_equal(tr[j], tw[j]), what
        assert same_state(o, sim)
    finally:
        sim.unseed()


@pytest.mark.parametrize("kind,n", [(O.CARTPOLE, 150), (O.MOUNTAINCAR, 97), (O.PENDULUM, 64)], ids=["CartPole-v1", "MountainCar-v0", "Pendulum-v1"])
def test_reference_semantics_without_auto_reset(hs, kind, n):
    """Auto-reset off = the reference's semantics: a finished env keeps stepping until the caller resets it; CartPole
    then counts steps_beyond_done and pays 1, 0, 0, ... (CartPoleEnv.cs:168-183), and `done` stays up."""
    o = O.OracleEnv(kind, n, seed=8, env_id_offset=3, auto_reset=False, mode=O.MODE_F32)
    sim = HostSim(hs, kind, n, 8, 3, auto_reset=False)
    assert np.array_equal(sim.reset_kernel(), o.reset())
    for t in range(60):
        a = o.sample_actions()
        oo, orr, od = o.step(a)
        so, sr, sd, idx, flag = sim.step_kernel(a)
        assert np.array_equal(so, oo) and np.array_equal(sr, orr) and np.array_equal(sd, od), "step %d" % t
        assert np.array_equal(np.sort(idx), np.nonzero(od)[0])
    for k, all_out in ((21, True), (40, False), (160, True)):
        tr = sim.rollout(k, all_out, 64)
        tw = o.rollout_random(k)
        for j, what in enumerate(("obs", "reward", "done", "actions")):
            assert np.array_equal(tr[j], tw[j]), "%s differs in the launch of %d steps" % (what, k)
    st, aux, ot = o.get_state()
    assert ot == sim.t and np.array_equal(sim.abi_state(), st.astype(F32))
    if kind == O.CARTPOLE:
        assert np.array_equal(sim.sbd, aux[:, 0]) and (sim.sbd > 0).any()      # poles fell and kept being stepped
        assert (tw[1][-1][sim.sbd > 0] == 0).all()                             # ... for a reward of 0
    # the caller's reset of the finished envs (README.md:36-40), as a masked reset
    mask = (tw[2][-1] != 0).astype(np.uint8)
    assert np.array_equal(sim.reset_kernel(mask), o.reset(mask))
    assert same_state(o, sim)


@pytest.mark.parametrize("n,od", [(1000, 4), (257, 8), (64, 3), (31, 2)])
def test_normalize_kernels_against_numpy(hs, n, od):
    """normalize.cuh (gymcuda_normalize_device) on the host: running mean / variance of the observation components and of
    the discounted return over several batches (warp shuffles + shared-memory rows + one atomic per value and CTA), the
    in-place normalisation with clipping, frozen statistics, and the return reset at `done` -- against numpy float64."""
    from hostsim_lib import NormalizeModel, _p
    rng = np.random.default_rng(n)
    model = NormalizeModel(n, od, gamma=0.97, eps=1e-6, clip_obs=1.5, clip_reward=4.0)
    acc = np.zeros(20, np.float64); ret = np.zeros(n, np.float32)
    scale = rng.uniform(0.1, 30.0, od); shift = rng.uniform(-5, 5, od)
    for it in range(6):
        obs = (rng.standard_normal((n, od)) * scale + shift).astype(F32)
        rew = rng.uniform(-2, 3, n).astype(F32)
        done = (rng.random(n) < 0.2).astype(np.uint8)
        update = it != 4                                         # one evaluation-mode call: statistics frozen
        want_o, want_r = model(obs, rew, done, update)
        got_o, got_r = obs.copy(), rew.copy()
        hs.hostsim_normalize(_p(got_o), _p(got_r), _p(done), _p(ret), _p(acc), n, od, 0.97, 1e-6, 1.5, 4.0, int(update))
        assert np.allclose(got_o, want_o, rtol=1e-6, atol=1e-6) and np.allclose(got_r, want_r, rtol=1e-6, atol=1e-6)
        assert np.array_equal(ret, model.ret)
        assert acc[18] == model.count and acc[19] == model.count_ret and np.allclose(acc[:od], model.s, rtol=1e-12) and np.allclose(acc[8:8 + od], model.q, rtol=1e-12)
        assert (np.abs(got_o) <= 1.5).all() and (np.abs(got_o) == 1.5).any()         # the clip is active
    # obs only / reward only
    obs = rng.standard_normal((n, od)).astype(F32); got = obs.copy()
    want, _ = model(obs, None, None, True)
    hs.hostsim_normalize(_p(got), None, None, _p(ret), _p(acc), n, od, 0.97, 1e-6, 1.5, 4.0, 1)
    assert np.allclose(got, want, rtol=1e-6, atol=1e-6)
    assert acc[18] == model.count and acc[19] == model.count_ret and acc[18] == acc[19] + n   # the obs-only call did not touch the return count
    rew = rng.uniform(-2, 3, n).astype(F32); got_r = rew.copy()
    _, want_r = model(None, rew, None, True)
    hs.hostsim_normalize(None, _p(got_r), None, _p(ret), _p(acc), n, od, 0.97, 1e-6, 1.5, 4.0, 1)
    assert np.allclose(got_r, want_r, rtol=1e-6, atol=1e-6) and acc[18] == acc[19]


def terminal_checker(kind, n, seed, off, time_limit):
    """Two oracles: A runs with auto-reset (what the kernel does); before every step its state is copied into B, which runs
    WITHOUT auto-reset -- B's observation after the step is the terminal observation of every env whose step ended an episode."""
    a = O.OracleEnv(kind, n, seed=seed, env_id_offset=off, auto_reset=True, mode=O.MODE_F32, time_limit=time_limit)
    b = O.OracleEnv(kind, n, seed=seed, env_id_offset=off, auto_reset=False, mode=O.MODE_F32, time_limit=time_limit)
    return a, b


@pytest.mark.parametrize("name,kind,n,limit,k", [("CartPole-v1", O.CARTPOLE, 300, 0, 60), ("MountainCar-v0", O.MOUNTAINCAR, 70, 25, 60),
                                                   ("Acrobot-v1", O.ACROBOT, 90, 20, 45), ("LunarLander-v2", O.LUNARLANDER, 40, 90, 200)],
                         ids=["CartPole-v1", "MountainCar-v0", "Acrobot-v1", "LunarLander-v2"])
def test_step_kernel_terminal_observations(hs, name, kind, n, limit, k):
    """gymcuda_set_terminal_obs: rows of the side buffer of the envs whose step returned done == the observation of the state
    the episode ended in (terminated or truncated by the time limit); the other rows are not touched; the step's own outputs
    and the state are what they are without the side buffer (LunarLander: the fused crash + zero-step solve is bypassed)."""
    rng = np.random.default_rng(17)
    a, b = terminal_checker(kind, n, 5, 77, limit)
    sim = HostSim(hs, kind, n, 5, 77)
    if limit:
        sim.limit = limit
    assert np.array_equal(sim.reset_kernel(), a.reset())
    b.reset()
    term = np.full((n, sim.od), -7.0, F32)
    seen = 0
    for _ in range(k):
        act = rng.integers(0, sim.actn, n).astype(np.int32)
        st, aux, t = a.get_state()
        b.set_state(st, aux, t)
        want_o, want_r, want_d = a.step(act)
        term_o, _, term_d = b.step(act)
        before = term.copy()
        got_o, got_r, got_d, _, _ = sim.step_kernel(act, terminal_obs=term)
        assert np.array_equal(got_o, want_o) and np.array_equal(got_r, want_r) and np.array_equal(got_d, want_d)
        d = want_d != 0
        assert np.array_equal(term[d], term_o[d]) and np.array_equal(term[~d], before[~d])
        seen += int(d.sum())
        assert same_state(a, sim)
    assert seen > 0


@pytest.mark.parametrize("name,kind,n", [("CartPole-v1", O.CARTPOLE, 100), ("Pendulum-v1", O.PENDULUM, 72), ("MountainCar-v0", O.MOUNTAINCAR, 64),
                                         ("MountainCarContinuous-v0", O.MOUNTAINCAR_CONT, 33), ("Acrobot-v1", O.ACROBOT, 50)],
                         ids=["CartPole-v1", "Pendulum-v1", "MountainCar-v0", "MountainCarContinuous-v0", "Acrobot-v1"])
@pytest.mark.parametrize("shape", ["generic", "chunked64", "chunked512"])
def test_rollout_kernel_with_caller_supplied_actions(hs, name, kind, n, shape):
    """gymcuda_step_many*: the rollout kernel fed from `actions_in` -- the generic variant, and the chunked all-outputs SUPPLIED
    variant in its 64- and 512-thread shapes (started at an unaligned step index: head, chunks, tail) -- == k oracle steps with
    those actions (auto-reset, default time limits); an invalid action leaves its env unstepped for that step, is counted and
    raises the host flag."""
    if shape == "chunked512" and kind == O.ACROBOT:
        pytest.skip("Acrobot has no 512-thread shape (ROLLOUT_CHUNK off)")
    rng = np.random.default_rng(23)
    o, sim = pair(hs, kind, n)
    if shape != "generic":           # 3 warm-up steps: the chunked launch then starts at t = 3 (a 5-step head before the first chunk)
        for _ in range(3):
            a0 = np.ones(n, np.int32) if sim.actn > 0 else np.zeros((n, sim.ad), F32)
            o.step(a0); sim.step(a0)
    k = 37
    if sim.actn > 0:
        acts = rng.integers(0, sim.actn, (k, n)).astype(np.int32)
    else:
        acts = rng.uniform(-2.5, 2.5, (k, n, sim.ad)).astype(F32)
    bad = np.zeros((k, n), bool)
    if kind != O.CARTPOLE:           # CartPole accepts anything (Debug.Assert only, CartPoleEnv.cs:139)
        bad = rng.random((k, n)) < 0.03
        if sim.actn > 0:
            acts[bad] = 7
        else:
            acts[bad] = np.nan
    obs, rew, done, _, episodes = sim.rollout(k, all_out=shape != "generic", block=512 if shape == "chunked512" else 64, actions_in=acts)
    want_eps = 0
    for j in range(k):
        wo, wr, wd = o.step(acts[j])
        assert o.invalid == int(bad[j].sum())
        ok = ~bad[j]
        assert np.array_equal(obs[j][ok], wo[ok]) and np.array_equal(rew[j][ok], wr[ok]) and np.array_equal(done[j][ok], wd[ok])
        assert np.array_equal(obs[j][bad[j]], wo[bad[j]])              # unchanged observation of the unstepped env
        assert not rew[j][bad[j]].any() and not done[j][bad[j]].any()
        want_eps += int(wd.sum())
    assert episodes == want_eps
    assert sim.rollout_invalid == (int(bad.any()), int(bad.sum()))
    assert same_state(o, sim)
