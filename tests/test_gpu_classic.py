"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU oracle.

T1  teacher-forced single step from random + adversarial float32 states:
      vs ORACLE F64 (the reference's arithmetic): done bit-exact, reward exact/1e-5, state <= 1e-5
      vs ORACLE F32 (engine arithmetic twin):     everything bit-exact
T2  free-running rollouts with the in-kernel random policy + auto-reset vs the F32 twin, bit-exact,
    and vs the F64 oracle teacher-forced along the trajectory.
T3  the reference test's loop shape (i % 2 actions, reset on done; CartpoleEnvironment.cs:19-30).
T4  invariances: sharding by env_id_offset, step-vs-rollout, k split.
"""
import numpy as np
import pytest

import oracle_lib as O
import gymnet_b200 as G
from helpers import KINDS, RTOL, STATE_SCALE, random_actions, random_states, rel_err

pytestmark = pytest.mark.gpu

CLASSIC = ["CartPole-v1", "Pendulum-v1", "MountainCar-v0", "MountainCarContinuous-v0", "Acrobot-v1"]


def adversarial_states(name, rng, n):
    """States whose successor lands within a few float32 ulps of a termination threshold."""
    s = random_states(name, rng, n)
    if name == "CartPole-v1":
        tau = np.float64(np.float32(0.02))
        xthr, tthr = np.float64(np.float32(2.4)), np.float64(np.float32(12 * 2 * np.pi / 360))
        k = n // 2
        sign = rng.choice([-1.0, 1.0], size=k)
        # x + tau*x_dot == +-x_thr up to a few ulps
        s[:k, 0] = (sign * xthr - tau * s[:k, 1].astype(np.float64)).astype(np.float32)
        jitter = rng.integers(-3, 4, size=k)
        for j in range(k):
            v = s[j, 0]
            for _ in range(abs(int(jitter[j]))):
                v = np.nextafter(v, np.float32(np.inf if jitter[j] > 0 else -np.inf), dtype=np.float32)
            s[j, 0] = v
        sign = rng.choice([-1.0, 1.0], size=n - k)
        s[k:, 2] = (sign * tthr - tau * s[k:, 3].astype(np.float64)).astype(np.float32)
    elif name in ("MountainCar-v0", "MountainCarContinuous-v0"):
        goal = 0.5 if name == "MountainCar-v0" else 0.45
        s[:, 1] = rng.uniform(0.0, 0.05, size=n).astype(np.float32)
        s[:, 0] = (goal - s[:, 1] + rng.uniform(-2e-3, 2e-3, size=n)).astype(np.float32)
    elif name == "Acrobot-v1":
        # near the swing-up height: -cos(t1) - cos(t1+t2) ~ 1
        t1 = rng.uniform(2.0, 2.2, size=n)
        s[:, 0] = t1; s[:, 1] = rng.uniform(-0.3, 0.3, size=n)
        s[:, 2:] = rng.uniform(-0.5, 0.5, size=(n, 2))
        s = s.astype(np.float32)
    return s


def state_err(name, a, b):
    """Relative state error; Acrobot's wrapped angles are compared modulo 2*pi (a value next to +-pi may
    legitimately land on either side of the wrap)."""
    a = np.asarray(a, np.float64).copy(); b = np.asarray(b, np.float64)
    if name == "Acrobot-v1":
        d = a[:, :2] - b[:, :2]
        a[:, :2] = b[:, :2] + (d + np.pi) % (2 * np.pi) - np.pi
    return rel_err(a, b, STATE_SCALE[name])


def obs_atol(name, state):
    """cos/sin of a float32 angle inherit the angle's own rounding: 1e-5 relative on |theta|."""
    if name == "Pendulum-v1":
        return (RTOL * np.maximum(1.0, np.abs(np.asarray(state, np.float64)[:, :1])))
    return RTOL * np.maximum(1.0, np.abs(STATE_SCALE[name]).max() if name == "Acrobot-v1" else 1.0)


def reward_atol(name, state_before, reward):
    """Pendulum's cost uses angle_normalize(th) of an UNWRAPPED float32 angle (upstream keeps th
    unbounded): the angle's own float32 spacing bounds the cost error by 2*pi*ulp32(|th|+pi)."""
    tol = RTOL * np.maximum(1.0, np.abs(reward.astype(np.float64)))
    if name == "Pendulum-v1":
        th = np.abs(np.asarray(state_before, np.float64)[:, 0]) + np.pi
        tol = tol + 2 * np.pi * np.spacing(th.astype(np.float32)).astype(np.float64)
    return tol


def teacher_forced(name, states, actions, sbd=None):
    n = len(states)
    kind = KINDS[name]
    aux = np.zeros((n, 3), np.int32)
    aux[:, 0] = -1 if sbd is None else sbd
    env = G.make(name, n, seed=3, auto_reset=False, time_limit=-1)
    env.ResetBatch()
    env.SetState(states, aux, 5)
    obs, rew, done = env.StepBatch(actions)
    st, ax, t = env.GetState()
    assert t == 6
    out = {"gpu": (obs, rew, done, st, ax)}
    for mode in (O.MODE_F64_F32STORE, O.MODE_F32):
        o = O.OracleEnv(kind, n, seed=3, auto_reset=False, time_limit=-1, mode=mode)
        o.reset()
        o.set_state(states.astype(np.float64), aux, 5)
        oo, orr, od = o.step(actions)
        ost, oax, _ = o.get_state()
        out[mode] = (oo, orr, od, ost, oax)
    env.Close()
    return out


@pytest.mark.parametrize("name", CLASSIC)
@pytest.mark.parametrize("adversarial", [False, True])
def test_t1_teacher_forced_single_step(name, adversarial):
    rng = np.random.default_rng(11 + adversarial)
    n = 1 << 18
    states = adversarial_states(name, rng, n) if adversarial else random_states(name, rng, n)
    env = G.make(name, 1)
    actions = random_actions(env, rng, n)
    env.Close()
    sbd = rng.integers(-1, 3, size=n).astype(np.int32) if name == "CartPole-v1" else None
    r = teacher_forced(name, states, actions, sbd)
    obs, rew, done, st, ax = r["gpu"]
    # --- vs the reference's arithmetic (F64, float32 store)
    oo, orr, od, ost, oax = r[O.MODE_F64_F32STORE]
    assert np.array_equal(done, od), "%d done flags differ from the F64 oracle" % int((done != od).sum())
    assert np.array_equal(ax, oax)
    assert state_err(name, st, ost).max() <= RTOL
    assert (np.abs(obs.astype(np.float64) - oo) <= obs_atol(name, ost)).all()
    assert (np.abs(rew.astype(np.float64) - orr) <= reward_atol(name, states, orr)).all()
    if adversarial and name != "Pendulum-v1":
        assert 0 < done.sum() < n   # both outcomes are exercised next to the threshold
    # --- vs the engine-arithmetic twin: bit for bit
    oo, orr, od, ost, oax = r[O.MODE_F32]
    assert np.array_equal(done, od) and np.array_equal(ax, oax)
    exact = np.ones(n, bool)
    if name == "Acrobot-v1":
        # beyond (9, 18) rad/s the engine steps in DOUBLE precision (engine arithmetic v3) through CUDA's sin / cos, the
        # twin through libm's: both within an ulp of double, so the float32-rounded results agree except where a 1e-16
        # difference straddles a float32 rounding boundary -- checked to 1e-6 there instead of bit for bit
        exact = (np.abs(states[:, 2]) <= 9.0) & (np.abs(states[:, 3]) <= 18.0)
        if not adversarial:
            assert 0.1 < exact.mean() < 0.9
            assert state_err(name, st[~exact], ost[~exact]).max() <= 1e-6
            assert (st[~exact] == ost[~exact].astype(np.float32)).all(axis=1).mean() > 0.999
            assert np.abs(obs[~exact].astype(np.float64) - oo[~exact]).max() <= 1e-5
    assert np.array_equal(st[exact], ost[exact].astype(np.float32))
    assert np.array_equal(obs[exact], oo[exact])
    assert np.array_equal(rew, orr)


def test_cartpole_done_on_exact_float32_threshold_ties():
    """The kernel decides `done` from the float32 position fmaf(tau, x_dot, x) and recomputes the reference's
    double-precision test (CartPoleEnv.cs:154,156,167) only when that float32 value EQUALS a threshold: a
    velocity too small to move the float32 sum still makes the reference's double sum exceed (or not) the
    threshold.  Every sign combination, against numpy float64 and both oracle modes."""
    xthr, tthr = np.float32(2.4), np.float32(12 * 2 * np.pi / 360)
    tau = np.float64(np.float32(0.02))
    rows = []
    for comp, thr in ((0, xthr), (2, tthr)):
        for side in (1.0, -1.0):
            for vel in (0.0, 1e-12, -1e-12, 1e-9, -1e-9, 3e-8, -3e-8, 1e-6, -1e-6):
                st = np.zeros(4, np.float32)
                st[comp] = np.float32(side) * thr
                st[comp + 1] = np.float32(vel)
                rows.append(st)
    states = np.stack(rows)
    n = len(states)
    actions = (np.arange(n) % 2).astype(np.int32)
    s64 = states.astype(np.float64)
    nx = s64[:, 0] + tau * s64[:, 1]
    nth = s64[:, 2] + tau * s64[:, 3]
    expect = ((np.abs(nx) > np.float64(xthr)) | (np.abs(nth) > np.float64(tthr))).astype(np.uint8)
    assert 0 < expect.sum() < n
    # the float32 sums of the tiny-velocity rows really are ties
    f32sum = (s64[:, 2] + tau * s64[:, 3]).astype(np.float32)
    assert (np.abs(f32sum[n // 2:]) == tthr).sum() >= 10
    r = teacher_forced("CartPole-v1", states, actions)
    assert np.array_equal(r["gpu"][2], expect)
    assert np.array_equal(r[O.MODE_F64_F32STORE][2], expect)
    assert np.array_equal(r[O.MODE_F32][2], expect)
    assert np.array_equal(r["gpu"][3], r[O.MODE_F32][3].astype(np.float32))


@pytest.mark.parametrize("name", CLASSIC)
def test_t2_free_running_rollout_bit_exact_vs_twin(name):
    n, k = 4096, 600
    env = G.make(name, n, seed=1234, auto_reset=True, env_id_offset=77)
    o = O.OracleEnv(KINDS[name], n, seed=1234, auto_reset=True, env_id_offset=77, mode=O.MODE_F32)
    assert np.array_equal(env.ResetBatch(), o.reset())
    obs, rew, done, act = env.RolloutRandom(k)
    oo, orr, od, oa = o.rollout_random(k)
    assert np.array_equal(act, oa)
    assert np.array_equal(done, od)
    assert np.array_equal(rew, orr)
    assert np.array_equal(obs, oo)
    st, ax, t = env.GetState()
    ost, oax, ot = o.get_state()
    assert t == ot == k
    assert np.array_equal(st, ost.astype(np.float32))
    assert done.sum() > 0 or name == "MountainCarContinuous-v0"
    assert env.Stats()["episodes"] == int(done.sum())
    env.Close()


@pytest.mark.parametrize("name", CLASSIC)
def test_t2_trajectory_teacher_forced_vs_f64(name):
    """Along a GPU trajectory, every transition agrees with the reference arithmetic: done bit-exact."""
    n, k = 2048, 300
    env = G.make(name, n, seed=5, auto_reset=True)
    env.ResetBatch()
    o = O.OracleEnv(KINDS[name], n, seed=5, auto_reset=True, mode=O.MODE_F64_F32STORE)
    o.reset()
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(k):
        st, ax, t = env.GetState()
        o.set_state(st.astype(np.float64), ax, t)
        a = random_actions(env, rng, n)
        obs, rew, done = env.StepBatch(a)
        oo, orr, od = o.step(a)
        assert np.array_equal(done, od)
        nd = done == 0   # post-reset states are identical draws; compare the rest numerically
        gs, _, _ = env.GetState()
        os_, _, _ = o.get_state()
        worst = max(worst, float(state_err(name, gs[nd], os_[nd]).max(initial=0.0)))
        assert np.array_equal(gs[~nd], os_[~nd].astype(np.float32))      # same reset draws
        assert np.abs(obs[~nd] - oo[~nd]).max(initial=0.0) <= RTOL        # obs of them: detmath vs libm cos/sin
        assert (np.abs(rew.astype(np.float64) - orr) <= reward_atol(name, st, orr)).all()
    assert worst <= RTOL
    env.Close()


def test_t3_reference_test_loop_shape():
    """tests/Gym.Tests/Envs/Classic/CartpoleEnvironment.cs:19-30: 1000 iterations, Reset on done else Step(i % 2)."""
    n = 64
    env = G.CartPoleVecEnv(n, seed=0)
    o = O.OracleEnv(O.CARTPOLE, n, seed=0, mode=O.MODE_F32)
    done = np.ones(n, bool)
    for i in range(1000):
        if done.any():
            m = None if done.all() else done.astype(np.uint8)   # first pass: every env is reset
            assert np.array_equal(env.ResetBatch(mask=m), o.reset(mask=m))
            done[:] = False
        else:
            a = np.full(n, i % 2, np.int32)
            obs, rew, d = env.StepBatch(a)
            oo, orr, od = o.step(a)
            assert np.array_equal(obs, oo) and np.array_equal(rew, orr) and np.array_equal(d, od)
            done = d.astype(bool)
    env.Close()


def test_ivecenv_surface_and_broadcast_step():
    env = G.CartPoleVecEnv(8, seed=9)
    obs = env.Reset()
    assert len(obs) == 8 and obs[0].shape == (4,) and obs[0].dtype == np.float32
    assert all(abs(v) <= 0.05 for o in obs for v in o)          # CartPoleEnv.cs:65
    steps = env.Step(1)                                           # IVecEnv.Step(int action): broadcast
    assert len(steps) == 8
    ob, reward, done, info = steps[0]                             # Step.Deconstruct
    assert reward == 1.0 and done is False and info is None
    assert env.ActionSpace.N == 2 and env.ObservationSpace.Shape == (4,)
    np.testing.assert_array_equal(env.ObservationSpace.High,
                                  np.array([4.800000190734863, np.finfo(np.float32).max, 0.41887903213500977,
                                            np.finfo(np.float32).max], np.float32))
    env.Close()


def test_cartpole_steps_beyond_done_quirk():
    """CartPoleEnv.cs:168-183: reward 1 on the terminal step, then 0 with steps_beyond_done counting up."""
    env = G.CartPoleVecEnv(2, seed=0, auto_reset=False)
    env.ResetBatch()
    st = np.array([[2.39, 3.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0]], np.float32)
    env.SetState(st, np.array([[-1, 0, 1], [-1, 0, 1]], np.int32), 0)
    a = np.array([1, 1], np.int32)
    _, r1, d1 = env.StepBatch(a)
    _, r2, d2 = env.StepBatch(a)
    _, r3, d3 = env.StepBatch(a)
    assert list(d1) == [1, 0] and list(r1) == [1.0, 1.0]
    assert d2[0] == 1 and r2[0] == 0.0 and d3[0] == 1 and r3[0] == 0.0
    _, ax, _ = env.GetState()
    assert ax[0, 0] == 2 and ax[1, 0] == -1
    env.Close()


def test_step_before_reset_is_estate():
    env = G.CartPoleVecEnv(4)
    with pytest.raises(G.GymCudaError) as ei:
        env.StepBatch(np.zeros(4, np.int32))
    assert ei.value.status == -6
    env.Close()


def test_invalid_actions():
    # CartPole: Debug.Assert only (CartPoleEnv.cs:139): accepted, != 1 means "left"
    env = G.CartPoleVecEnv(4, seed=1); env.ResetBatch()
    o1, _, _ = env.StepBatch(np.array([7, 0, -3, 0], np.int32))
    st, _, _ = env.GetState()
    assert np.array_equal(st[0, 1] < 0, True) and st[0, 1] != 0
    env.Close()
    # the others reject: call returns EACTION (-> InvalidActionError), offending envs unstepped
    env = G.AcrobotVecEnv(4, seed=1); before = env.ResetBatch()
    with pytest.raises(G.InvalidActionError):
        env.StepBatch(np.array([0, 3, 1, -1], np.int32))
    after = env.Observe()
    assert np.array_equal(after[1], before[1]) and np.array_equal(after[3], before[3])
    assert not np.array_equal(after[0], before[0])
    assert env.Stats()["invalid_actions"] == 2
    env.Close()


@pytest.mark.parametrize("name", ["CartPole-v1", "Acrobot-v1"])
def test_t4_sharding_invariance(name):
    """Global-env-id keyed streams: 2 shards of n/2 == 1 batch of n (what 8 GPUs rely on)."""
    n, k = 2048, 200
    full = G.make(name, n, seed=42, auto_reset=True)
    full.ResetBatch()
    fo, fr, fd, fa = full.RolloutRandom(k)
    for part in range(2):
        sh = G.make(name, n // 2, seed=42, auto_reset=True, env_id_offset=part * n // 2)
        sh.ResetBatch()
        so, sr, sd, sa = sh.RolloutRandom(k)
        sl = slice(part * n // 2, (part + 1) * n // 2)
        assert np.array_equal(so, fo[:, sl]) and np.array_equal(sd, fd[:, sl]) and np.array_equal(sa, fa[:, sl])
        sh.Close()
    full.Close()


@pytest.mark.parametrize("name", CLASSIC)
def test_t4_rollout_equals_repeated_step_and_k_split(name):
    """One 130-step rollout == the same steps split over launches that start at unaligned step indices (the
    kernel's 8-step chunks are aligned to the absolute step index: heads and tails run the generic loop)
    == 130 single steps fed the recorded actions."""
    n = 1024 if name == "CartPole-v1" else 256
    a_env = G.make(name, n, seed=8, auto_reset=True); a_env.ResetBatch()
    b_env = G.make(name, n, seed=8, auto_reset=True); b_env.ResetBatch()
    c_env = G.make(name, n, seed=8, auto_reset=True); c_env.ResetBatch()
    obs, rew, done, act = a_env.RolloutRandom(130)
    parts = [b_env.RolloutRandom(k) for k in (1, 3, 61, 8, 57)]
    for j in range(4):
        assert np.array_equal(np.concatenate([p[j] for p in parts]), (obs, rew, done, act)[j])
    for t in range(130):
        o, r, d = c_env.StepBatch(act[t])
        assert np.array_equal(o, obs[t]) and np.array_equal(d, done[t]) and np.array_equal(r, rew[t])
    sa, _, ta = a_env.GetState()
    for e in (b_env, c_env):
        se, _, te = e.GetState()
        assert te == ta == 130 and np.array_equal(se, sa)
    for e in (a_env, b_env, c_env):
        e.Close()


def test_done_compaction_matches_mask():
    n = 5000   # not a multiple of the block size: ragged tail
    env = G.CartPoleVecEnv(n, seed=2, auto_reset=True); env.ResetBatch()
    rng = np.random.default_rng(1)
    seen = 0
    for _ in range(60):
        _, _, done = env.StepBatch(rng.integers(0, 2, n).astype(np.int32))
        idx = env.DoneIndices()
        assert np.array_equal(np.sort(idx), np.nonzero(done)[0])
        seen += int(done.sum())
    assert seen > 0 and env.Stats()["episodes"] == seen
    env.Close()


def test_seed_each_and_reseed():
    n = 256
    env = G.CartPoleVecEnv(n, seed=0)
    o = O.OracleEnv(O.CARTPOLE, n, seed=0, mode=O.MODE_F32)
    seeds = np.arange(1000, 1000 + n, dtype=np.int32)
    env.Seed(seeds); o.seed_each(seeds)
    assert np.array_equal(env.ResetBatch(), o.reset())
    env.Seed(77); o.seed(77)
    assert np.array_equal(env.ResetBatch(), o.reset())
    with pytest.raises(ValueError):   # VecEnv.Seed(int[]) length check (VecEnv.cs:49)
        env.Seed(np.arange(3, dtype=np.int32))
    env.Close()


@pytest.mark.parametrize("name,limit", [("Pendulum-v1", 200), ("MountainCar-v0", 200), ("Acrobot-v1", 500)])
def test_time_limit_truncation(name, limit):
    n = 512
    env = G.make(name, n, seed=6, auto_reset=True); env.ResetBatch()
    assert env.TimeLimit == limit
    _, _, done, _ = env.RolloutRandom(limit + 1, want=("done",))
    assert done[limit - 1].all() or name == "Acrobot-v1"
    assert done[: limit - 1].sum() == 0 or name != "Pendulum-v1"
    env.Close()


def test_full_size_config2_properties():
    """BASELINE config 2 size (65 536 envs): size-independent properties of a long rollout."""
    n, k = 65536, 256
    env = G.CartPoleVecEnv(n, seed=0, auto_reset=True); env.ResetBatch()
    obs, rew, done, act = env.RolloutRandom(k)
    assert set(np.unique(act)) == {0, 1} and abs(act.mean() - 0.5) < 2e-3
    assert (rew == 1.0).all()                       # auto-reset on: never steps beyond done
    post = obs[done.astype(bool)]                   # observation returned at done is the post-reset one
    assert np.abs(post).max() <= 0.05
    alive = obs[~done.astype(bool)]
    assert np.abs(alive[:, 0]).max() <= 2.4000001 and np.abs(alive[:, 2]).max() <= 0.2094396
    lengths = k * n / max(1, done.sum())
    assert 15 < lengths < 35                        # random-policy CartPole episodes last ~22 steps
    assert env.Stats()["episodes"] == int(done.sum())
    env.Close()


def test_zero_copy_pinned_buffers_match_staged_copies():
    """gymcuda_step with page-locked host buffers (kernel reads/writes them over PCIe) == pageable path."""
    import ctypes as C
    from gymnet_b200 import _native as N
    n = 3000
    L = N.lib()
    a_env = G.CartPoleVecEnv(n, seed=4, auto_reset=True); a_env.ResetBatch()
    b_env = G.CartPoleVecEnv(n, seed=4, auto_reset=True); b_env.ResetBatch()
    nbytes = n * 4 + n * 16 + n * 4 + n
    ptr = C.c_void_p()
    N.check(L.gymcuda_host_alloc(C.byref(ptr), nbytes))
    buf = (C.c_uint8 * nbytes).from_address(ptr.value)
    raw = np.frombuffer(buf, dtype=np.uint8)
    act = raw[: n * 4].view(np.int32)
    obs = raw[n * 4: n * 20].view(np.float32).reshape(n, 4)
    rew = raw[n * 20: n * 24].view(np.float32)
    done = raw[n * 24:]
    rng = np.random.default_rng(5)
    for _ in range(40):
        a = rng.integers(0, 2, n).astype(np.int32)
        act[:] = a
        N.check(L.gymcuda_step(a_env._h, C.c_void_p(act.ctypes.data), C.c_void_p(obs.ctypes.data),
                               C.c_void_p(rew.ctypes.data), C.c_void_p(done.ctypes.data)))
        o, r, d = b_env.StepBatch(a)
        assert np.array_equal(obs, o) and np.array_equal(rew, r) and np.array_equal(done, d)
    assert np.array_equal(a_env.Observe(), b_env.Observe())
    del act, obs, rew, done, raw, buf
    N.check(L.gymcuda_host_free(ptr))
    a_env.Close(); b_env.Close()


@pytest.mark.parametrize("registered", ["out", "act", "obs_only", "all"])
def test_registered_and_mixed_host_buffers(registered):
    """gymcuda_host_register page-locks memory the caller owns (the C# shim's GCHandle-pinned result arrays): the
    step kernel then accesses it in place.  Every mix of registered and pageable buffers gives the same results,
    registrations are picked up (and dropped) between calls, and an invalid action is still reported."""
    import ctypes as C
    from gymnet_b200 import _native as N
    n = 2500
    L = N.lib()
    a_env = G.MountainCarVecEnv(n, seed=7, auto_reset=True); a_env.ResetBatch()
    b_env = G.MountainCarVecEnv(n, seed=7, auto_reset=True); b_env.ResetBatch()
    # page-aligned numpy arrays (cudaHostRegister works on any range; alignment keeps the pinned pages private)
    def arr(shape, dtype):
        count = int(np.prod(shape)); item = np.dtype(dtype).itemsize
        raw = np.empty(count * item + 8192, np.uint8)
        off = (-raw.ctypes.data) % 4096
        return raw[off: off + count * item].view(dtype).reshape(shape), raw
    act, k0 = arr((n,), np.int32); obs, k1 = arr((n, 2), np.float32); rew, k2 = arr((n,), np.float32); done, k3 = arr((n,), np.uint8)
    which = {"out": [obs, rew, done], "act": [act], "obs_only": [obs], "all": [act, obs, rew, done]}[registered]
    def step(a):
        act[:] = a
        return L.gymcuda_step(a_env._h, C.c_void_p(act.ctypes.data), C.c_void_p(obs.ctypes.data),
                              C.c_void_p(rew.ctypes.data), C.c_void_p(done.ctypes.data))
    rng = np.random.default_rng(5)
    for phase in range(3):   # pageable -> registered -> pageable again
        if phase == 1:
            for x in which:
                N.check(L.gymcuda_host_register(C.c_void_p(x.ctypes.data), x.nbytes))
        if phase == 2:
            for x in which:
                N.check(L.gymcuda_host_unregister(C.c_void_p(x.ctypes.data)))
        for _ in range(25):
            a = rng.integers(0, 3, n).astype(np.int32)
            N.check(step(a))
            o, r, d = b_env.StepBatch(a)
            assert np.array_equal(obs, o) and np.array_equal(rew, r) and np.array_equal(done, d)
        if phase == 1:
            bad = rng.integers(0, 3, n).astype(np.int32); bad[17] = 9
            assert step(bad) == N.EACTION
            with pytest.raises(N.InvalidActionError):
                b_env.StepBatch(bad)
            assert np.array_equal(obs, b_env.Observe())
    assert np.array_equal(a_env.Observe(), b_env.Observe())
    a_env.Close(); b_env.Close()


@pytest.mark.parametrize("name", ["CartPole-v1", "MountainCar-v0", "Pendulum-v1", "LunarLander-v2"])
def test_device_action_sampling_matches_rollout_and_oracle(name):
    """ActionSpace.Sample() on device: sample -> step reproduces the fused rollout; masked Discrete.Sample
    (Discrete.cs:19-25) matches the oracle restatement."""
    n, k = 512, 40
    a_env = G.make(name, n, seed=12, auto_reset=True); a_env.ResetBatch()
    b_env = G.make(name, n, seed=12, auto_reset=True); b_env.ResetBatch()
    o = O.OracleEnv(KINDS[name], n, seed=12, auto_reset=True, mode=O.MODE_F32); o.reset()
    obs, rew, done, act = a_env.RolloutRandom(k)
    for t in range(k):
        a = b_env.SampleActions()
        assert np.array_equal(a.reshape(act[t].shape), act[t])
        assert np.array_equal(a, o.sample_actions())
        ob, r, d = b_env.StepBatch(a); o.step(a)
        assert np.array_equal(ob, obs[t]) and np.array_equal(d, done[t])
    if a_env.act_n > 0:
        rng = np.random.default_rng(0)
        mask = (rng.random((n, a_env.act_n)) < 0.5).astype(np.uint8)
        mask[0] = 0                                   # nothing valid -> Start (= 0)
        mask[1] = 0; mask[1, a_env.act_n - 1] = 1     # a single valid entry
        got = b_env.SampleActions(mask)
        assert np.array_equal(got, o.sample_actions(mask))
        assert got[0] == 0 and got[1] == a_env.act_n - 1
        rows = np.arange(n)
        assert ((mask[rows, got] == 1) | (mask.sum(1) == 0)).all()
    else:
        with pytest.raises((NotImplementedError, ValueError)):
            b_env.SampleActions(np.ones((n, 1), np.uint8))
    a_env.Close(); b_env.Close()


def test_config1_single_env_10k_steps_plumbing():
    """BASELINE config 1: CartPole-v1, ONE env, 10 000 steps, random actions a_t = philox(seed 0, env 0, t) & 1,
    reset on done by the caller (README.md:36-40 loop) -- replayed through the C ABI with N = 1 and compared
    step for step with the CPU oracle (reference arithmetic for done/reward, engine twin bit for bit)."""
    env = G.CartPoleVecEnv(1, seed=0)
    twin = O.OracleEnv(O.CARTPOLE, 1, seed=0, mode=O.MODE_F32)
    ref = O.OracleEnv(O.CARTPOLE, 1, seed=0, mode=O.MODE_F64_F32STORE)
    obs = env.ResetBatch(); assert np.array_equal(obs, twin.reset()); ref.reset()
    episodes = 0
    for t in range(10000):
        a = env.SampleActions()
        assert np.array_equal(a, twin.sample_actions())
        st, ax, tt = env.GetState() if t % 500 == 0 else (None, None, None)
        if st is not None:
            ref.set_state(st.astype(np.float64), ax, tt)
            ro, rr, rd = ref.step(a)
        obs, rew, done = env.StepBatch(a)
        oo, orr, od = twin.step(a)
        assert np.array_equal(obs, oo) and rew[0] == orr[0] and done[0] == od[0]
        if st is not None:
            assert done[0] == rd[0] and rew[0] == rr[0] and np.abs(obs - ro).max() <= 1e-5
        if done[0]:
            episodes += 1
            assert np.array_equal(env.ResetBatch(), twin.reset())
    assert 300 < episodes < 700        # ~22 steps per random-policy episode
    env.Close()


@pytest.mark.parametrize("n", [1, 31, 1000, 4097])
def test_ragged_batch_sizes(n):
    env = G.CartPoleVecEnv(n, seed=3, auto_reset=True); env.ResetBatch()
    o = O.OracleEnv(O.CARTPOLE, n, seed=3, auto_reset=True, mode=O.MODE_F32); o.reset()
    tr = env.RolloutRandom(70); tw = o.rollout_random(70)
    for x, y in zip(tr, tw):
        assert np.array_equal(x, y)
    a = np.ones(n, np.int32)
    for x, y in zip(env.StepBatch(a), o.step(a)):
        assert np.array_equal(x, y)
    env.Close()


@pytest.mark.parametrize("name", ["Pendulum-v1", "Acrobot-v1"])
@pytest.mark.parametrize("n", [20, 96, 116, 1030, 4100])
def test_staged_obs_store_ragged_sizes(name, n):
    """Pendulum (3 floats) and Acrobot (6 floats) observations leave a full warp through a shared-memory tile as
    16 B vectors; a partial last warp, or n not a multiple of 4 (the row base is then not 16 B aligned), falls
    back to per-lane stores.  All shapes must give the oracle's trajectory bit for bit."""
    env = G.make(name, n, seed=4, auto_reset=True); env.ResetBatch()
    o = O.OracleEnv(KINDS[name], n, seed=4, auto_reset=True, mode=O.MODE_F32); o.reset()
    for k in (3, 37):   # unaligned start for the second launch: head + chunks + tail
        tr = env.RolloutRandom(k); tw = o.rollout_random(k)
        for x, y in zip(tr, tw):
            assert np.array_equal(x, y)
    env.Close()


@pytest.mark.parametrize("name,n", [("CartPole-v1", 65536), ("CartPole-v1", 60001), ("Pendulum-v1", 65536),
                                    ("MountainCar-v0", 75776), ("MountainCarContinuous-v0", 57000)])
def test_one_wave_cta_shape_matches_oracle(name, n):
    """Batches that fit one wave of 512-thread CTAs (3/4 .. 1 x 148 x 512 envs) run the rollout kernel in that
    shape (launch_rollout): same trajectory as the oracle, ragged last CTA and last warp included."""
    env = G.make(name, n, seed=12, auto_reset=True); env.ResetBatch()
    o = O.OracleEnv(KINDS[name], n, seed=12, auto_reset=True, mode=O.MODE_F32); o.set_threads(8); o.reset()
    for k in (5, 27):
        tr = env.RolloutRandom(k); tw = o.rollout_random(k)
        for x, y in zip(tr, tw):
            assert np.array_equal(x, y)
    env.Close()


def _rollout_from_states(name, states, k=40, seed=6):
    n = len(states)
    env = G.make(name, n, seed=seed, auto_reset=True); env.ResetBatch()
    o = O.OracleEnv(KINDS[name], n, seed=seed, auto_reset=True, mode=O.MODE_F32); o.reset()
    st, ax, t = env.GetState()
    env.SetState(states, ax, t)
    o.set_state(states.astype(np.float64), ax, t)
    tr = env.RolloutRandom(k); tw = o.rollout_random(k)
    for x, y in zip(tr, tw):
        assert np.array_equal(x, y, equal_nan=True)
    sg, _, _ = env.GetState(); so, _, _ = o.get_state()
    assert np.array_equal(sg, so.astype(np.float32), equal_nan=True)
    env.Close()
    return tr


def test_cartpole_rollout_from_out_of_range_states():
    """The unrolled chunk runs the reduced-range step (bare sincos polynomial, division without the range check)
    only for warps whose lanes all have |theta| <= pi/4 and |theta_dot| <= 64; anything else -- set through
    SetState here -- takes the generic chunk until auto-reset brings the warp back in range."""
    rng = np.random.default_rng(2)
    n = 512
    st = rng.uniform([-2, -3, -0.2, -3], [2, 3, 0.2, 3], size=(n, 4)).astype(np.float32)
    st[0:40, 2] = rng.uniform(0.8, 3.0, 40)          # beyond pi/4: Cody-Waite sincos
    st[40:80, 3] = rng.uniform(65, 500, 40)           # fast poles: generic division
    st[80:90, 2] = [1e3, -1e3, 4e4, -4e4, 1e6, 3e9, 2e14, -2e14, 0.7853982, -0.7853982]   # incl. the double reduction and NaN
    st[90:96, 3] = [64.0, -64.0, 64.00001, 1e8, 1e20, -1e30]
    st[96:100, 0] = [1e10, -1e30, 2.4, -2.4]
    tr = _rollout_from_states("CartPole-v1", st)
    assert tr[2][0, :100].sum() > 50                  # most of them terminate at once and are reset


def test_pendulum_rollout_from_huge_angles():
    """angle_normalize beyond 2^21 rad takes CUDA's fmodf, sincos beyond 32768 rad the double-precision reduction:
    both off the branch-free fast paths, both defined by IEEE alone -- same bits as the oracle's libm fmod."""
    n = 256
    rng = np.random.default_rng(3)
    st = rng.uniform([-10, -8], [10, 8], size=(n, 2)).astype(np.float32)
    st[:16, 0] = [3e4, -3e4, 32768.0, -32769.0, 1e5, -1e6, 2097152.0, 2097153.0, -2.1e6, 5e7, -6e7, 1e9, 1e12, -9e13, 3.1415927, -3.1415927]
    st[16:32, 0] = rng.uniform(-2e6, 2e6, 16)
    st[32:48, 0] = np.float32(2 * np.pi) * np.arange(-8, 8, dtype=np.float32)      # exact multiples: remainder 0
    _rollout_from_states("Pendulum-v1", st, k=24)


@pytest.mark.parametrize("name", ["MountainCar-v0", "MountainCarContinuous-v0", "Acrobot-v1"])
def test_rollout_from_edge_states(name):
    rng = np.random.default_rng(5)
    n = 256
    st = random_states(name, rng, n)
    if name == "Acrobot-v1":
        st[:8, 0] = [3.1415927, -3.1415927, 0.0, 1e4, -4e4, 2.0, 2.1, -2.1]
        st[8:12, 2] = [12.566371, -12.566371, 20.0, -20.0]
    else:
        st[:8, 0] = [-1.2, 0.6, 0.5, 0.45, 0.4999999, 0.45000002, -1.1999999, 0.0]
        st[8:12, 1] = [0.07, -0.07, 0.0, 1e-8]
    _rollout_from_states(name, st, k=24)


def test_episode_statistics_and_truncation_bits():
    """episode return/length statistics and a truncation flag distinct from termination, fused
    into the step / rollout kernels (what BasePlaySession.cs:58-69 keeps by hand)."""
    n, k = 2048, 450
    env = G.MountainCarVecEnv(n, seed=2, auto_reset=True, episode_stats=True, done_bits=True); env.ResetBatch()
    o = O.OracleEnv(O.MOUNTAINCAR, n, seed=2, auto_reset=True, mode=O.MODE_F32, done_bits=True); o.reset()
    obs, rew, done, act = env.RolloutRandom(k)
    oo, orr, od, oa = o.rollout_random(k)
    assert np.array_equal(done, od) and np.array_equal(obs, oo)
    assert set(np.unique(done)) <= {0, 1, 2} and (done == 2).sum() >= n       # the 200-step limit truncates
    st = env.Stats()
    fin = done != 0
    assert st["episodes"] == int(fin.sum())
    # expected sums straight from the trajectory
    ret = np.zeros(n); length = np.zeros(n, int); rs = 0.0; ls = 0
    for t in range(k):
        ret += rew[t]; length += 1
        f = fin[t]
        rs += ret[f].sum(); ls += int(length[f].sum()); ret[f] = 0; length[f] = 0
    assert st["length_sum"] == ls and abs(st["return_sum"] - rs) <= 1e-6 * max(1.0, abs(rs))
    # the per-launch step path accumulates the same way
    env2 = G.CartPoleVecEnv(512, seed=3, auto_reset=True, episode_stats=True); env2.ResetBatch()
    rng = np.random.default_rng(0); rs = 0.0; ls = 0; ret = np.zeros(512); length = np.zeros(512, int)
    for _ in range(120):
        _, r, d = env2.StepBatch(rng.integers(0, 2, 512).astype(np.int32))
        ret += r; length += 1; f = d != 0
        rs += ret[f].sum(); ls += int(length[f].sum()); ret[f] = 0; length[f] = 0
    s2 = env2.Stats()
    assert s2["length_sum"] == ls and abs(s2["return_sum"] - rs) < 1e-6 * max(1.0, rs) and env2.TimeLimit == 0
    env.Close(); env2.Close()


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
