"""CPU: the engine's DEVICE source, compiled for the host, against the oracle -- bit for bit.

tests/hostsim/hostsim.cpp includes gym.net_b200/csrc/{detmath,philox,env_classic,lunar,lunar_core}.cuh through a stub
<cuda_runtime.h> (every __device__ function becomes plain C++, -ffp-contract=off like nvcc -fmad=false) and replays
the per-thread body of step_kernel / reset_kernel.  What the GPU tests establish through the C ABI on a B200 --
kernel == oracle F32 on free-running trajectories with auto-reset -- is established here for the same source lines
on any machine: reset draws, action validation, transitions, rewards, termination, time limits, the LunarLander
solver through landings and crashes."""
import numpy as np
import pytest

import oracle_lib as O

from hostsim_lib import DEFAULT_LIMIT, PRM, HostSim, _p, build


@pytest.fixture(scope="module")
def hs():
    return build()


CASES = [("CartPole-v1", O.CARTPOLE, 256, 300), ("Pendulum-v1", O.PENDULUM, 128, 450), ("MountainCar-v0", O.MOUNTAINCAR, 128, 450),
         ("MountainCarContinuous-v0", O.MOUNTAINCAR_CONT, 64, 1020), ("Acrobot-v1", O.ACROBOT, 128, 600),
         ("LunarLander-v2", O.LUNARLANDER, 96, 400), ("LunarLanderContinuous-v2", O.LUNARLANDER_CONT, 48, 300)]


@pytest.mark.parametrize("name,kind,n,k", CASES, ids=[c[0] for c in CASES])
def test_device_source_on_host_equals_oracle(hs, name, kind, n, k):
    seed, off = 21, 1000
    o = O.OracleEnv(kind, n, seed=seed, env_id_offset=off, auto_reset=True, mode=O.MODE_F32)
    sim = HostSim(hs, kind, n, seed, off)
    assert np.array_equal(sim.reset(), o.reset())
    episodes = 0
    for t in range(k):
        a = o.sample_actions()                       # the random policy of the rollout kernel, from the oracle
        oo, orr, od = o.step(a)
        so, sr, sdn, bad = sim.step(a)
        assert bad == 0
        assert np.array_equal(sdn, od), "done differs at step %d" % t
        assert np.array_equal(sr, orr), "reward differs at step %d" % t
        assert np.array_equal(so, oo), "observation differs at step %d" % t
        episodes += int(od.sum())
        if t % 97 == 0 or t == k - 1:
            st, aux, ot = o.get_state()
            assert ot == sim.t
            assert np.array_equal(sim.abi_state(), st.astype(np.float32)), "state differs at step %d" % t
    assert episodes > 0


def test_invalid_actions_leave_the_env_unstepped(hs):
    n = 8
    sim = HostSim(hs, O.MOUNTAINCAR, n, 3); sim.reset()
    before = sim.abi_state()
    a = np.array([0, 1, 2, 3, -1, 2, 1, 0], np.int32)
    obs, rew, done, bad = sim.step(a)
    assert bad == 2
    after = sim.abi_state()
    assert np.array_equal(after[[3, 4]], before[[3, 4]]) and not np.array_equal(after[[0, 1, 2]], before[[0, 1, 2]])


def test_div_inrange_is_the_ieee_quotient_on_its_stated_ranges(hs):
    """detmath.cuh div_inrange against numpy's float32 division on the operand ranges its callers state
    (CartPole: den in [0.62, 0.67], |num| up to 2^8; Acrobot: den in [0.56, 4.5], |num| up to ~1e5), zero included.
    (The host build seeds the Newton step with 1.0f / y instead of MUFU.RCP; the GPU tests cover the real one.)"""
    rng = np.random.default_rng(0)
    m = 1 << 20
    y = np.concatenate([rng.uniform(0.62, 0.67, m), rng.uniform(0.56, 4.5, m)]).astype(np.float32)
    x = np.concatenate([rng.uniform(-256, 256, m), rng.standard_normal(m) * 1e4]).astype(np.float32)
    x[:1000] = 0.0
    x[1000:2000] = np.float32(2.0) ** rng.integers(-60, 8, 1000).astype(np.float32)
    q = np.empty_like(x)
    hs.hostsim_div_inrange(_p(x), _p(y), _p(q), x.size)
    assert np.array_equal(q, x / y)


def test_sincos_det_host_build_equals_oracle(hs):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-40000, 40000, 200000), rng.uniform(-1, 1, 100000), rng.standard_normal(1000) * 1e9,
                        [0.0, 0.7853981852531433, -0.7853981852531433, 0.78539824, 32768.0, -32768.0, 32769.0, 1e14, 2e14]]).astype(np.float32)
    s = np.empty_like(x); c = np.empty_like(x); so = np.empty_like(x); co = np.empty_like(x)
    hs.hostsim_sincos(_p(x), _p(s), _p(c), x.size)
    O.lib().oracle_sincosf(_p(x), _p(so), _p(co), x.size)
    assert np.array_equal(s, so, equal_nan=True) and np.array_equal(c, co, equal_nan=True)


ROLLOUT_CASES = [("CartPole-v1", O.CARTPOLE, 200), ("Pendulum-v1", O.PENDULUM, 201), ("MountainCar-v0", O.MOUNTAINCAR, 200),
                 ("MountainCarContinuous-v0", O.MOUNTAINCAR_CONT, 200), ("Acrobot-v1", O.ACROBOT, 201), ("LunarLander-v2", O.LUNARLANDER, 40)]


@pytest.mark.parametrize("name,kind,n", ROLLOUT_CASES, ids=[c[0] for c in ROLLOUT_CASES])
@pytest.mark.parametrize("shape", ["all_out_64", "all_out_512", "generic_64"])
def test_rollout_kernel_source_on_host_equals_oracle(hs, name, kind, n, shape):
    """kernels.cuh's rollout_kernel compiled for the host and run one thread at a time (one-lane warps; the staged
    observation store, which needs a real warp, is kept off by n % 4 != 0 for 3- and 6-float observations): the action
    generator, head / unrolled chunks / tail for launches that start at unaligned step indices, the reduced-range and
    limit-free chunk variants, the pre-generated resets and the 32-bit row index -- against the oracle's rollout."""
    all_out = shape != "generic_64"
    block = 512 if shape == "all_out_512" else 64
    if block == 512 and kind in (O.ACROBOT, O.LUNARLANDER):
        pytest.skip("no one-wave shape for envs whose step is not unrolled")
    seed, off = 5, 77
    o = O.OracleEnv(kind, n, seed=seed, env_id_offset=off, auto_reset=True, mode=O.MODE_F32)
    sim = HostSim(hs, kind, n, seed, off)
    assert np.array_equal(sim.reset(), o.reset())
    episodes = 0
    launches = {O.LUNARLANDER: (3, 37, 120), O.ACROBOT: (3, 37, 8, 210, 260)}.get(kind, (3, 37, 8, 210))   # Acrobot: past its 500-step limit
    for k in launches:
        tr = sim.rollout(k, all_out, block)
        tw = o.rollout_random(k)
        for j, what in enumerate(("obs", "reward", "done", "actions")):
            assert np.array_equal(tr[j], tw[j]), "%s differs in the launch of %d steps" % (what, k)
        episodes += int(tw[2].sum())   # (the kernel's own episode counter is a warp reduction: meaningless with one-lane warps)
        st, _, ot = o.get_state()
        assert ot == sim.t and np.array_equal(sim.abi_state(), st.astype(np.float32))
    assert episodes > 0 or kind == O.MOUNTAINCAR_CONT


def test_cartpole_teacher_forced_ten_million_states(hs):
    """SURVEY section 7 step 3 at its stated size, on the CPU: 10^7 teacher-forced CartPole states -- uniform over and
    beyond the episode range, plus states constructed so that the successor's x or theta lands within a few float32
    ulps of +-x_threshold / +-theta_threshold -- through (1) the reference arithmetic (oracle F64: the C# doubles of
    CartPoleEnv.cs:137-186), (2) the engine arithmetic (oracle F32) and (3) the DEVICE source compiled for the host.
    done: identical in all three.  State: (3) == (2) bit for bit, and within 1e-5 of (1)."""
    rng = np.random.default_rng(7)
    chunk, chunks = 1_000_000, 10
    tau = np.float64(np.float32(0.02)); xthr = np.float64(np.float32(2.4)); ththr = np.float64(np.float32(12 * 2 * np.pi / 360))
    n_done = 0
    for c in range(chunks):
        s = rng.uniform([-2.6, -3, -0.25, -3.5], [2.6, 3, 0.25, 3.5], size=(chunk, 4)).astype(np.float32)
        k = chunk // 4                                  # a quarter each: successor x / theta at a threshold (+- a few ulps)
        sign = rng.choice([-1.0, 1.0], 2 * k)
        jit = rng.integers(-3, 4, 2 * k)
        x0 = (sign[:k] * xthr - tau * s[:k, 1].astype(np.float64)).astype(np.float32)
        s[:k, 0] = np.nextafter(x0, np.where(jit[:k] > 0, np.float32(np.inf), np.float32(-np.inf))) if c % 2 else x0
        t0 = (sign[k:] * ththr - tau * s[k:2 * k, 3].astype(np.float64)).astype(np.float32)
        s[k:2 * k, 2] = np.nextafter(t0, np.where(jit[k:] > 0, np.float32(np.inf), np.float32(-np.inf))) if c % 2 else t0
        a = rng.integers(0, 2, chunk).astype(np.int32)
        ax = np.zeros((chunk, 3), np.int32); ax[:, 0] = -1
        out = {}
        for mode in (O.MODE_F64, O.MODE_F32):
            e = O.OracleEnv(O.CARTPOLE, chunk, seed=0, time_limit=-1, mode=mode)
            e.set_threads(8)
            e.reset(); e.set_state(s.astype(np.float64), ax, 0)
            obs, rew, done = e.step(a)
            out[mode] = (done, e.get_state()[0], rew)
            e.close()
        d64, s64, r64 = out[O.MODE_F64]; d32, s32, r32 = out[O.MODE_F32]
        assert np.array_equal(d64, d32), "engine arithmetic disagrees with the reference arithmetic on done (chunk %d)" % c
        assert np.array_equal(r64, r32)
        den = np.maximum(np.maximum(np.abs(s32), np.abs(s64)), [2.4, 1, 0.21, 1])
        assert (np.abs(s32 - s64) / den).max() <= 1e-5
        sim = HostSim(hs, O.CARTPOLE, chunk, 0, auto_reset=False)
        sim.state[:] = s
        so, sr, sd, bad = sim.step(a)
        assert bad == 0 and np.array_equal(sd, d32) and np.array_equal(sr, r32)
        assert np.array_equal(sim.abi_state(), s32.astype(np.float32))
        n_done += int(d32.sum())
    assert 0.2 < n_done / (chunk * chunks) < 0.8


@pytest.mark.parametrize("name,kind,lo,hi,scale,rtol", [
    ("Pendulum-v1", O.PENDULUM, [-40, -8], [40, 8], [3.14, 8], 1e-5),
    ("MountainCar-v0 near the goal", O.MOUNTAINCAR, [0.40, -0.01], [0.56, 0.07], [1.2, 0.07], 1e-5),
    ("MountainCarContinuous-v0 near the goal", O.MOUNTAINCAR_CONT, [0.36, -0.01], [0.52, 0.07], [1.2, 0.07], 1e-5),
    ("MountainCar-v0 at the left wall", O.MOUNTAINCAR, [-1.2, -0.07], [-1.1, 0.02], [1.2, 0.07], 1e-5),
    ("Acrobot-v1 around the terminal height", O.ACROBOT, [1.6, -2.2, -3, -5], [3.14, 2.2, 3, 5], [3.14, 3.14, 12, 28], 1e-5),
    ("Acrobot-v1 fast", O.ACROBOT, [-3.14, -3.14, -9, -18], [3.14, 3.14, 9, 18], [3.14, 3.14, 12, 28], 1e-5),
    # the WHOLE velocity clamp box (|dtheta1| <= 4 pi, |dtheta2| <= 9 pi): towards its corners one RK4 step of 0.2 s changes
    # the velocities by tens of rad/s and amplifies float32 rounding about a hundredfold, so beyond (9, 18) rad/s the engine
    # steps in double precision (engine arithmetic v3) and the 1e-5 of north_star holds everywhere
    ("Acrobot-v1 whole clamp box", O.ACROBOT, [-3.1416, -3.1416, -12.5664, -28.2744], [3.1416, 3.1416, 12.5664, 28.2744], [3.14, 3.14, 12, 28], 1e-5),
], ids=lambda v: v if isinstance(v, str) else None)
def test_upstream_envs_teacher_forced_two_million_states(hs, name, kind, lo, hi, scale, rtol):
    """2 x 10^6 teacher-forced states per case, concentrated where `done`, the clamps and the wall rule decide: reference
    arithmetic (oracle F64) vs engine arithmetic (oracle F32) vs the device source on the host -- done identical in all
    three, device source == engine arithmetic bit for bit, state within `rtol` (1e-5) of the reference arithmetic."""
    rng = np.random.default_rng(11)
    chunk = 500_000
    d = O.dims(kind)
    seen_done = 0
    for c in range(4):
        s = rng.uniform(lo, hi, size=(chunk, len(lo))).astype(np.float32)
        a = (rng.integers(0, d["act_n"], chunk).astype(np.int32) if d["act_n"] else rng.uniform(-2.5, 2.5, (chunk, 1)).astype(np.float32))
        ax = np.zeros((chunk, 3), np.int32); ax[:, 0] = -1
        out = {}
        for mode in (O.MODE_F64_F32STORE, O.MODE_F32):
            e = O.OracleEnv(kind, chunk, seed=0, time_limit=-1, mode=mode)
            e.set_threads(8)
            e.reset(); e.set_state(s.astype(np.float64), ax, 0)
            obs, rew, done = e.step(a)
            out[mode] = (done, e.get_state()[0], rew, obs)
            e.close()
        d64, s64, r64, o64 = out[O.MODE_F64_F32STORE]; d32, s32, r32, o32 = out[O.MODE_F32]
        assert np.array_equal(d64, d32), "done differs between the two arithmetics (chunk %d)" % c
        diff = s32 - s64
        if kind == O.ACROBOT:
            diff[:, :2] = (diff[:, :2] + np.pi) % (2 * np.pi) - np.pi
        den = np.maximum(np.maximum(np.abs(s32), np.abs(s64)), np.array(scale))
        assert (np.abs(diff) / den).max() <= rtol
        sim = HostSim(hs, kind, chunk, 0, auto_reset=False)
        sim.limit = 0
        sim.state[:] = s
        so, sr, sd, bad = sim.step(a)
        assert bad == 0 and np.array_equal(sd, d32) and np.array_equal(sr, r32) and np.array_equal(so, o32)
        assert np.array_equal(sim.abi_state(), s32.astype(np.float32))
        seen_done += int(d32.sum())
    if kind != O.PENDULUM and "wall" not in name:
        assert 0.02 < seen_done / (4 * chunk) < 0.98


@pytest.mark.parametrize("kind", [O.LUNARLANDER, O.LUNARLANDER_CONT], ids=["discrete", "continuous"])
def test_lunarlander_wind_and_gravity_parameters_on_host(hs, kind):
    """The constructor parameters of LunarLanderEnv (LunarLanderEnv.cs:383-410): wind and turbulence on (the double
    tanh(sin(..) + sin(..)) of :588-596 with its two phases drawn once per env, kept across episodes) and a weaker gravity,
    device source on the host vs oracle, bit for bit (both call the host libm here; on the GPU the wind terms may differ
    from libm in an ulp, which tests/test_gpu_lunar.py allows for)."""
    n, seed, off = 64, 9, 300
    prm = (-6.5, 18.0, 1.9, 1)
    o = O.OracleEnv(kind, n, seed=seed, env_id_offset=off, auto_reset=True, time_limit=120, mode=O.MODE_F32)
    O.lib().oracle_set_lunar_params(o.h, prm[0], prm[3], prm[1], prm[2])
    sim = HostSim(hs, kind, n, seed, off, prm=prm)
    sim.limit = 120
    assert np.array_equal(sim.reset(), o.reset())
    for k in (5, 120, 150):
        tr = sim.rollout(k, True, 64)
        tw = o.rollout_random(k)
        for j, what in enumerate(("obs", "reward", "done", "actions")):
            assert np.array_equal(tr[j], tw[j]), "%s differs in the launch of %d steps" % (what, k)
    st, _, ot = o.get_state()
    assert ot == sim.t and np.array_equal(sim.abi_state(), st.astype(np.float32))
    # the wind really blew: the same envs without it fly differently
    first = {}
    for use_wind in (0, 1):
        e = O.OracleEnv(kind, n, seed=seed, env_id_offset=off, auto_reset=True, time_limit=120, mode=O.MODE_F32)
        O.lib().oracle_set_lunar_params(e.h, prm[0], use_wind, prm[1], prm[2])
        e.reset()
        first[use_wind] = e.rollout_random(60)[0]
    assert not np.array_equal(first[0], first[1])
