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
