"""CPU: pin the oracle.  Known-answer vectors for the generator, accuracy of detmath, the golden
fixtures (independent Python restatement, tests/golden/make_golden.py), the reference quirks of
SURVEY Appendix A, and the agreement of the two oracle arithmetics (F64 = reference, F32 = engine)."""
import math
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-5


# ---------------------------------------------------------------- Philox4x32-10
def test_philox_random123_known_answers():
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        assert [int(v) for v in O.philox4x32_10(ctr, key)] == want


def test_draw_addressing():
    seed, env, index = 0x1234567890abcdef, 77, (5 << 32) | 9
    got = O.draw(seed, env, index, 2, 3)
    want = O.philox4x32_10([9, 5, 2 | (3 << 8), 0x12345678], [0x90abcdef, 77])
    assert np.array_equal(got, want)


def test_random_policy_streams():
    e = O.OracleEnv(O.CARTPOLE, 64, seed=5, auto_reset=True, mode=O.MODE_F32); e.reset()
    _, _, _, a = e.rollout_random(512)
    assert set(np.unique(a)) == {0, 1} and abs(a.mean() - 0.5) < 0.02
    # bit b of word w of block t>>7
    blk = O.draw(5, 3, 1, 1)
    assert a[128 + 37, 3] == (int(blk[1]) >> 5) & 1
    e = O.OracleEnv(O.ACROBOT, 64, seed=5, auto_reset=True, mode=O.MODE_F32); e.reset()
    _, _, _, a = e.rollout_random(300)
    assert set(np.unique(a)) == {0, 1, 2}
    assert a[9, 2] == (int(O.draw(5, 2, 2, 1)[1]) * 3) >> 32
    e = O.OracleEnv(O.PENDULUM, 64, seed=5, auto_reset=True, mode=O.MODE_F32); e.reset()
    _, _, _, a = e.rollout_random(300)
    assert a.min() >= -2 and a.max() < 2 and abs(a.mean()) < 0.05


# ---------------------------------------------------------------- detmath
def test_sincosf_det_accuracy():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-0.8, 0.8, 300000), rng.uniform(-100, 100, 300000),
                        rng.uniform(-1e5, 1e5, 300000), rng.uniform(3e4, 1e7, 50000),
                        [0.0, -0.0, 0.7853981852531433, -0.7853981852531433, 32768.0, 1e14]]).astype(np.float32)
    s, c = O.sincosf(x)
    xd = x.astype(np.float64)
    small = np.abs(xd) <= 1e5
    for got, ref in ((s, np.sin(xd)), (c, np.cos(xd))):
        ulp = np.spacing(np.abs(ref.astype(np.float32))).astype(np.float64)
        assert (np.abs(got - ref)[small] / ulp[small]).max() <= 2.0
        assert np.abs(got - ref).max() <= 2e-7 * 4   # absolute, everywhere up to 1e14
    s, c = O.sincosf(np.array([np.inf, np.nan, 3e20], np.float32))
    assert np.isnan(s).all() and np.isnan(c).all()


# ---------------------------------------------------------------- golden fixtures vs the C oracle
def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _oracle_step(kind, state, action, aux=None, mode=O.MODE_F64):
    n = len(state)
    ax = np.zeros((n, 3), np.int32); ax[:, 0] = -1
    if aux is not None:
        ax[:, 0] = aux
    e = O.OracleEnv(kind, n, seed=0, time_limit=-1, mode=mode)
    e.reset(); e.set_state(np.asarray(state, np.float64), ax, 0)
    obs, rew, done = e.step(action)
    st, ax2, _ = e.get_state()
    return obs, rew, done, st, ax2


def test_golden_cartpole_bit_exact():
    g = _load("cartpole")
    obs, rew, done, st, ax = _oracle_step(O.CARTPOLE, g["state"], g["action"], g["sbd"])
    assert np.array_equal(done, g["done"]) and np.array_equal(rew, g["reward"]) and np.array_equal(ax[:, 0], g["next_sbd"])
    assert np.array_equal(st, g["next_state"])      # same libm, same operation order: identical doubles
    assert 0 < done.sum() < len(done)


@pytest.mark.parametrize("name,kind", [("pendulum", O.PENDULUM), ("mountaincar", O.MOUNTAINCAR),
                                       ("mountaincar_cont", O.MOUNTAINCAR_CONT), ("acrobot", O.ACROBOT)])
def test_golden_upstream_envs(name, kind):
    g = _load(name)
    obs, rew, done, st, _ = _oracle_step(kind, g["state"], g["action"])
    assert np.array_equal(done, g["done"])
    assert np.abs(st - g["next_state"]).max() <= 1e-12
    assert np.abs(rew - g["reward"]).max() <= 1e-5


# ---------------------------------------------------------------- reference quirks
def test_cartpole_constants_are_float32_rounded():
    # A.1: thresholds are float32 (CartPoleEnv.cs:34,36), not the upstream doubles
    x = np.array([[2.4000000953674316 - 0.02 * 1.0, 1.0, 0.0, 0.0],      # lands exactly ON x_thr: not done (strict >)
                  [np.nextafter(np.float32(2.4), np.float32(3)) - 0.0, 0.0, 0.0, 0.0],
                  [0.0, 0.0, 0.20943951606750488, 0.0],                    # exactly ON theta_thr: not done
                  [0.0, 0.0, 0.2094395160675049 + 3e-17, 0.0]])            # double just above: done
    obs, rew, done, st, ax = _oracle_step(O.CARTPOLE, x, np.zeros(4, np.int32))
    assert list(done) == [int(st[0, 0] > 2.4000000953674316), 1, 0, 1]


def test_cartpole_no_step_limit_and_reset_range():
    # A.2: no TimeLimit in the reference; reset = uniform(-0.05, 0.05, 4) with steps_beyond_done = -1
    e = O.OracleEnv(O.CARTPOLE, 256, seed=3, auto_reset=True, mode=O.MODE_F64)
    obs = e.reset()
    assert np.abs(obs).max() <= 0.05 and np.abs(obs).max() > 0.04
    st, ax, t = e.get_state()
    assert (ax[:, 0] == -1).all() and (ax[:, 2] == 1).all() and t == 0
    e2 = O.OracleEnv(O.CARTPOLE, 1, seed=3, mode=O.MODE_F64); obs = e2.reset()
    for i in range(700):                      # a PD controller keeps it up past CartPole-v1's 500: never truncated
        x, xd, th, thd = obs[0]
        obs, _, d = e2.step(np.array([1 if (th + 0.3 * thd + 0.05 * x + 0.1 * xd) > 0 else 0], np.int32))
        assert d[0] == 0


def test_invalid_action_policies():
    e = O.OracleEnv(O.CARTPOLE, 2, seed=0, mode=O.MODE_F64); e.reset()
    e.step(np.array([5, 0], np.int32)); assert e.invalid == 0     # Debug.Assert only (CartPoleEnv.cs:139)
    e = O.OracleEnv(O.MOUNTAINCAR, 2, seed=0, mode=O.MODE_F64); before = e.reset()
    obs, _, _ = e.step(np.array([3, 0], np.int32))
    assert e.invalid == 1 and np.array_equal(obs[0], before[0])


# ---------------------------------------------------------------- the two arithmetics agree
@pytest.mark.parametrize("name,kind,lo,hi,scale", [
    ("cartpole", O.CARTPOLE, [-2.6, -3, -0.25, -3.5], [2.6, 3, 0.25, 3.5], [2.4, 1, 0.21, 1]),
    ("pendulum", O.PENDULUM, [-10, -8], [10, 8], [3.14, 8]),
    ("mountaincar", O.MOUNTAINCAR, [-1.2, -0.07], [0.6, 0.07], [1.2, 0.07]),
    ("mountaincar_cont", O.MOUNTAINCAR_CONT, [-1.2, -0.07], [0.6, 0.07], [1.2, 0.07]),
    ("acrobot", O.ACROBOT, [-3.14, -3.14, -6, -12], [3.14, 3.14, 6, 12], [3.14, 3.14, 12, 28]),
])
def test_engine_arithmetic_vs_reference_arithmetic(name, kind, lo, hi, scale):
    rng = np.random.default_rng(1)
    n = 200000
    s = rng.uniform(lo, hi, size=(n, len(lo))).astype(np.float32)
    d = O.dims(kind)
    a = (rng.integers(0, d["act_n"], n).astype(np.int32) if d["act_n"] else rng.uniform(-2, 2, n).astype(np.float32))
    o64, r64, d64, s64, _ = _oracle_step(kind, s, a, mode=O.MODE_F64_F32STORE)
    o32, r32, d32, s32, _ = _oracle_step(kind, s, a, mode=O.MODE_F32)
    assert np.array_equal(d64, d32)                      # termination is bit-exact
    diff = s32 - s64
    if name == "acrobot":
        diff[:, :2] = (diff[:, :2] + np.pi) % (2 * np.pi) - np.pi
    den = np.maximum(np.maximum(np.abs(s32), np.abs(s64)), np.array(scale))
    assert (np.abs(diff) / den).max() <= RTOL
    if name == "cartpole":                               # positions come from the same double operations
        assert np.array_equal(s32[:, 0], s64[:, 0]) and np.array_equal(s32[:, 2], s64[:, 2])


def test_cartpole_adversarial_thresholds_cpu():
    """States whose successor sits within a few ulps of +-x_thr / +-theta_thr: F32 engine arithmetic
    must give the reference's `done` (it evaluates the position update in double)."""
    rng = np.random.default_rng(2)
    n = 100000
    tau = np.float64(np.float32(0.02)); xthr = np.float64(np.float32(2.4))
    s = rng.uniform([-2.6, -3, -0.2, -3], [2.6, 3, 0.2, 3], size=(n, 4)).astype(np.float32)
    sign = rng.choice([-1.0, 1.0], n)
    s[:, 0] = (sign * xthr - tau * s[:, 1].astype(np.float64)).astype(np.float32)
    a = rng.integers(0, 2, n).astype(np.int32)
    _, _, d64, _, _ = _oracle_step(O.CARTPOLE, s, a, mode=O.MODE_F64)
    _, _, d32, _, _ = _oracle_step(O.CARTPOLE, s, a, mode=O.MODE_F32)
    assert np.array_equal(d64, d32) and 0.2 < d64.mean() < 0.8


def test_partition_invariance_cpu():
    full = O.OracleEnv(O.CARTPOLE, 64, seed=9, auto_reset=True, mode=O.MODE_F32); full.reset()
    fo, _, fd, fa = full.rollout_random(100)
    for part in range(4):
        sh = O.OracleEnv(O.CARTPOLE, 16, seed=9, env_id_offset=16 * part, auto_reset=True, mode=O.MODE_F32); sh.reset()
        so, _, sd, sa = sh.rollout_random(100)
        assert np.array_equal(so, fo[:, 16 * part:16 * part + 16]) and np.array_equal(sa, fa[:, 16 * part:16 * part + 16])


def test_free_running_drift_report_runs():
    """SURVEY 8(c) T2, free-running half (tests/tools/f32_vs_f64_drift.py): the float32 engine arithmetic against the float64
    reference arithmetic with NO teacher forcing.  CartPole's episodes are short and its termination test is evaluated
    exactly, so the two agree on (almost) every `done` for hundreds of steps; chaotic Pendulum / Acrobot trajectories
    separate by construction -- which is why the parity criterion is per-step (teacher-forced), not per-trajectory."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("drift", os.path.join(root, "tests", "tools", "f32_vs_f64_drift.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    r = mod.drift("CartPole-v1", 512, 300)
    assert r["frac_diverged"] <= 0.01 and r["max_rel_state_err_while_in_step"] < 0.05
    r = mod.drift("MountainCar-v0", 256, 300)
    assert r["diverged"] == 0 and r["max_rel_state_err_while_in_step"] < 1e-4
