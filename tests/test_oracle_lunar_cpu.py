"""CPU: physical and behavioural pins of the LunarLander oracle.  The reference's rigid-body arithmetic
lives in Aether.Physics2D (absent) so there are no reference vectors; these tests pin the restatement to
the physics it claims to implement (Newton, joint constraints, non-penetration, sleeping) and to the
reference quirks of SURVEY Appendix A that are visible from the env interface."""
import numpy as np
import pytest

import oracle_lib as O

SCALE, FPS, W, H = 30.0, 50.0, 600 / 30.0, 400 / 30.0
M_FUSELAGE = 4.816666603088379      # density 5 x area of LANDER_POLY / SCALE^2 (LunarLanderEnv.cs:189,238)


def bodies(st):
    return st[:, :21].reshape(len(st), 3, 7)   # (c.x, c.y, a, v.x, v.y, w, sleep_time)


def make(n=64, seed=5, **kw):
    e = O.OracleEnv(O.LUNARLANDER, n, seed=seed, mode=O.MODE_F32, **kw)
    return e, e.reset()


def test_reset_returns_observation_after_one_tick():
    """A.6: Reset runs Step(0) (LunarLanderEnv.cs:567-571): the lander already moves; INITIAL_RANDOM kick."""
    e, obs = make()
    st, ax, t = e.get_state()
    b = bodies(st)
    assert t == 0 and (ax[:, -1] == 1).all()
    v = b[:, 0, 3:5]
    assert np.abs(v).max() <= 1000.0 / M_FUSELAGE / FPS + 10.0 / FPS + 1e-3    # F/m*dt + g*dt
    assert np.abs(v[:, 0]).max() > 1.0                                          # the random force did act
    assert np.abs(obs[:, 0]).max() < 0.05 and (obs[:, 1] > 1.3).all()            # top centre of the viewport
    assert (obs[:, 6:] == 0).all()
    assert np.allclose(st[:, 64], obs_shaping(obs), rtol=1e-5, atol=1e-3)        # prev_shaping = shaping of that obs


def obs_shaping(o):
    return (-100 * np.hypot(o[:, 0], o[:, 1]) - 100 * np.hypot(o[:, 2], o[:, 3]) - 100 * np.abs(o[:, 4])
            + 10 * o[:, 6] + 10 * o[:, 7])


def test_free_fall_and_joint_constraints():
    e, obs = make()
    prev = bodies(e.get_state()[0])
    for step in range(30):
        obs, r, d = e.step(np.zeros(e.n, np.int32))
        st, ax, _ = e.get_state()
        b = bodies(st)
        # Newton: no engine, no contact -> the fuselage's vertical speed drops by about g*dt (legs tug a little)
        dv = b[:, 0, 4] - prev[:, 0, 4]
        assert np.abs(dv + 10.0 / FPS).max() < 0.02
        # revolute joints hold: leg anchor (+-20/30, 18/30 in leg frame) coincides with the fuselage origin
        for leg, ax_local in ((1, -20 / SCALE), (2, 20 / SCALE)):
            a_f, a_l = b[:, 0, 2], b[:, leg, 2]
            org_f = b[:, 0, :2] - rot(a_f, np.array([0.0, 0.10130719095468521]))
            anchor = b[:, leg, :2] + rot(a_l, np.array([ax_local, 18 / SCALE]) - np.array([0.03333333507180214, 0.13333334028720856]))
            assert np.abs(anchor - org_f).max() < 0.02                          # within a few linearSlop
            ref = -0.05 if leg == 1 else 0.05
            ang = a_l - a_f - ref
            lo, hi = (0.4, 0.9) if leg == 1 else (-0.9, -0.4)
            # limits (:276-277): the legs are created folded (angle 0) and the position solver opens them over the
            # first ~8 steps (each 8-degree limit correction is mostly undone by the point constraint of the very light leg)
            if step >= 12:
                assert (ang >= lo - 0.08).all() and (ang <= hi + 0.08).all()
        prev = b
    assert (d == 0).all()


def rot(a, v):
    c, s = np.cos(a), np.sin(a)
    return np.stack([c * v[0] - s * v[1], s * v[0] + c * v[1]], axis=-1)


def test_main_engine_impulse():
    """ApplyLinearImpulse of -o*13 at the nozzle (:655-670): |dv| = 13*|o|/m with |o| in [4/30 - 2/30, 4/30 + 2/30]+."""
    e, obs = make(n=256)
    st0 = bodies(e.get_state()[0])[:, 0]
    obs, r, d = e.step(np.full(e.n, 2, np.int32))
    st1 = bodies(e.get_state()[0])[:, 0]
    dv = st1[:, 3:5] - st0[:, 3:5] + np.array([0.0, 10.0 / FPS])
    mag = np.hypot(dv[:, 0], dv[:, 1])
    assert mag.min() > 13 * (2 / SCALE) / M_FUSELAGE * 0.85 and mag.max() < 13 * (6.5 / SCALE) / M_FUSELAGE * 1.15
    assert (dv[:, 1] > 0).all()                                                 # pushes up when upright


def test_reward_is_shaping_difference_minus_fuel():
    e, obs = make(n=128)
    prev = obs_shaping(obs)
    for a, fuel in ((0, 0.0), (2, 0.3), (1, 0.03), (3, 0.03)):
        obs, r, d = e.step(np.full(e.n, a, np.int32))
        sh = obs_shaping(obs)
        assert np.abs(r - (sh - prev - fuel)).max() < 2e-3                      # :748-760
        prev = sh


def test_landing_contacts_sleep_and_terminal_rewards():
    e, obs = make(n=96, seed=1000)
    ended = np.zeros(e.n, bool); final = np.zeros(e.n)
    max_pen = 0.0
    for t in range(900):
        a = pid(obs)
        obs, r, d = e.step(a)
        final = np.where(~ended & (d > 0), r, final)
        ended |= d > 0
        st, ax, _ = e.get_state()
        # non-penetration: no leg vertex sinks more than a few slops below the helipad plane while resting on it
        b = bodies(st)
        on_pad = (np.abs(obs[:, 0]) < 0.15) & ((obs[:, 6] > 0) | (obs[:, 7] > 0)) & ~ended
        if on_pad.any():
            low = np.minimum(b[on_pad, 1, 1], b[on_pad, 2, 1]) - 0.14           # leg centre - half height bound
            max_pen = max(max_pen, float(np.max(0.33 * 3 * H / 4 - low - 0.3)))
        if ended.all():
            break
    assert set(np.unique(final[ended])) <= {-100.0, 100.0}                      # :762-771
    assert (final == 100.0).sum() >= 5                                           # landers do come to rest and fall asleep
    assert max_pen < 0.1


def pid(s):
    at = np.clip(s[:, 0] * 0.5 + s[:, 2], -0.4, 0.4); ht = 0.55 * np.abs(s[:, 0])
    atd = (at - s[:, 4]) * 0.5 - s[:, 5]; htd = (ht - s[:, 1]) * 0.5 - s[:, 3] * 0.5
    legs = (s[:, 6] > 0) | (s[:, 7] > 0)
    atd = np.where(legs, 0.0, atd); htd = np.where(legs, -s[:, 3] * 0.5, htd)
    a = np.zeros(len(s), np.int32); a[atd > 0.05] = 1; a[atd < -0.05] = 3
    a[(htd > np.abs(atd)) & (htd > 0.05)] = 2
    return a


def test_one_sided_out_of_view_quirk():
    """A.5: done only for pos.x > 1, never for < -1 (LunarLanderEnv.cs:762)."""
    e, obs = make(n=2)
    st, ax, t = e.get_state()
    for i, x in enumerate((W + 1.0, -1.0 - W * 0.05)):       # right of the viewport / left of it
        shift = x - st[i, 0]
        for body in range(3):
            st[i, 7 * body] += shift
    e.set_state(st, ax, t)
    obs, r, d = e.step(np.zeros(2, np.int32))
    assert obs[0, 0] > 1 and d[0] == 1 and r[0] == -100.0
    assert obs[1, 0] < -1 and d[1] == 0


def test_invalid_action_and_continuous_mode():
    e, obs = make(n=4)
    before = e.get_state()[0].copy()
    e.step(np.array([0, 4, -1, 3], np.int32))
    assert e.invalid == 2
    after = e.get_state()[0]
    assert np.array_equal(after[1], before[1]) and np.array_equal(after[2], before[2])
    c = O.OracleEnv(O.LUNARLANDER_CONT, 8, seed=2, mode=O.MODE_F32); c.reset()
    # main engine only fires for a0 > 0 (m_power in [0.5, 1], :640); side engines for |a1| > 0.5 (:620)
    acts = np.array([[-1, 0], [0.0, 0.4], [1, 0], [0.5, 0], [0, 1], [0, -1], [0, 0.6], [2, -3]], np.float32)
    s0 = bodies(c.get_state()[0])[:, 0].copy()
    c.step(acts)
    s1 = bodies(c.get_state()[0])[:, 0]
    dv = s1[:, 4] - s0[:, 4] + 10.0 / FPS
    assert np.abs(dv[:2]).max() < 0.01 and (dv[2:4] > 0.1).all()
    dw = np.abs(s1[:, 5] - s0[:, 5])
    assert dw[4] > 1e-3 and dw[5] > 1e-3 and dw[6] > 1e-3 and dw[0] < 1e-2


def test_wind_phase_persists_across_episodes():
    """A.9: _wind_idx / _torque_idx are constructor draws (:409-410), advanced only while airborne (:588-596)."""
    e = O.OracleEnv(O.LUNARLANDER, 8, seed=4, mode=O.MODE_F32)
    O.lib().oracle_set_lunar_params(e.h, -10.0, 1, 15.0, 1.5)
    e.reset()
    a0 = e.get_state()[1]
    assert (np.abs(a0[:, 24]) <= 9999 + 1).all() and (a0[:, 24] != a0[0, 24]).any()   # per-env phase in [-9999, 9999)
    e.step(np.zeros(8, np.int32))
    a1 = e.get_state()[1]
    assert (a1[:, 24] == a0[:, 24] + 1).all()
    e.reset()
    assert (e.get_state()[1][:, 24] == a1[:, 24] + 1).all()      # reset's zero step advances it; it is NOT re-drawn


def test_mass_data_constants():
    """The hard-coded polygon mass data (oracle/lunar.hpp and the CUDA twin) equal b2PolygonShape::ComputeMass in float32."""
    import importlib.util, os, re
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("lunar_mass_data", os.path.join(here, "golden", "lunar_mass_data.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    c = mod.constants()
    root = os.path.dirname(here)
    for path in (os.path.join(root, "oracle", "lunar.hpp"), os.path.join(root, "gym.net_b200", "csrc", "lunar_core.cuh")):
        text = open(path).read()
        for part in ("fuselage", "leg"):
            for key in ("mass", "inv_mass", "inertia", "inv_inertia"):
                assert repr(c[part][key]) + "f" in text, "%s: %s.%s = %r not found" % (path, part, key, c[part][key])
            assert repr(c[part]["centroid"][1]) + "f" in text
