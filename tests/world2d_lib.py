"""ctypes binding of oracle/world2d/libworld2d.so -- the structurally independent LunarLander oracle (a generic
Box2D-2.3-lineage engine with LunarLanderEnv.cs built on top).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W2D_DIR = os.path.join(ROOT, "oracle", "world2d")
LIB_PATH = os.path.join(W2D_DIR, "libworld2d.so")

STATE_DIM, AUX_DIM = 80, 29
STREAM_RESET, STREAM_ACTION, STREAM_DYNAMICS, STREAM_CTOR = 0, 1, 2, 3
F_GAME_OVER, F_LEG0, F_LEG1, F_FUSELAGE, F_AWAKE, F_FIRST_STEP, F_CONTINUOUS = 1, 2, 4, 8, 16, 32, 64


class Options(C.Structure):
    _fields_ = [("continuous", C.c_int32), ("gravity", C.c_float), ("use_wind", C.c_int32), ("wind_power", C.c_float),
                ("turbulence_power", C.c_float), ("wind_idx", C.c_int32), ("torque_idx", C.c_int32),
                ("begin_contact_false", C.c_int32), ("continuous_physics", C.c_int32), ("reverse_seed_order", C.c_int32),
                ("contact_list_head_insertion", C.c_int32), ("det_sincos", C.c_int32), ("force_at_origin", C.c_int32),
                ("canonical_contact_order", C.c_int32)]


_lib = None


def build(force=False):
    srcs = [os.path.join(W2D_DIR, f) for f in ("lunar_sim.cpp", "world2d.hpp", "Makefile")] + [os.path.join(ROOT, "oracle", "detmath.hpp")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", W2D_DIR, "-s"], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.w2d_lunar_default_options.argtypes = [C.POINTER(Options)]
        L.w2d_lunar_create.restype = C.c_void_p
        L.w2d_lunar_create.argtypes = [C.POINTER(Options)]
        L.w2d_lunar_destroy.argtypes = [C.c_void_p]
        L.w2d_lunar_reset.argtypes = [C.c_void_p] * 4
        L.w2d_lunar_step.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
        L.w2d_lunar_export.argtypes = [C.c_void_p] * 3
        L.w2d_lunar_import.argtypes = [C.c_void_p] * 3
        L.w2d_lunar_toi_events.argtypes = [C.c_void_p]
        L.w2d_lunar_toi_events.restype = C.c_int32
        L.w2d_lunar_num_contacts.argtypes = [C.c_void_p]
        L.w2d_lunar_num_contacts.restype = C.c_int32
        L.w2d_lunar_mass_data.argtypes = [C.c_void_p, C.c_void_p]
        L.w2d_lunar_polygon.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.w2d_lunar_polygon.restype = C.c_int32
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


f32 = np.float32


def uniformf(lo, hi, w):
    """oracle/philox.hpp uniformf: lo + (hi - lo) * u01(w), separate float32 multiply and add."""
    u = f32(np.uint32(w) >> np.uint32(8)) * f32(2.0 ** -24)
    return f32(f32(lo) + f32(f32(f32(hi) - f32(lo)) * u))


def reset_draws(seed, gid, ordinal):
    """The 14 uniforms of LunarLanderEnv.Reset from the engine's RESET stream (lunar_core.cuh reset / DESIGN.md RNG spec)."""
    b = [O.draw(seed, gid, ordinal, STREAM_RESET, sub) for sub in range(4)]
    words = [b[0][2], b[0][3], b[1][0], b[1][1], b[1][2], b[1][3], b[2][0], b[2][1], b[2][2], b[2][3], b[3][0], b[3][1]]
    view_h = f32(f32(400.0) / f32(30.0))
    out = [uniformf(-1000.0, 1000.0, b[0][0]), uniformf(-1000.0, 1000.0, b[0][1])]
    out += [uniformf(0.0, f32(view_h / f32(2.0)), w) for w in words]
    return np.array(out, np.float32)


def step_draws(seed, gid, t):
    b = O.draw(seed, gid, t, STREAM_DYNAMICS)
    return np.array([uniformf(-1.0, 1.0, b[0]), uniformf(-1.0, 1.0, b[1])], np.float32)


def ctor_draws(seed, gid):
    b = O.draw(seed, gid, 0, STREAM_CTOR)
    return -9999 + int((int(b[0]) * 19998) >> 32), -9999 + int((int(b[1]) * 19998) >> 32)


class LunarWorld:
    """One LunarLanderEnv instance on the generic engine."""

    def __init__(self, **kw):
        o = Options()
        lib().w2d_lunar_default_options(C.byref(o))
        for k, v in kw.items():
            if not hasattr(o, k):
                raise TypeError("unknown option %r" % k)
            setattr(o, k, v)
        self.opt = o
        self.h = lib().w2d_lunar_create(C.byref(o))
        self.continuous = bool(o.continuous)

    def close(self):
        if self.h:
            lib().w2d_lunar_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self, draws14, zero_step_draws2):
        d = np.ascontiguousarray(draws14, np.float32); z = np.ascontiguousarray(zero_step_draws2, np.float32)
        obs = np.empty(8, np.float32)
        lib().w2d_lunar_reset(self.h, _p(d), _p(z), _p(obs))
        return obs

    def step(self, action, draws2):
        z = np.ascontiguousarray(draws2, np.float32)
        obs = np.empty(8, np.float32)
        r = C.c_float(); dn = C.c_int32()
        if self.continuous:
            a = np.ascontiguousarray(action, np.float32).reshape(2)
            lib().w2d_lunar_step(self.h, 0, _p(a), _p(z), _p(obs), C.byref(r), C.byref(dn))
        else:
            lib().w2d_lunar_step(self.h, int(action), None, _p(z), _p(obs), C.byref(r), C.byref(dn))
        return obs, np.float32(r.value), int(dn.value)

    def export_state(self):
        s = np.empty(STATE_DIM, np.float32); a = np.empty(AUX_DIM, np.int32)
        lib().w2d_lunar_export(self.h, _p(s), _p(a))
        return s, a

    def import_state(self, state, aux):
        s = np.ascontiguousarray(state, np.float32).reshape(-1)[:STATE_DIM].copy()
        a = np.ascontiguousarray(aux, np.int32).reshape(-1)[:AUX_DIM].copy()
        lib().w2d_lunar_import(self.h, _p(s), _p(a))

    def toi_events(self):
        return lib().w2d_lunar_toi_events(self.h)

    def mass_data(self):
        out = np.empty((3, 8), np.float32)
        lib().w2d_lunar_mass_data(self.h, _p(out))
        return out

    def polygon(self, body):
        v = np.zeros((8, 2), np.float32); n = np.zeros((8, 2), np.float32)
        k = lib().w2d_lunar_polygon(self.h, body, _p(v), _p(n))
        return v[:k], n[:k]


def pid_action(s):
    """The reference test's heuristic (tests/Gym.Tests/Envs/Aether/LunarLanderEnvironment.cs:102-150), discrete."""
    angle_targ = min(max(s[0] * 0.5 + s[2] * 1.0, -0.4), 0.4)
    hover_targ = 0.55 * abs(s[0])
    angle_todo = (angle_targ - s[4]) * 0.5 - s[5] * 1.0
    hover_todo = (hover_targ - s[1]) * 0.5 - s[3] * 0.5
    if s[6] > 0 or s[7] > 0:
        angle_todo = 0.0
        hover_todo = -s[3] * 0.5
    if hover_todo > abs(angle_todo) and hover_todo > 0.05:
        return 2
    if angle_todo < -0.05:
        return 3
    if angle_todo > 0.05:
        return 1
    return 0


def generate_transitions(count, seed=2024, T=1000, landers=64, p_pid=(1.0, 1.0, 0.7, 0.3), continuous=False, rng_seed=5, max_episode=600, **opts):
    """`count` single-step transitions of the generic oracle, free-running over `landers` episodes at a time (restarted
    when done) under a mix of the PID heuristic and random actions, so that free flight, touch-down on one and two legs,
    belly contact, resting and falling asleep all occur (lander k follows the heuristic with probability p_pid[k % len]).  Transition i is stepped with the dispersion draws of
    (seed, env id i, step index T): exactly what env i of a batch does after set_state(..., t=T).
    Returns dict of arrays: state0, aux0, action, state1, aux1, obs, reward, done."""
    rng = np.random.default_rng(rng_seed)
    out = {k: [] for k in ("state0", "aux0", "action", "state1", "aux1", "obs", "reward", "done")}
    worlds, obs_now, age = [], [], []
    episode = 0

    def fresh():
        nonlocal episode
        wi, ti = ctor_draws(seed, 1000000 + episode)
        w = LunarWorld(continuous=int(continuous), wind_idx=wi, torque_idx=ti, **opts)
        o = w.reset(reset_draws(seed, 1000000 + episode, 0), step_draws(seed, 1000000 + episode, 0))
        episode += 1
        return w, o

    for _ in range(landers):
        w, o = fresh()
        worlds.append(w); obs_now.append(o); age.append(0)
    i = 0
    while i < count:
        for k in range(landers):
            if i >= count:
                break
            w = worlds[k]
            s0, a0 = w.export_state()
            if continuous:
                act = rng.uniform(-1, 1, 2).astype(np.float32)
                if rng.random() < p_pid[k % len(p_pid)]:   # the heuristic's continuous branch (:128-132)
                    s = obs_now[k]
                    angle_targ = min(max(s[0] * 0.5 + s[2] * 1.0, -0.4), 0.4)
                    hover_targ = 0.55 * abs(s[0])
                    angle_todo = (angle_targ - s[4]) * 0.5 - s[5] * 1.0
                    hover_todo = (hover_targ - s[1]) * 0.5 - s[3] * 0.5
                    if s[6] > 0 or s[7] > 0:
                        angle_todo = 0.0
                        hover_todo = -s[3] * 0.5
                    act = np.clip(np.array([hover_todo * 20 - 1, -angle_todo * 20], np.float32), -1, 1)
            else:
                act = pid_action(obs_now[k]) if rng.random() < p_pid[k % len(p_pid)] else int(rng.integers(0, 4))
            o, r, d = w.step(act, step_draws(seed, i, T))
            s1, a1 = w.export_state()
            out["state0"].append(s0); out["aux0"].append(a0); out["action"].append(act)
            out["state1"].append(s1); out["aux1"].append(a1); out["obs"].append(o); out["reward"].append(r); out["done"].append(d)
            i += 1
            age[k] += 1
            obs_now[k] = o
            if d or age[k] >= max_episode:
                w.close()
                worlds[k], obs_now[k] = fresh()
                age[k] = 0
    for w in worlds:
        w.close()
    res = {k: np.array(v) for k, v in out.items()}
    res["action"] = res["action"].astype(np.float32 if continuous else np.int32)
    res["done"] = res["done"].astype(np.uint8)
    res["reward"] = res["reward"].astype(np.float32)
    return res


def categories(aux0, state0):
    """Coarse class of a pre-step state, for coverage reports: free / near (pairs, not touching) / legs / belly / asleep."""
    touch = aux0[:, 0:3]
    flags = aux0[:, 3]
    free = aux0[:, 26] == -1
    cat = np.full(len(aux0), "near", dtype=object)
    cat[free] = "free"
    legs = (touch[:, 1] != 0) | (touch[:, 2] != 0)
    cat[legs & (touch[:, 0] == 0)] = "legs"
    cat[(touch[:, 1] != 0) & (touch[:, 2] != 0) & (touch[:, 0] == 0)] = "two_legs"
    cat[touch[:, 0] != 0] = "belly"
    cat[(flags & F_AWAKE) == 0] = "asleep"
    return cat


# Single-step sensitivity to the sin/cos implementation.  The generic oracle run with the engine's deterministic float32
# sincos agrees with the engine bit for bit; run with the reference's (float)Math.Sin((double)a) it differs from ITSELF by
# the bounds below, because a one-ulp change of a rotation row is amplified by the conditioning of the revolute-joint rows
# (leg inertia 1.1e-4 against 0.78 for the fuselage: inverse inertias 8934 and 1.28) and by the 2-point block solver's
# split of a leg's load between its two manifold points.  (index groups of the 80-word state, absolute error against
# max(|value|, 1) unless noted)
TOLERANCES = (
    ("fuselage pose", [0, 1, 2], 1e-5),
    ("fuselage velocity", [3, 4, 5], 2e-4),
    ("leg poses", [7, 8, 9, 14, 15, 16], 1e-4),
    ("leg velocities", [10, 11, 12, 17, 18, 19], 1e-2),
    ("sleep timers", [6, 13, 20], 0.0),
    ("joint impulses", list(range(21, 29)), 5e-3),
    ("contact impulses", list(range(29, 53)), 1e-2),
    ("terrain", list(range(53, 64)), 0.0),
    ("shaping, force, torque", list(range(64, 68)), 1e-2),
    ("proxy boxes", list(range(68, 80)), 1e-5),
)
OBS_ATOL, REWARD_ATOL = 2e-4, 1e-2


def compare_transitions(tag, want, got_state, got_aux, got_obs, got_reward, got_done, exact):
    """want: dict from generate_transitions; got_*: the engine's results for the same inputs.  Integer words (contact flags,
    touching masks, limit states, contact ids, pair lists) and done must be identical; floats identical (exact) or within
    TOLERANCES (the generic oracle on the reference's sin/cos)."""
    aux1 = want["aux1"]
    bad_aux = np.argwhere(got_aux[:, :AUX_DIM] != aux1)
    assert len(bad_aux) == 0, "%s: %d int words differ, first (transition, word) %s: got %d want %d" % (
        tag, len(bad_aux), tuple(bad_aux[0]), got_aux[tuple(bad_aux[0])], aux1[tuple(bad_aux[0])])
    assert np.array_equal(np.asarray(got_done).astype(np.uint8), want["done"]), "%s: done flags differ" % tag
    gs = np.asarray(got_state, np.float32)[:, :STATE_DIM]; ws = np.asarray(want["state1"], np.float32)
    go = np.asarray(got_obs, np.float32); wo = np.asarray(want["obs"], np.float32)
    gr = np.asarray(got_reward, np.float32); wr = np.asarray(want["reward"], np.float32)
    if exact:
        for name, g, w in (("state", gs, ws), ("obs", go, wo), ("reward", gr, wr)):
            same = (g == w) | (np.isnan(g) & np.isnan(w))
            assert same.all(), "%s: %s differs in %d values (bit-exact mode), max |diff| %.3e, first at %s" % (
                tag, name, int((~same).sum()), float(np.abs(g.astype(np.float64) - w).max()), tuple(np.argwhere(~same)[0]))
        return
    for name, idx, tol in TOLERANCES:
        g = gs[:, idx].astype(np.float64); w = ws[:, idx].astype(np.float64)
        err = np.abs(g - w) / np.maximum(np.abs(w), 1.0)
        assert err.max() <= tol, "%s: %s: error %.3e > %.1e at transition %d" % (tag, name, float(err.max()), tol, int(err.max(axis=1).argmax()))
    assert np.abs(go.astype(np.float64) - wo).max() <= OBS_ATOL, "%s: observation error %.3e" % (tag, float(np.abs(go.astype(np.float64) - wo).max()))
    big = want["done"] != 0   # the terminal +-100 are exact
    assert np.array_equal(gr[big], wr[big])
    assert np.abs(gr.astype(np.float64) - wr)[~big].max() <= REWARD_ATOL, "%s: reward error %.3e" % (tag, float(np.abs(gr.astype(np.float64) - wr)[~big].max()))
