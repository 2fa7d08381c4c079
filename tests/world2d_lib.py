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
