"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

CARTPOLE, PENDULUM, MOUNTAINCAR, MOUNTAINCAR_CONT, ACROBOT, LUNARLANDER, LUNARLANDER_CONT = range(7)
MODE_F64, MODE_F64_F32STORE, MODE_F32 = 0, 1, 2
FLAG_AUTO_RESET = 1

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB_PATH)
        for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".hpp", ".h", "Makefile"))
    ):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_dims.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 5
        L.oracle_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_seed_each.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_lunar_params.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_float, C.c_float]
        L.oracle_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_reset_masked.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_step.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.oracle_step.restype = C.c_int
        L.oracle_step_many.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_step_many.restype = C.c_int
        L.oracle_sample_actions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_rollout_random.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.oracle_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.oracle_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.oracle_philox4x32_10.argtypes = [C.c_void_p] * 3
        L.oracle_draw.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        L.oracle_sincosf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def dims(kind):
    v = [C.c_int() for _ in range(5)]
    if lib().oracle_dims(kind, *[C.byref(x) for x in v]) != 0:
        raise ValueError("unknown env kind %r" % kind)
    return dict(zip(("state_dim", "aux_dim", "obs_dim", "act_dim", "act_n"), (x.value for x in v)))


class OracleEnv:
    """Serial per-instance loop over `n` scalar envs: the reference's VecEnvWrapper shape."""

    def __init__(self, kind, n, seed=0, env_id_offset=0, auto_reset=False, time_limit=0, mode=MODE_F64_F32STORE,
                 done_bits=False):
        self.kind, self.n, self.mode = kind, n, mode
        self.d = dims(kind)
        self.h = lib().oracle_create(kind, n, seed, env_id_offset,
                                     (FLAG_AUTO_RESET if auto_reset else 0) | (4 if done_bits else 0), time_limit, mode)
        if not self.h:
            raise ValueError("oracle_create failed")
        self.discrete = self.d["act_n"] > 0

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def seed(self, seed):
        lib().oracle_seed(self.h, seed)

    def seed_each(self, seeds):
        s = np.ascontiguousarray(seeds, dtype=np.int32)
        assert s.shape == (self.n,)
        lib().oracle_seed_each(self.h, _p(s))

    def set_threads(self, k):
        lib().oracle_set_threads(self.h, k)

    def reset(self, mask=None):
        obs = np.empty((self.n, self.d["obs_dim"]), np.float32)
        if mask is None:
            lib().oracle_reset(self.h, _p(obs))
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            lib().oracle_reset_masked(self.h, _p(m), _p(obs))
        return obs

    def _actions(self, actions):
        if self.discrete:
            a = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n)
        else:
            a = np.ascontiguousarray(actions, dtype=np.float32).reshape(self.n, self.d["act_dim"])
        return a

    def step(self, actions):
        a = self._actions(actions)
        obs = np.empty((self.n, self.d["obs_dim"]), np.float32)
        rew = np.empty(self.n, np.float32)
        done = np.empty(self.n, np.uint8)
        bad = lib().oracle_step(self.h, _p(a), _p(obs), _p(rew), _p(done))
        self.invalid = bad
        return obs, rew, done

    def step_many(self, actions):
        """actions [K][n](,act_dim): K steps, no outputs (timing loop)."""
        a = np.ascontiguousarray(actions, dtype=np.int32 if self.discrete else np.float32)
        return lib().oracle_step_many(self.h, a.shape[0], _p(a))

    def sample_actions(self, mask=None):
        out = np.empty(self.n, np.int32) if self.discrete else np.empty((self.n, self.d["act_dim"]), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8).reshape(self.n, self.d["act_n"])
        lib().oracle_sample_actions(self.h, _p(m), _p(out))
        return out

    def rollout_random(self, k, want_obs=True):
        obs = np.empty((k, self.n, self.d["obs_dim"]), np.float32) if want_obs else None
        rew = np.empty((k, self.n), np.float32)
        done = np.empty((k, self.n), np.uint8)
        act = (np.empty((k, self.n), np.int32) if self.discrete
               else np.empty((k, self.n, self.d["act_dim"]), np.float32))
        lib().oracle_rollout_random(self.h, k, _p(obs), _p(rew), _p(done), _p(act))
        return obs, rew, done, act

    def get_state(self):
        st = np.empty((self.n, self.d["state_dim"]), np.float64)
        aux = np.empty((self.n, self.d["aux_dim"]), np.int32)
        t = C.c_uint64()
        lib().oracle_get_state(self.h, _p(st), _p(aux), C.byref(t))
        return st, aux, t.value

    def set_state(self, state, aux, t):
        st = np.ascontiguousarray(state, dtype=np.float64).reshape(self.n, self.d["state_dim"])
        ax = np.ascontiguousarray(aux, dtype=np.int32).reshape(self.n, self.d["aux_dim"])
        lib().oracle_set_state(self.h, _p(st), _p(ax), t)


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.empty(4, np.uint32)
    lib().oracle_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def draw(seed, env_id, index, stream, sub=0):
    o = np.empty(4, np.uint32)
    lib().oracle_draw(seed, env_id, index, stream, sub, _p(o))
    return o


def sincosf(x):
    x = np.ascontiguousarray(x, np.float32)
    s = np.empty_like(x); c = np.empty_like(x)
    lib().oracle_sincosf(_p(x), _p(s), _p(c), x.size)
    return s, c
