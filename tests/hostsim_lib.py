"""TEST INFRASTRUCTURE: builds tests/hostsim/hostsim.cpp (the engine's device headers compiled for the host) and wraps
its entry points.  Used by test_hostsim_cpu.py (one thread at a time) and test_hostsim_simt_cpu.py (CTAs of fibers)."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS_DIR = os.path.join(ROOT, "tests", "hostsim")
CSRC = os.path.join(ROOT, "gym.net_b200", "csrc")
LIB = os.path.join(HS_DIR, "_hostsim.so")


def build():
    srcs = [os.path.join(HS_DIR, "hostsim.cpp"), os.path.join(HS_DIR, "stubs", "cuda_runtime.h"), os.path.join(HS_DIR, "simt.hpp")] + [
        os.path.join(CSRC, f) for f in ("detmath.cuh", "philox.cuh", "env_classic.cuh", "lunar.cuh", "lunar_core.cuh", "kernels.cuh", "normalize.cuh")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-w", "-I" + os.path.join(HS_DIR, "stubs"),
                        "-I" + CSRC, "-shared", "-fPIC", "-o", LIB, srcs[0]], check=True)
    L = C.CDLL(LIB)
    V, I, U32, U64, F = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float
    L.hostsim_step.argtypes = [I, V, V, V, V, V, V, V, V, V, I, U32, U64, U64, I, I, F, F, F, I]
    L.hostsim_reset.argtypes = [I, V, V, V, V, V, V, I, U32, U64, U64, F, F, F, I]
    L.hostsim_ctor.argtypes = [I, V, V, I, U32, U64]
    L.hostsim_rollout.argtypes = [I, V, V, V, V, V, V, V, V, V, V, V, V, I, I, I, U32, U64, U64, I, I, I, I, F, F, F, I]
    L.hostsim_set_simt.argtypes = [I]
    L.hostsim_set_trio.argtypes = [I]
    L.hostsim_set_terminal_obs.argtypes = [V]
    L.hostsim_set_actions_in.argtypes = [V, V]
    L.hostsim_set_seeds.argtypes = [V]
    L.hostsim_set_schedule.argtypes = [I, U64]
    L.hostsim_step_kernel.argtypes = [I, V, V, V, V, V, V, V, V, V, V, V, V, V, V, V, I, V, I, U32, U64, U64, I, I, I, C.c_int32, U32, F, F, F, I]
    L.hostsim_reset_kernel.argtypes = [I, V, V, V, V, V, V, V, V, I, U32, U64, U64, F, F, F, I]
    L.hostsim_sample_kernel.argtypes = [I, V, V, I, U32, U64, U64]
    L.hostsim_partition.argtypes = [V, I, V, V]
    L.hostsim_normalize.argtypes = [V, V, V, V, V, I, I, F, F, F, F, I]
    L.hostsim_set_gather.argtypes = [I, I, U32, V, V, V]
    L.hostsim_gather_wait.argtypes = [V, I, U32]
    L.hostsim_div_inrange.argtypes = [V, V, V, C.c_size_t]
    L.hostsim_sincos.argtypes = [V, V, V, C.c_size_t]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


DEFAULT_LIMIT = {O.CARTPOLE: 0, O.PENDULUM: 200, O.MOUNTAINCAR: 200, O.MOUNTAINCAR_CONT: 999, O.ACROBOT: 500,
                 O.LUNARLANDER: 0, O.LUNARLANDER_CONT: 0}
PRM = (-10.0, 15.0, 1.5, 0)   # gravity, wind_power, turbulence_power, enable_wind (LunarLanderEnv.cs:351-354)


class HostSim:
    """Device-layout buffers + the replayed kernel bodies."""

    def __init__(self, L, kind, n, seed, env_off=0, auto_reset=True, prm=PRM):
        self.L, self.kind, self.n, self.seed, self.off, self.auto = L, kind, n, seed, env_off, auto_reset
        self.prm = tuple(prm)          # gravity, wind_power, turbulence_power, enable_wind
        d = O.dims(kind)
        self.sd, self.od, self.ad, self.actn = d["state_dim"], d["obs_dim"], d["act_dim"], d["act_n"]
        self.lunar = kind >= O.LUNARLANDER
        self.auxw = d["aux_dim"] - 2 if self.lunar else 0
        self.state = np.zeros((self.sd, n) if self.lunar else (n, self.sd), np.float32)
        self.aux = np.zeros((max(self.auxw, 1), n), np.int32)
        self.sbd = np.full(n, -1, np.int32); self.ept = np.zeros(n, np.int32); self.episode = np.zeros(n, np.int32)
        self.limit = DEFAULT_LIMIT[kind]
        self.t = 0
        L.hostsim_ctor(kind, _p(self.state), _p(self.aux), n, env_off, seed)

    def reset(self):
        obs = np.empty((self.n, self.od), np.float32)
        self.L.hostsim_reset(self.kind, _p(self.state), _p(self.aux), _p(self.sbd), _p(self.ept), _p(self.episode), _p(obs),
                             self.n, self.off, self.seed, self.t, *self.prm)
        return obs

    def step(self, actions):
        a = np.ascontiguousarray(actions)
        obs = np.empty((self.n, self.od), np.float32); rew = np.empty(self.n, np.float32); done = np.empty(self.n, np.uint8)
        bad = self.L.hostsim_step(self.kind, _p(self.state), _p(self.aux), _p(self.sbd), _p(self.ept), _p(self.episode), _p(a),
                                  _p(obs), _p(rew), _p(done), self.n, self.off, self.seed, self.t, self.limit, int(self.auto), *self.prm)
        self.t += 1
        return obs, rew, done, bad

    def rollout(self, k, all_out=True, block=64, ep_ret=None, sums=None, done_bits=0, actions_in=None):
        """actions_in [k][n](, ad): gymcuda_step_many* (all_out False: the generic variant fed with the caller's actions; all_out
        True: the chunked SUPPLIED variant); self.rollout_invalid then holds [host flag, rejected count]."""
        n = self.n
        obs = np.empty((k, n, self.od), np.float32); rew = np.empty((k, n), np.float32); done = np.empty((k, n), np.uint8)
        act = np.empty((k, n), np.int32) if self.actn > 0 else np.empty((k, n, self.ad), np.float32)
        stats = np.zeros(2, np.uint64)
        if actions_in is not None:
            a_in = np.ascontiguousarray(actions_in); flag = np.zeros(1, np.int32)
            self.L.hostsim_set_actions_in(_p(a_in), _p(flag))
            act = None
        rc = self.L.hostsim_rollout(self.kind, _p(self.state), _p(self.aux), _p(self.sbd), _p(self.ept), _p(self.episode), _p(obs),
                                    _p(rew), _p(done), None if act is None else _p(act), _p(stats), None if ep_ret is None else _p(ep_ret),
                                    None if sums is None else _p(sums), done_bits, n, k, self.off, self.seed, self.t, self.limit,
                                    int(self.auto), int(all_out), block, *self.prm)
        assert rc == 0, rc
        self.t += k
        if actions_in is not None:
            self.rollout_invalid = (int(flag[0]), int(stats[1]))
        return obs, rew, done, act, int(stats[0])

    # ---- the kernels themselves (run them under hostsim_set_simt(1)) ------------------------------------------------
    def step_kernel(self, actions, *, perm=None, ep_ret=None, sums=None, done_bits=0, bcast=None, done_offset=0, terminal_obs=None):
        """step_kernel as gymcuda_step_device launches it.  Returns obs, reward, done, done_idx[:count], invalid flag.
        self.stats ([episodes, invalid]) and self.done_count ([2], by parity of the launch number) persist like on the device."""
        n = self.n
        if not hasattr(self, "stats"):
            self.stats = np.zeros(2, np.uint64); self.done_count = np.zeros(2, np.int32); self.seq = 0
        a = np.ascontiguousarray(actions) if actions is not None else None
        obs = np.empty((n, self.od), np.float32); rew = np.empty(n, np.float32)
        done_buf = np.zeros(n + 8, np.uint8); done = done_buf[done_offset:done_offset + n]   # offset != 0 mod 4: unpacked done stores
        idx = np.full(n, -1, np.int32); flag = np.zeros(2, np.int32)
        if terminal_obs is not None:
            self.L.hostsim_set_terminal_obs(_p(terminal_obs))
        rc = self.L.hostsim_step_kernel(self.kind, _p(self.state), _p(self.aux), _p(self.sbd), _p(self.ept), _p(self.episode),
                                        None if perm is None else _p(perm), None if a is None else _p(a), _p(obs), _p(rew),
                                        C.c_void_p(done_buf.ctypes.data + done_offset), _p(idx), _p(self.done_count), _p(self.stats),
                                        None if ep_ret is None else _p(ep_ret), None if sums is None else _p(sums), done_bits, _p(flag),
                                        n, self.off, self.seed, self.t, self.limit, int(self.auto), int(bcast is not None),
                                        0 if bcast is None else int(bcast), self.seq, *self.prm)
        assert rc == 0, rc
        count = int(self.done_count[self.seq & 1])
        assert self.done_count[(self.seq + 1) & 1] == 0          # the kernel zeroes the next launch's counter
        self.seq += 1
        self.t += 1
        return obs, rew, done.copy(), idx[:count].copy(), int(flag[0])

    def reset_kernel(self, mask=None, ep_ret=None):
        obs = np.empty((self.n, self.od), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        rc = self.L.hostsim_reset_kernel(self.kind, _p(self.state), _p(self.aux), _p(self.sbd), _p(self.ept), _p(self.episode),
                                         None if m is None else _p(m), _p(obs), None if ep_ret is None else _p(ep_ret), self.n,
                                         self.off, self.seed, self.t, *self.prm)
        assert rc == 0, rc
        return obs

    def sample_kernel(self, mask=None):
        out = np.empty(self.n, np.int32) if self.actn > 0 else np.empty((self.n, self.ad), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        rc = self.L.hostsim_sample_kernel(self.kind, None if m is None else _p(m), _p(out), self.n, self.off, self.seed, self.t)
        assert rc == 0, rc
        return out

    def partition(self):
        """contact_partition() of gymcuda.cu: the thread -> env map of the LunarLander kernels."""
        nb = (self.n + 255) // 256
        block_free = np.zeros(nb + 1, np.int32); perm = np.full(self.n, -1, np.int32)
        assert self.L.hostsim_partition(_p(self.aux), self.n, _p(block_free), _p(perm)) == nb
        return perm, block_free

    def seed_each(self, seeds):
        """gymcuda_seed_each: per-env seeds, generators restarted (episode ordinals and t back to 0, constructor draws redone)."""
        self._seeds = np.ascontiguousarray(seeds, np.int32)
        assert self._seeds.shape == (self.n,)
        self.L.hostsim_set_seeds(_p(self._seeds))
        self.episode[:] = 0
        self.t = 0
        self.L.hostsim_ctor(self.kind, _p(self.state), _p(self.aux), self.n, self.off, self.seed)

    def unseed(self):
        self.L.hostsim_set_seeds(None)

    def abi_state(self):
        """[n][state_dim] like gymcuda_get_state."""
        return self.state.T.copy() if self.lunar else self.state.copy()




class NormalizeModel:
    """numpy restatement of gymcuda_normalize (include/gymcuda.h; the VecNormalize recipe): the checker of both the host
    build of normalize.cuh (tests/test_hostsim_simt_cpu.py) and the CUDA build (tests/test_gpu_classic.py)."""

    def __init__(self, n, od, gamma=0.99, eps=1e-8, clip_obs=10.0, clip_reward=10.0):
        self.n, self.od = n, od
        self.gamma, self.eps, self.clip_obs, self.clip_reward = np.float32(gamma), np.float32(eps), np.float32(clip_obs), np.float32(clip_reward)
        self.s = np.zeros(od); self.q = np.zeros(od); self.sr = 0.0; self.qr = 0.0; self.count = 0.0; self.count_ret = 0.0   # separate counts: obs-only / reward-only calls
        self.ret = np.zeros(n, np.float32)

    def __call__(self, obs, reward, done, update=True):
        """Returns normalised copies (float32) of obs / reward (None stays None)."""
        if update:
            if obs is not None:
                x = obs.astype(np.float64)
                self.s += x.sum(axis=0); self.q += (x * x).sum(axis=0)
                self.count += self.n
            if reward is not None:
                r = (self.ret * self.gamma).astype(np.float32) + reward          # float32 multiply, then float32 add
                r64 = r.astype(np.float64)
                self.sr += r64.sum(); self.qr += (r64 * r64).sum()
                self.ret = np.where(done != 0, np.float32(0), r).astype(np.float32) if done is not None else r
                self.count_ret += self.n
        out_o, out_r = obs, reward
        if obs is not None and self.count > 0:
            mean = self.s / self.count
            var = np.maximum(self.q / self.count - mean * mean, 0.0)
            out_o = np.clip(((obs.astype(np.float64) - mean) / np.sqrt(var + np.float64(self.eps))).astype(np.float32), -self.clip_obs, self.clip_obs)
        if reward is not None and self.count_ret > 0:
            mean = self.sr / self.count_ret
            var = max(self.qr / self.count_ret - mean * mean, 0.0)
            out_r = np.clip((reward.astype(np.float64) / np.sqrt(var + np.float64(self.eps))).astype(np.float32), -self.clip_reward, self.clip_reward)
        return out_o, out_r
