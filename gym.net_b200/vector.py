"""Gym.Environments.Vector -- the host-side mirror of the reference's Env / VecEnv surface for the
batched CUDA path.  Python twin of csharp/Gym.Environments.Vector/CudaVecEnv.cs; both sit on the
same C ABI (include/gymcuda.h).

Reference members mirrored (paths relative to the reference):
  Step {Observation, Reward, Done, Information}   src/Gym/Observations/Step.cs:7-20
  VecEnv ctor / Reset / Step(int) / Seed / Close  src/Gym/Envs/VecEnv.cs:12-53, IVecEnv.cs:8-19
  VecEnvWrapper.Step broadcast of one action      src/Gym/Envs/VecEnvWrapper.cs:22-24
and extended with what a real VectorEnv needs (per-env actions, auto-reset, fused rollouts).
"""
import ctypes as C

import numpy as np

from . import _native as N
from .spaces import Box, Discrete


class Step:
    """Gym.Observations.Step (Step.cs:7-29)."""

    __slots__ = ("Observation", "Reward", "Done", "Information")

    def __init__(self, observation=None, reward=0.0, done=False, information=None):
        self.Observation, self.Reward, self.Done, self.Information = observation, reward, done, information

    def __iter__(self):   # Deconstruct (Step.cs:22-27): var (obs, reward, done, info) = env.Step(a)
        return iter((self.Observation, self.Reward, self.Done, self.Information))

    def __repr__(self):
        return "Reward: %s, Done: %s, Information: %s, Observation: %s" % (
            self.Reward, self.Done, self.Information, self.Observation)


def _widen_seed(seed):
    """int -> the 64-bit engine seed, by the rule of the C# shim ((ulong)(uint) seed for the reference's 32-bit Seed(int)): a
    negative seed is its 32-bit two's-complement pattern, zero-extended; non-negative seeds up to 2^64 - 1 pass through."""
    seed = int(seed)
    return seed & 0xFFFFFFFF if seed < 0 else seed & (2**64 - 1)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CudaVecEnv:
    """VecEnv over `num_envs` instances of one env family living on one GPU.

    Source-compatible members: Reset() -> NDArray[], Step(int action) -> Step[] (broadcast),
    Seed(int), Seed(int[]), Close(), ActionSpace, ObservationSpace, NumberOfEnvironments.
    Batched members: ResetBatch(), StepBatch(actions) -> (obs, reward, done), RolloutRandom(k).
    """

    ENV_KIND = None

    def __init__(self, num_envs, seed=0, device=0, env_id_offset=0, auto_reset=False, time_limit=0,
                 episode_stats=False, done_bits=False, **params):
        if self.ENV_KIND is None:
            raise TypeError("use a concrete family: CartPoleVecEnv, PendulumVecEnv, ...")
        L = N.lib()
        cfg = N.Config()
        N.check(L.gymcuda_config_default(C.byref(cfg), self.ENV_KIND, int(num_envs)))
        cfg.device, cfg.seed, cfg.env_id_offset = int(device), _widen_seed(seed), int(env_id_offset)
        cfg.flags = ((N.FLAG_AUTO_RESET if auto_reset else 0) | (N.FLAG_EPISODE_STATS if episode_stats else 0)
                     | (N.FLAG_DONE_BITS if done_bits else 0))
        cfg.time_limit = int(time_limit)
        for k, v in params.items():   # gravity, enable_wind, wind_power, turbulence_power (LunarLanderEnv ctor)
            if not hasattr(cfg, k):
                raise TypeError("unknown env parameter %r" % k)
            setattr(cfg, k, v)
        h = C.c_void_p()
        N.check(L.gymcuda_create(C.byref(cfg), C.byref(h)))
        self._h, self._L = h, L
        info = N.SpaceInfo()
        N.check(L.gymcuda_space(h, C.byref(info)))
        self.info = info
        self.NumberOfEnvironments = int(num_envs)
        self.obs_dim, self.act_dim, self.act_n = info.obs_dim, info.act_dim, info.act_n
        self.state_dim, self.aux_dim, self.TimeLimit = info.state_dim, info.aux_dim, info.time_limit
        low = np.array(info.obs_low[:self.obs_dim], np.float32)
        high = np.array(info.obs_high[:self.obs_dim], np.float32)
        self.ObservationSpace = Box(low, high, dtype=np.float32)
        if self.act_n > 0:
            self.ActionSpace = Discrete(self.act_n)
        else:
            self.ActionSpace = Box(np.array(info.act_low[:self.act_dim], np.float32),
                                   np.array(info.act_high[:self.act_dim], np.float32), dtype=np.float32)
        self.Metadata = {"render.modes": [], "video.frames_per_second": 50}
        self.RewardRange = (-float("inf"), float("inf"))

    # ---- lifecycle -------------------------------------------------------------------------------
    def Close(self):
        if getattr(self, "_h", None):
            self._L.gymcuda_destroy(self._h)
            self._h = None

    CloseEnvironment = Close

    def __del__(self):
        try:
            self.Close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.Close()

    # ---- Seed ------------------------------------------------------------------------------------
    def Seed(self, seed):
        """Seed(int) (VecEnv.cs:44-46) or Seed(int[]) (VecEnv.cs:48-53)."""
        if np.isscalar(seed):
            N.check(self._L.gymcuda_seed(self._h, _widen_seed(seed)))
        else:
            s = np.ascontiguousarray(seed, dtype=np.int32)
            N.check(self._L.gymcuda_seed_each(self._h, _ptr(s), int(s.size)))

    # ---- batched API -----------------------------------------------------------------------------
    def _out(self):
        n = self.NumberOfEnvironments
        return (np.empty((n, self.obs_dim), np.float32), np.empty(n, np.float32), np.empty(n, np.uint8))

    def ResetBatch(self, mask=None):
        obs = np.empty((self.NumberOfEnvironments, self.obs_dim), np.float32)
        if mask is None:
            N.check(self._L.gymcuda_reset(self._h, _ptr(obs)))
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            if m.shape != (self.NumberOfEnvironments,):
                raise ValueError("mask must have one entry per environment")
            N.check(self._L.gymcuda_reset_masked(self._h, _ptr(m), _ptr(obs)))
        return obs

    def _actions(self, actions):
        n = self.NumberOfEnvironments
        if self.act_n > 0:
            a = np.ascontiguousarray(actions, dtype=np.int32)
            if a.shape != (n,):
                raise ValueError("expected %d discrete actions" % n)
        else:
            a = np.ascontiguousarray(actions, dtype=np.float32).reshape(n, self.act_dim)
        return a

    def StepBatch(self, actions):
        a = self._actions(actions)
        obs, rew, done = self._out()
        N.check(self._L.gymcuda_step(self._h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(done)))
        return obs, rew, done

    def StepMany(self, actions, want=("obs", "reward", "done")):
        """k env steps with caller-supplied actions [k][n](, act_dim) in one launch (gymcuda_step_many)."""
        n = self.NumberOfEnvironments
        if self.act_n > 0:
            a = np.ascontiguousarray(actions, dtype=np.int32)
            if a.ndim != 2 or a.shape[1] != n:
                raise ValueError("expected actions of shape [k][%d]" % n)
        else:
            a = np.ascontiguousarray(actions, dtype=np.float32)
            a = a.reshape(a.shape[0], n, self.act_dim)
        k = a.shape[0]
        obs = np.empty((k, n, self.obs_dim), np.float32) if "obs" in want else None
        rew = np.empty((k, n), np.float32) if "reward" in want else None
        done = np.empty((k, n), np.uint8) if "done" in want else None
        N.check(self._L.gymcuda_step_many(self._h, k, _ptr(a), _ptr(obs), _ptr(rew), _ptr(done)))
        return obs, rew, done

    def StepManyDevice(self, k_steps, d_actions, d_obs=0, d_reward=0, d_done=0):
        N.check(self._L.gymcuda_step_many_device(self._h, int(k_steps), C.c_void_p(d_actions), C.c_void_p(d_obs or None),
                                                 C.c_void_p(d_reward or None), C.c_void_p(d_done or None)))

    def SetTerminalObs(self, buffer):
        """Under auto-reset: `buffer` ([n][obs_dim] float32 numpy array, a raw device pointer, or None to turn the side
        buffer off) receives the observation of the TERMINAL state of every env whose step returns done (the step itself
        returns the post-reset observation).  The array must stay alive while it is registered."""
        if buffer is None:
            self._terminal = None
            N.check(self._L.gymcuda_set_terminal_obs(self._h, None))
        elif isinstance(buffer, int):
            self._terminal = None
            N.check(self._L.gymcuda_set_terminal_obs(self._h, C.c_void_p(buffer)))
        else:
            if buffer.dtype != np.float32 or buffer.shape != (self.NumberOfEnvironments, self.obs_dim) or not buffer.flags.c_contiguous:
                raise ValueError("terminal-observation buffer must be a C-contiguous float32 [num_envs][obs_dim] array")
            self._terminal = buffer
            N.check(self._L.gymcuda_set_terminal_obs(self._h, _ptr(buffer)))

    def RolloutRandom(self, k_steps, want=("obs", "reward", "done", "actions")):
        n, k = self.NumberOfEnvironments, int(k_steps)
        obs = np.empty((k, n, self.obs_dim), np.float32) if "obs" in want else None
        rew = np.empty((k, n), np.float32) if "reward" in want else None
        done = np.empty((k, n), np.uint8) if "done" in want else None
        act = None
        if "actions" in want:
            act = np.empty((k, n), np.int32) if self.act_n > 0 else np.empty((k, n, self.act_dim), np.float32)
        N.check(self._L.gymcuda_rollout_random(self._h, k, _ptr(obs), _ptr(rew), _ptr(done), _ptr(act)))
        return obs, rew, done, act

    def SampleActions(self, mask=None):
        """ActionSpace.Sample() for every env on the device (Discrete.Sample(mask) / Box.Sample())."""
        n = self.NumberOfEnvironments
        out = np.empty(n, np.int32) if self.act_n > 0 else np.empty((n, self.act_dim), np.float32)
        m = None
        if mask is not None:
            if self.act_n == 0:
                raise NotImplementedError("Box.sample cannot be provided a mask.")
            m = np.ascontiguousarray(mask, dtype=np.uint8).reshape(n, self.act_n)
        N.check(self._L.gymcuda_sample_actions(self._h, _ptr(m), _ptr(out)))
        return out

    def DoneIndices(self):
        cnt = C.c_int32()
        idx = np.empty(self.NumberOfEnvironments, np.int32)
        N.check(self._L.gymcuda_done_indices(self._h, _ptr(idx), C.byref(cnt)))
        return idx[:cnt.value].copy()

    def GetState(self):
        n = self.NumberOfEnvironments
        st = np.empty((n, self.state_dim), np.float32)
        aux = np.empty((n, self.aux_dim), np.int32)
        t = C.c_uint64()
        N.check(self._L.gymcuda_get_state(self._h, _ptr(st), _ptr(aux), C.byref(t)))
        return st, aux, t.value

    def SetState(self, state, aux, t):
        n = self.NumberOfEnvironments
        st = np.ascontiguousarray(state, dtype=np.float32).reshape(n, self.state_dim)
        ax = np.ascontiguousarray(aux, dtype=np.int32).reshape(n, self.aux_dim)
        N.check(self._L.gymcuda_set_state(self._h, _ptr(st), _ptr(ax), int(t)))

    def Observe(self):
        obs = np.empty((self.NumberOfEnvironments, self.obs_dim), np.float32)
        N.check(self._L.gymcuda_observe(self._h, _ptr(obs)))
        return obs

    def Render(self, env_ids=None, width=600, height=400, count=None):
        """Env.Render(mode: "rgb_array") for a subset of the batch, rasterised on the device: uint8 [count, height, width, 3]
        (CartPole and LunarLander, the two envs the reference can render)."""
        if env_ids is None:
            k = int(count if count is not None else min(self.NumberOfEnvironments, 1))
            ids = None
        else:
            ids = np.ascontiguousarray(env_ids, dtype=np.int32); k = int(ids.size)
        out = np.empty((k, int(height), int(width), 3), np.uint8)
        N.check(self._L.gymcuda_render(self._h, _ptr(ids), k, int(width), int(height), _ptr(out)))
        return out

    def Stats(self, reset=False):
        s = N.Stats()
        N.check(self._L.gymcuda_get_stats(self._h, C.byref(s), 1 if reset else 0))
        return {"env_steps": s.env_steps, "episodes": s.episodes, "invalid_actions": s.invalid_actions,
                "return_sum": s.return_sum, "length_sum": s.length_sum}

    # ---- observation / reward normalisation (the VecNormalize recipe, on device) -----------------------
    def NormalizeConfig(self, gamma=0.99, epsilon=1e-8, clip_obs=10.0, clip_reward=10.0):
        N.check(self._L.gymcuda_normalize_config(self._h, gamma, epsilon, clip_obs, clip_reward))

    def Normalize(self, obs=None, reward=None, done=None, update=True):
        """In place on float32 host arrays: obs <- clip((obs - mean) / sqrt(var + eps)), reward <- clip(reward /
        sqrt(var of the discounted return + eps)); update=True adds this batch to the running statistics first."""
        n = self.NumberOfEnvironments
        for name, a, shape, dt in (("obs", obs, (n, self.obs_dim), np.float32), ("reward", reward, (n,), np.float32),
                                   ("done", done, (n,), np.uint8)):
            if a is not None and not (isinstance(a, np.ndarray) and a.dtype == dt and a.shape == shape and a.flags.c_contiguous):
                raise ValueError("%s must be a C-contiguous %s array of shape %s (it is normalised in place)" % (name, np.dtype(dt).name, shape))
        N.check(self._L.gymcuda_normalize(self._h, _ptr(obs), _ptr(reward), _ptr(done), 1 if update else 0))
        return obs, reward

    def NormalizeDevice(self, d_obs=0, d_reward=0, d_done=0, update=True):
        N.check(self._L.gymcuda_normalize_device(self._h, C.c_void_p(d_obs or 0), C.c_void_p(d_reward or 0),
                                                 C.c_void_p(d_done or 0), 1 if update else 0))

    def NormalizeStats(self):
        mean = np.zeros(self.obs_dim, np.float64); var = np.zeros(self.obs_dim, np.float64)
        rv = C.c_double(); cnt = C.c_double()
        N.check(self._L.gymcuda_normalize_get(self._h, _ptr(mean), _ptr(var), C.byref(rv), C.byref(cnt)))
        return {"obs_mean": mean, "obs_var": var, "return_var": rv.value, "count": cnt.value}

    def NormalizeReset(self):
        N.check(self._L.gymcuda_normalize_reset(self._h))

    # ---- device-pointer API (torch tensors / raw pointers) ------------------------------------------
    def SetDeviceClock(self, on=True):
        """Step index and launch sequence number on the device: makes a captured StepDevice replayable (CUDA graphs)."""
        N.check(self._L.gymcuda_set_device_clock(self._h, 1 if on else 0))

    def SetStream(self, cuda_stream):
        N.check(self._L.gymcuda_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def Sync(self):
        N.check(self._L.gymcuda_sync(self._h))

    def StepDevice(self, d_actions, d_obs, d_reward, d_done):
        N.check(self._L.gymcuda_step_device(self._h, C.c_void_p(d_actions), C.c_void_p(d_obs or 0),
                                            C.c_void_p(d_reward or 0), C.c_void_p(d_done or 0)))

    NO_OBS = N.NO_OBS   # StepDevice(d_obs=NO_OBS): no observation copy (read them through ObsViewDevice, or call Observe)

    def ObsViewDevice(self):
        """Device pointer to [n][obs_dim] float32 = the current observations without a copy (the library's state array); only
        for the env kinds whose observation is their state vector (CartPole, MountainCar, MountainCarContinuous)."""
        p = C.c_void_p()
        N.check(self._L.gymcuda_obs_view_device(self._h, C.byref(p)))
        return p.value

    def RolloutRandomDevice(self, k_steps, d_obs=0, d_reward=0, d_done=0, d_actions=0):
        N.check(self._L.gymcuda_rollout_random_device(self._h, int(k_steps), C.c_void_p(d_obs or 0),
                                                      C.c_void_p(d_reward or 0), C.c_void_p(d_done or 0),
                                                      C.c_void_p(d_actions or 0)))

    def CommInit(self, unique_id, rank, world_size):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        N.check(self._L.gymcuda_comm_init(self._h, buf, int(rank), int(world_size)))

    def AllGatherObs(self, d_out, d_obs=0):
        N.check(self._L.gymcuda_allgather_obs(self._h, C.c_void_p(d_obs or 0), C.c_void_p(d_out)))

    # ---- fused step + obs gather over NVLink peer memory -----------------------------------------------
    def GatherCreate(self, rank, world_size):
        buf = (C.c_uint8 * 64)()
        N.check(self._L.gymcuda_gather_create(self._h, int(rank), int(world_size), buf))
        return bytes(buf)

    def GatherOpen(self, handles):
        blob = b"".join(bytes(h) for h in handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        N.check(self._L.gymcuda_gather_open(self._h, buf))

    def StepGatherDevice(self, d_actions, d_reward=0, d_done=0):
        """Step + all-gather in one kernel; returns the device pointer of this rank's [world][n][obs_dim] view."""
        out = C.c_void_p()
        N.check(self._L.gymcuda_step_gather_device(self._h, C.c_void_p(d_actions), C.c_void_p(d_reward or 0),
                                                   C.c_void_p(d_done or 0), C.byref(out)))
        return out.value

    def GatherWait(self):
        N.check(self._L.gymcuda_gather_wait(self._h))

    # ---- source-compatible IVecEnv members -----------------------------------------------------------
    def Reset(self):
        """IVecEnv.Reset() -> NDArray[] (one observation array per env)."""
        return list(self.ResetBatch())

    def Step(self, action):
        """IVecEnv.Step(int action) -> Step[]: ONE action broadcast to every env (VecEnvWrapper.cs:22-24)."""
        if self.act_n == 0:
            raise NotImplementedError("IVecEnv.Step(int) needs a Discrete action space; use StepBatch")
        obs, rew, done = self._out()
        N.check(self._L.gymcuda_step_broadcast(self._h, int(action), _ptr(obs), _ptr(rew), _ptr(done)))
        return [Step(obs[i], float(rew[i]), bool(done[i]), self._information(obs[i])) for i in range(self.NumberOfEnvironments)]

    def _information(self, obs_row):
        """Step.Information of one env (null for CartPole, CartPoleEnv.cs:185)."""
        return None


def shard_envs(total_envs, rank, world_size):
    """Contiguous shard of a global batch: (num_envs, env_id_offset) of `rank`.  Global env id =
    env_id_offset + local id keys the Philox streams, so trajectories do not depend on the GPU count."""
    if not (0 <= rank < world_size) or total_envs < world_size:
        raise ValueError("bad shard request: total=%d rank=%d world=%d" % (total_envs, rank, world_size))
    base, extra = divmod(int(total_envs), int(world_size))
    n = base + (1 if rank < extra else 0)
    off = rank * base + min(rank, extra)
    return n, off


def bind_host_to_device(device):
    """Pin the calling process to the CPU cores next to GPU `device` (its PCIe root's NUMA node) and return them.

    The host-buffer step writes ~21 B per env-step into page-locked host memory; with one process per GPU, a
    process that runs (and therefore first-touches its pinned buffers) on the other socket pushes every one of
    those bytes across the socket interconnect.  Call this BEFORE allocating host buffers.  Returns None when
    the topology cannot be read (no sysfs entry, single node): the affinity is then left alone."""
    import os
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(int(device)), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/local_cpulist" % (dom[-4:], rest)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def nccl_unique_id():
    buf = (C.c_uint8 * 128)()
    N.check(N.lib().gymcuda_nccl_unique_id(buf))
    return bytes(buf)


class CartPoleVecEnv(CudaVecEnv):
    """Batched Gym.Environments.Envs.Classic.CartPoleEnv (CartPoleEnv.cs)."""
    ENV_KIND = N.CARTPOLE


class PendulumVecEnv(CudaVecEnv):
    ENV_KIND = N.PENDULUM


class MountainCarVecEnv(CudaVecEnv):
    ENV_KIND = N.MOUNTAINCAR


class MountainCarContinuousVecEnv(CudaVecEnv):
    ENV_KIND = N.MOUNTAINCAR_CONT


class AcrobotVecEnv(CudaVecEnv):
    ENV_KIND = N.ACROBOT


class LunarLanderVecEnv(CudaVecEnv):
    """Batched Gym.Environments.Envs.Aether.LunarLanderEnv (LunarLanderEnv.cs); continuous=True for the Box action space."""
    ENV_KIND = N.LUNARLANDER

    def __init__(self, num_envs, continuous=False, **kw):
        self.ENV_KIND = N.LUNARLANDER_CONT if continuous else N.LUNARLANDER
        self.ContinuousMode = bool(continuous)
        super().__init__(num_envs, **kw)

    def _information(self, o):
        """The Dict LunarLanderEnv.Step fills (LunarLanderEnv.cs:740-746); every entry is a view of the observation."""
        return {"pos": (float(o[0]), float(o[1])), "velocity": (float(o[2]), float(o[3])), "angle": float(o[4]),
                "omega": float(o[5]), "LeftContact": bool(o[6]), "RightContact": bool(o[7])}


FAMILIES = {
    "CartPole-v1": CartPoleVecEnv,
    "Pendulum-v1": PendulumVecEnv,
    "MountainCar-v0": MountainCarVecEnv,
    "MountainCarContinuous-v0": MountainCarContinuousVecEnv,
    "Acrobot-v1": AcrobotVecEnv,
    "LunarLander-v2": LunarLanderVecEnv,
}


def make(name, num_envs, **kw):
    return FAMILIES[name](num_envs, **kw)
