"""gym.net_b200: Blackwell-native vectorised environment engine behind the Gym.NET API surface.

Holds only what the hot path needs: csrc/ (CUDA kernels + the C ABI, libgymcuda.so) and the
host-side mirror of the reference's Env / VecEnv / Space interface for that path.  The directory
name is not a valid Python identifier; import it through the repo-root shim `gymnet_b200`.
"""
from . import _native
from ._native import GymCudaError, InvalidActionError
from .spaces import Box, Discrete, Space
from .vector import (AcrobotVecEnv, CartPoleVecEnv, CudaVecEnv, LunarLanderVecEnv,
                     MountainCarContinuousVecEnv, MountainCarVecEnv, PendulumVecEnv, Step, make,
                     bind_host_to_device, nccl_unique_id, shard_envs)

__all__ = ["GymCudaError", "InvalidActionError", "Box", "Discrete", "Space", "CudaVecEnv", "Step", "make",
           "CartPoleVecEnv", "PendulumVecEnv", "MountainCarVecEnv", "MountainCarContinuousVecEnv",
           "AcrobotVecEnv", "LunarLanderVecEnv", "nccl_unique_id", "shard_envs", "bind_host_to_device"]
