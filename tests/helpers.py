"""Shared helpers of the parity tests.  The oracle (tests/oracle_lib.py) is the checker; the CUDA
path under test is always reached through the C ABI (gymnet_b200 -> ctypes -> libgymcuda.so)."""
import numpy as np

import oracle_lib as O
import gymnet_b200 as G

KINDS = {
    "CartPole-v1": O.CARTPOLE,
    "Pendulum-v1": O.PENDULUM,
    "MountainCar-v0": O.MOUNTAINCAR,
    "MountainCarContinuous-v0": O.MOUNTAINCAR_CONT,
    "Acrobot-v1": O.ACROBOT,
    "LunarLander-v2": O.LUNARLANDER,
}

# natural magnitude of each state component: |a-b| <= RTOL * max(|a|, |b|, scale)
STATE_SCALE = {
    "CartPole-v1": np.array([2.4, 1.0, 0.21, 1.0]),
    "Pendulum-v1": np.array([3.14, 8.0]),
    "MountainCar-v0": np.array([1.2, 0.07]),
    "MountainCarContinuous-v0": np.array([1.2, 0.07]),
    "Acrobot-v1": np.array([3.14, 3.14, 12.0, 28.0]),
}
RTOL = 1e-5   # north_star: float state within 1e-5 relative


def rel_err(a, b, scale):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)
    return np.abs(a - b) / den


def random_actions(env, rng, n):
    if env.act_n > 0:
        return rng.integers(0, env.act_n, size=n).astype(np.int32)
    lo = np.array(env.info.act_low[:env.act_dim]); hi = np.array(env.info.act_high[:env.act_dim])
    return rng.uniform(lo, hi, size=(n, env.act_dim)).astype(np.float32)


def random_states(name, rng, n):
    """Float32 states spread over (and beyond) the region each env visits."""
    if name == "CartPole-v1":
        s = rng.uniform([-2.6, -3, -0.25, -3.5], [2.6, 3, 0.25, 3.5], size=(n, 4))
    elif name == "Pendulum-v1":
        s = rng.uniform([-10, -8], [10, 8], size=(n, 2))
    elif name in ("MountainCar-v0", "MountainCarContinuous-v0"):
        s = rng.uniform([-1.2, -0.07], [0.6, 0.07], size=(n, 2))
    elif name == "Acrobot-v1":
        # the whole reachable state box: angles wrapped to [-pi, pi], velocities clamped to (4 pi, 9 pi); beyond (9, 18)
        # rad/s the engine steps in double precision, so 1e-5 holds over all of it
        s = rng.uniform([-3.14159, -3.14159, -12.5663, -28.2743], [3.14159, 3.14159, 12.5663, 28.2743], size=(n, 4))
    else:
        raise KeyError(name)
    return s.astype(np.float32)
