"""Host-side helpers that need no GPU: NUMA binding degrades gracefully, shards tile the batch."""
import os

import gymnet_b200 as G


def test_bind_host_to_device_is_harmless_without_a_gpu():
    before = os.sched_getaffinity(0)
    cpus = G.bind_host_to_device(0)          # no nvidia-smi / no sysfs entry here: leaves the affinity alone
    assert cpus is None or set(cpus) <= before
    if cpus is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_shards_tile_the_global_batch():
    for total, world in ((65536, 8), (1048576, 8), (1000003, 7), (8, 8)):
        off = 0
        for rank in range(world):
            n, o = G.shard_envs(total, rank, world)
            assert o == off and n >= 1
            off += n
        assert off == total


def test_linear_slop_squared_threshold_is_the_exact_sqrt_boundary():
    """lunar_core.cuh joint_position compares dot(C, C) with LINEAR_SLOP_SQ_MAX instead of sqrtf(dot(C, C)) with linearSlop:
    the constant must be the LARGEST float32 whose correctly rounded square root does not exceed float32(0.005)."""
    import math
    import re
    import numpy as np
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym.net_b200", "csrc", "lunar_core.cuh")).read()
    c = np.float32(float(re.search(r"LINEAR_SLOP_SQ_MAX = ([0-9.e+-]+)f", src).group(1)))
    slop = np.float32(0.005)
    root = lambda x: np.float32(math.sqrt(float(x)))   # double sqrt rounded once: the correctly rounded float32 root
    assert root(c) <= slop and root(np.nextafter(c, np.float32(np.inf), dtype=np.float32)) > slop
    x = c
    for _ in range(1000):   # monotone on both sides of the boundary
        x = np.nextafter(x, np.float32(-np.inf), dtype=np.float32)
        assert root(x) <= slop
