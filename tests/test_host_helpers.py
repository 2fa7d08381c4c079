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
