"""CPU, world_size 2 over gloo: the N>1 host logic.  Each rank owns a contiguous shard of the global
batch (gymnet_b200.shard_envs -> num_envs, env_id_offset), steps it independently (no data-path
collective), and the optional observation all-gather lays shards out [world][n][obs_dim] -- the layout
gymcuda_allgather_obs produces with ncclAllGather on the GPU box.  Without a GPU the per-rank stepping
is done by the CPU oracle (engine-arithmetic twin); the GPU version of this test is in test_gpu_multi.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from gymnet_b200 import shard_envs

TOTAL, K = 96, 40


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, off = shard_envs(TOTAL, rank, world)
    env = O.OracleEnv(O.CARTPOLE, n, seed=11, env_id_offset=off, auto_reset=True, mode=O.MODE_F32)
    env.reset()
    obs, rew, done, act = env.rollout_random(K)
    last = torch.from_numpy(np.ascontiguousarray(obs[-1]))
    gathered = [torch.empty_like(last) for _ in range(world)]
    dist.all_gather(gathered, last)                     # [world][n][obs_dim]
    steps = torch.tensor([float(n * K)]); dist.all_reduce(steps)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.stack(gathered).numpy())
        np.save(os.path.join(out_dir, "steps.npy"), steps.numpy())
    dist.destroy_process_group()


def test_shard_envs_partitions():
    for total, world in [(96, 2), (65536, 8), (10, 3), (7, 7)]:
        parts = [shard_envs(total, r, world) for r in range(world)]
        assert sum(n for n, _ in parts) == total
        assert parts[0][1] == 0 and all(parts[i][1] + parts[i][0] == parts[i + 1][1] for i in range(world - 1))
    with pytest.raises(ValueError):
        shard_envs(4, 2, 2)


def test_two_rank_sharded_rollout_matches_single_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    gathered = np.load(tmp_path / "gathered.npy")
    assert gathered.shape == (world, TOTAL // world, 4)
    full = O.OracleEnv(O.CARTPOLE, TOTAL, seed=11, auto_reset=True, mode=O.MODE_F32)
    full.reset()
    obs, _, _, _ = full.rollout_random(K)
    assert np.array_equal(gathered.reshape(TOTAL, 4), obs[-1])
    assert np.load(tmp_path / "steps.npy")[0] == TOTAL * K
