"""torchrun worker of tests/test_gpu_multi.py: one process per GPU, independent env shards, then the
optional NCCL all-gather of observations (gymcuda_allgather_obs).  torch.distributed only ships the
128-byte ncclUniqueId between ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G  # noqa: E402


class _RawCuda:
    """CUDA array interface over a raw device pointer returned by the C ABI."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _second_id(rank):
    ids = [G.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    return ids[0]


def main():
    out_dir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    total = 4096
    n, off = G.shard_envs(total, rank, world)
    env = G.CartPoleVecEnv(n, seed=21, device=local, env_id_offset=off, auto_reset=True)
    ids = [G.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    env.CommInit(ids[0], rank, world)
    env.ResetBatch()
    env.RolloutRandom(50, want=())
    out = torch.empty((world, n, env.obs_dim), dtype=torch.float32, device="cuda")
    env.AllGatherObs(out.data_ptr())        # observations recomputed from state, gathered over NVLink
    env.Sync()
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), out.cpu().numpy())
    dist.barrier()
    env.Close()
    # fused step + gather over peer memory (no NCCL on the step path) vs ncclAllGather of the same step
    fz = G.CartPoleVecEnv(n, seed=33, device=local, env_id_offset=off, auto_reset=True)
    fz.CommInit(_second_id(rank), rank, world)
    handles = [None] * world
    dist.all_gather_object(handles, fz.GatherCreate(rank, world))
    fz.GatherOpen(handles)
    fz.ResetBatch()
    f_rew = torch.empty((n,), dtype=torch.float32, device="cuda")
    f_done = torch.empty((n,), dtype=torch.uint8, device="cuda")
    ref_out = torch.empty((world, n, 4), dtype=torch.float32, device="cuda")
    gen = torch.Generator(device="cuda"); gen.manual_seed(100 + rank)
    ok = True
    for step in range(25):
        a = torch.randint(0, 2, (n,), dtype=torch.int32, device="cuda", generator=gen)
        torch.cuda.synchronize()
        ptr = fz.StepGatherDevice(a.data_ptr(), f_rew.data_ptr(), f_done.data_ptr())
        fz.GatherWait()
        fz.AllGatherObs(ref_out.data_ptr())          # NCCL gather of the same observations (last_obs = own slot)
        fz.Sync()
        fused = torch.as_tensor(_RawCuda(ptr, (world, n, 4)), device="cuda")   # zero-copy view of the gather buffer
        ok = ok and bool(torch.equal(fused, ref_out))
    flag = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(out_dir, "fused_ok.npy"), flag.numpy())
    dist.barrier()
    fz.Close()
    # config 5 shape: LunarLander shards, one step, obs all-gather of the step's observations
    n2, off2 = G.shard_envs(2048, rank, world)
    ll = G.LunarLanderVecEnv(n2, seed=5, device=local, env_id_offset=off2, auto_reset=True)
    ll.CommInit(_second_id(rank), rank, world)
    ll.ResetBatch()
    d_obs = torch.empty((n2, 8), dtype=torch.float32, device="cuda")
    d_rew = torch.empty((n2,), dtype=torch.float32, device="cuda")
    d_done = torch.empty((n2,), dtype=torch.uint8, device="cuda")
    acts = torch.full((n2,), 2, dtype=torch.int32, device="cuda")
    for _ in range(20):
        ll.StepDevice(acts.data_ptr(), d_obs.data_ptr(), d_rew.data_ptr(), d_done.data_ptr())
    out2 = torch.empty((world, n2, 8), dtype=torch.float32, device="cuda")
    ll.AllGatherObs(out2.data_ptr(), d_obs.data_ptr())
    ll.Sync()
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered_lunar.npy"), out2.cpu().numpy())
    dist.barrier()
    ll.Close()
    # LunarLander with the observation gather: the partitioned step (three partition kernels + two step kernels) fills the
    # rank's own slot, gather_push_kernel sends it to the peers -- against ncclAllGather of the same step, 30 steps
    n3, off3 = G.shard_envs(4096, rank, world)
    lg = G.LunarLanderVecEnv(n3, seed=9, device=local, env_id_offset=off3, auto_reset=True, time_limit=25)
    lg.CommInit(_second_id(rank), rank, world)
    handles = [None] * world
    dist.all_gather_object(handles, lg.GatherCreate(rank, world))
    lg.GatherOpen(handles)
    lg.ResetBatch()
    g_rew = torch.empty((n3,), dtype=torch.float32, device="cuda"); g_done = torch.empty((n3,), dtype=torch.uint8, device="cuda")
    ref3 = torch.empty((world, n3, 8), dtype=torch.float32, device="cuda")
    ok = True
    for step in range(30):
        a = torch.randint(0, 4, (n3,), dtype=torch.int32, device="cuda", generator=gen)
        torch.cuda.synchronize()
        ptr = lg.StepGatherDevice(a.data_ptr(), g_rew.data_ptr(), g_done.data_ptr())
        lg.GatherWait()
        lg.AllGatherObs(ref3.data_ptr())
        lg.Sync()
        fused = torch.as_tensor(_RawCuda(ptr, (world, n3, 8)), device="cuda")
        ok = ok and bool(torch.equal(fused, ref3))
        dist.barrier()   # nobody starts the next step's pushes into a buffer a peer is still comparing
    flag = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(out_dir, "fused_lunar_ok.npy"), flag.numpy())
    dist.barrier()
    lg.Close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
