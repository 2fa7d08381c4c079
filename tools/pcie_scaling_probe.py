"""Why does the host-buffer (e2e) step scale worse than the kernels?  Run under torchrun with N ranks of ONE box:
every rank drives its own GPU with 65 536 CartPole envs at the same time, three ways, and reports the slowest rank:
  zero_copy   gymcuda_step with page-locked host buffers (the kernel reads / writes host memory over PCIe; bench.py's e2e)
  dma         gymcuda_step_device + ONE copy-engine D2H of obs|reward|done (1.38 MB) + H2D of the actions + synchronise
  dma_only    the two DMAs + synchronise, no kernel (the PCIe floor of the step's bytes)
Prints one JSON line on rank 0: per-mode us per step (max over ranks) and the D2H GB/s per GPU it implies, plus the CPU
affinity / NUMA node of every rank.  python -m torch.distributed.run --nproc-per-node N tools/pcie_scaling_probe.py [bind]"""
import ctypes as C, json, os, sys, time
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G  # noqa: E402

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
cpus = G.bind_host_to_device(local) if bind else None
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 65536
env = G.make("CartPole-v1", n, seed=0, device=local, env_id_offset=rank * n, auto_reset=True)
env.ResetBatch()
L = G._native.lib()
h_out = torch.empty(n * 21, dtype=torch.uint8).pin_memory(); d_out = torch.empty(n * 21, dtype=torch.uint8, device=dev)
h_act = torch.zeros(n, dtype=torch.int32).pin_memory(); d_act = torch.zeros(n, dtype=torch.int32, device=dev)
d_obs = d_out[: n * 16].view(torch.float32).view(n, 4); d_rew = d_out[n * 16: n * 20].view(torch.float32); d_done = d_out[n * 20:]
zargs = (env._h, C.c_void_p(h_act.data_ptr()), C.c_void_p(h_out.data_ptr()), C.c_void_p(h_out.data_ptr() + n * 16), C.c_void_p(h_out.data_ptr() + n * 20))
stream = torch.cuda.current_stream(); env.SetStream(stream.cuda_stream)


def zero_copy():
    L.gymcuda_step(*zargs)


def dma():
    d_act.copy_(h_act, non_blocking=True)
    env.StepDevice(d_act.data_ptr(), d_obs.data_ptr(), d_rew.data_ptr(), d_done.data_ptr())
    h_out.copy_(d_out, non_blocking=True); torch.cuda.synchronize()


def dma_only():
    d_act.copy_(h_act, non_blocking=True); h_out.copy_(d_out, non_blocking=True); torch.cuda.synchronize()


def timed(fn, reps=1500):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    us = (time.perf_counter() - t0) / reps * 1e6
    if world > 1:
        t = torch.tensor([us], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us = float(t.item())
    return us


res = {"ranks": world, "bind": bind}
for name, fn in (("zero_copy", zero_copy), ("dma", dma), ("dma_only", dma_only)):
    us = timed(fn)
    res[name + "_us"] = round(us, 2)
    res[name + "_d2h_gbs_per_gpu"] = round(n * 21 / us / 1e3, 2)
    res[name + "_env_steps_per_s"] = world * n / (us * 1e-6)
aff = sorted(os.sched_getaffinity(0))
info = [None] * world
mine = {"rank": rank, "cpus": "%d-%d (%d)" % (aff[0], aff[-1], len(aff))}
try:
    bus = torch.cuda.get_device_properties(local).pci_bus_id
    mine["numa"] = open("/sys/bus/pci/devices/0000:%02x:00.0/numa_node" % bus).read().strip()
except Exception:
    pass
if world > 1:
    dist.all_gather_object(info, mine)
else:
    info = [mine]
if rank == 0:
    res["ranks_info"] = info
    print(json.dumps(res))
env.Close()
if world > 1:
    dist.destroy_process_group()
