"""What the host-buffer step is made of on this box: DMA times of its buffers, the bare launch + sync, and the
zero-copy gymcuda_step itself.  python tools/pcie_probe.py"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G  # noqa: E402

n = 65536
dev = torch.device("cuda", 0)
env = G.make("CartPole-v1", n, seed=0, auto_reset=True)
env.ResetBatch()
L = G._native.lib()
h_out = torch.empty(n * 21, dtype=torch.uint8).pin_memory()
d_out = torch.empty(n * 21, dtype=torch.uint8, device=dev)
h_act = torch.zeros(n, dtype=torch.int32).pin_memory()
d_act = torch.zeros(n, dtype=torch.int32, device=dev)
d_obs = torch.empty((n, 4), device=dev); d_rew = torch.empty(n, device=dev); d_done = torch.empty(n, dtype=torch.uint8, device=dev)


def timeit(fn, reps=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def d2h():
    h_out.copy_(d_out, non_blocking=True); torch.cuda.synchronize()


def h2d():
    d_act.copy_(h_act, non_blocking=True); torch.cuda.synchronize()


def launch_sync():
    env.StepDevice(d_act.data_ptr(), d_obs.data_ptr(), d_rew.data_ptr(), d_done.data_ptr()); env.Sync()


args = (env._h, C.c_void_p(h_act.data_ptr()), C.c_void_p(h_out.data_ptr()), C.c_void_p(h_out.data_ptr() + n * 16),
        C.c_void_p(h_out.data_ptr() + n * 20))


def zero_copy():
    L.gymcuda_step(*args)


pag_obs = torch.empty((n, 4)); pag_rew = torch.empty(n); pag_done = torch.empty(n, dtype=torch.uint8); pag_act = torch.zeros(n, dtype=torch.int32)
pargs = (env._h, C.c_void_p(pag_act.data_ptr()), C.c_void_p(pag_obs.data_ptr()), C.c_void_p(pag_rew.data_ptr()), C.c_void_p(pag_done.data_ptr()))


def pageable():
    L.gymcuda_step(*pargs)


# what the C# shim does: result arrays registered once (gymcuda_host_register), the caller's action array pageable
reg_obs = torch.empty((n, 4)); reg_rew = torch.empty(n); reg_done = torch.empty(n, dtype=torch.uint8)
for t_ in (reg_obs, reg_rew, reg_done):
    G._native.check(L.gymcuda_host_register(C.c_void_p(t_.data_ptr()), t_.numel() * t_.element_size()))
rargs = (env._h, C.c_void_p(pag_act.data_ptr()), C.c_void_p(reg_obs.data_ptr()), C.c_void_p(reg_rew.data_ptr()), C.c_void_p(reg_done.data_ptr()))


def shim_like():
    L.gymcuda_step(*rargs)


print("D2H 1.376 MB pinned DMA + sync      %.1f us" % timeit(d2h))
print("H2D 0.262 MB pinned DMA + sync      %.1f us" % timeit(h2d))
print("step_device launch + sync           %.1f us" % timeit(launch_sync))
print("gymcuda_step zero-copy (pinned)     %.1f us" % timeit(zero_copy))
print("gymcuda_step pageable host buffers  %.1f us" % timeit(pageable, 300))
print("gymcuda_step registered results + pageable actions (C# shim)  %.1f us" % timeit(shim_like, 1000))
