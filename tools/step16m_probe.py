"""CartPole step_kernel at 16 777 216 envs (state 256 MiB >> L2) for N launches -- target of ncu captures."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G
n = 16777216; steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
env = G.CartPoleVecEnv(n, seed=0, auto_reset=True); env.ResetBatch()
dev = torch.device("cuda", 0)
obs = torch.empty((n, 4), device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
act = torch.randint(0, 2, (n,), dtype=torch.int32, device=dev)
for _ in range(steps):
    env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
env.Sync()
