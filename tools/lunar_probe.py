"""Step LunarLander (65 536 envs, random policy, auto-reset) for N launches -- target of ncu captures."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G
n = 65536; steps = int(sys.argv[1]) if len(sys.argv) > 1 else 160
env = G.LunarLanderVecEnv(n, seed=0, auto_reset=True, time_limit=1000); env.ResetBatch()
dev = torch.device("cuda", 0)
obs = torch.empty((n, 8), device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
act = torch.randint(0, 4, (n,), dtype=torch.int32, device=dev)
for _ in range(steps):
    env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
env.Sync()
print("legs down", float((obs[:, 6:] > 0).any(1).float().mean()))
