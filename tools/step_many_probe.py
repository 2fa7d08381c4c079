"""Launch gymcuda_step_many_device (the rollout kernel fed with the caller's actions) a few times -- the target of ncu captures.
   python tools/step_many_probe.py [env[:num_envs[:k]] ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    for spec in (sys.argv[1:] or ["CartPole-v1:65536:512"]):
        parts = spec.split(":")
        name, n, K = parts[0], int(parts[1]), int(parts[2])
        env = G.make(name, n, seed=0, auto_reset=True)
        env.ResetBatch()
        od, ad = env.obs_dim, env.act_dim
        obs = torch.empty((K, n, od), dtype=torch.float32, device=dev)
        rew = torch.empty((K, n), dtype=torch.float32, device=dev)
        done = torch.empty((K, n), dtype=torch.uint8, device=dev)
        if env.act_n > 0:
            act = torch.randint(0, env.act_n, (K, n, ad), dtype=torch.int32, device=dev)
        else:
            act = torch.rand((K, n, ad), device=dev) * 2 - 1
        for _ in range(4):
            env.StepManyDevice(K, act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
        env.Sync()
        print(name, n, K, "episodes/env", float(done.sum().item()) / n, flush=True)


if __name__ == "__main__":
    main()
