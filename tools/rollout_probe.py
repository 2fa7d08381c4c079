"""Launch the fused rollout kernel of the named envs a few times -- the target of ncu captures.
   python tools/rollout_probe.py [env[:num_envs[:inner]] ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G  # noqa: E402

DEFAULT = ["CartPole-v1:65536:512", "Pendulum-v1:262144:128", "MountainCarContinuous-v0:262144:128",
           "MountainCar-v0:262144:128", "Acrobot-v1:131072:128"]


def main():
    dev = torch.device("cuda", 0)
    for spec in (sys.argv[1:] or DEFAULT):
        parts = spec.split(":")
        name, n, K = parts[0], int(parts[1]), int(parts[2])
        env = G.make(name, n, seed=0, auto_reset=True)
        env.ResetBatch()
        od, ad = env.obs_dim, env.act_dim
        obs = torch.empty((K, n, od), dtype=torch.float32, device=dev)
        rew = torch.empty((K, n), dtype=torch.float32, device=dev)
        done = torch.empty((K, n), dtype=torch.uint8, device=dev)
        act = torch.empty((K, n, ad), dtype=torch.int32 if env.act_n > 0 else torch.float32, device=dev)
        for _ in range(3):
            env.RolloutRandomDevice(K, obs.data_ptr(), rew.data_ptr(), done.data_ptr(), act.data_ptr())
        env.Sync()
        print(name, n, K, "episodes/env", float(done.sum().item()) / n, flush=True)
        env.Close()


if __name__ == "__main__":
    main()
