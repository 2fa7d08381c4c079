#!/usr/bin/env python
"""Throughput of every env family at the BASELINE.json config sizes (configs 2-5), both entry points:
   rollout  = fused random-policy rollout kernel, trajectory streamed to HBM (device buffers)
   step     = one launch per step (gymcuda_step_device), actions resident on device
Prints one JSON line per (env, mode).  Roofline denominators: MEASURED_PEAKS.json hbm_gbs.
Run on the GPU box:  python tools/measure_envs.py > gpurun_out/envs.jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gymnet_b200 as G  # noqa: E402

CONFIGS = [   # (name, envs per GPU, rollout inner steps, kwargs)
    ("CartPole-v1", 65536, 512, {}),
    ("Pendulum-v1", 262144, 128, {}),
    ("MountainCarContinuous-v0", 262144, 128, {}),
    ("MountainCar-v0", 262144, 128, {}),
    ("Acrobot-v1", 131072, 128, {}),
    ("LunarLander-v2", 65536, 32, {"time_limit": 1000}),
]


def main():
    only = os.environ.get("MEASURE_MODE", "")          # "rollout" / "step" / "" = both
    names = [x for x in os.environ.get("MEASURE_ENVS", "").split(",") if x]
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    for name, n, K, kw in CONFIGS:
        if names and name not in names:
            continue
        env = G.make(name, n, seed=0, auto_reset=True, **kw)
        env.SetStream(stream.cuda_stream)
        env.ResetBatch()
        od, ad = env.obs_dim, env.act_dim
        adt = torch.int32 if env.act_n > 0 else torch.float32
        obs = torch.empty((K, n, od), dtype=torch.float32, device=dev)
        rew = torch.empty((K, n), dtype=torch.float32, device=dev)
        done = torch.empty((K, n), dtype=torch.uint8, device=dev)
        act = torch.empty((K, n, ad), dtype=adt, device=dev)
        reps = 10
        for _ in range(3):
            env.RolloutRandomDevice(K, obs.data_ptr(), rew.data_ptr(), done.data_ptr(), act.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(reps):
            env.RolloutRandomDevice(K, obs.data_ptr(), rew.data_ptr(), done.data_ptr(), act.data_ptr())
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bytes_step = od * 4 + 4 + 1 + ad * 4
        rate = n * K / (ms * 1e-3)
        print(json.dumps({"env": name, "mode": "rollout", "num_envs": n, "inner": K, "ms_per_launch": ms,
                          "env_steps_per_s": rate, "algo_bytes_per_env_step": bytes_step,
                          "achieved_gbs": rate * bytes_step / 1e9, "frac_of_measured_hbm": rate * bytes_step / 1e9 / peak,
                          "episodes_per_env": float(done.sum().item()) / n}), flush=True)
        if only == "rollout":
            env.Close()
            del obs, rew, done, act
            continue
        # per-launch step with device-resident actions
        a1 = act[0].contiguous()
        steps = 200
        for _ in range(10):
            env.StepDevice(a1.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(steps):
            env.StepDevice(a1.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        sd = env.state_dim
        bytes_step = 2 * sd * 4 + ad * 4 + od * 4 + 4 + 1
        rate = n / (ms * 1e-3)
        print(json.dumps({"env": name, "mode": "step_device", "num_envs": n, "ms_per_launch": ms, "env_steps_per_s": rate,
                          "algo_bytes_per_env_step": bytes_step, "achieved_gbs": rate * bytes_step / 1e9,
                          "frac_of_measured_hbm": rate * bytes_step / 1e9 / peak}), flush=True)
        env.Close()
        del obs, rew, done, act


if __name__ == "__main__":
    main()
