"""Per-launch time of CartPole step_kernel at 16 777 216 envs over the first 60 steps after a reset (the done fraction
comes in waves: nobody can fail before step ~8), prints JSON."""
import json, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G
n = 16777216
env = G.CartPoleVecEnv(n, seed=0, auto_reset=True); env.ResetBatch()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); env.SetStream(stream.cuda_stream)
obs = torch.empty((n, 4), device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
act = torch.randint(0, 2, (n,), dtype=torch.int32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(61)]
torch.cuda.synchronize()
for k in range(60):
    ev[k].record(stream)
    env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
ev[60].record(stream); torch.cuda.synchronize()
us = [ev[k].elapsed_time(ev[k + 1]) * 1e3 for k in range(60)]
print(json.dumps({"variant": sys.argv[1] if len(sys.argv) > 1 else "", "us_no_done_steps_2_6": sum(us[2:7]) / 5, "us_steady_steps_40_59": sum(us[40:]) / 20,
                  "us": [round(x) for x in us], "episodes": env.Stats()["episodes"]}))
