"""Per-launch time of the LunarLander step for the first N steps after a reset (65 536 landers, random policy,
auto-reset): all landers start in free flight and reach the ground around step 100, so the series separates the
cost of the joint-only solve from the cost of landers in contact.  Prints JSON lines."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G
n = int(os.environ.get("LUNAR_N", "65536")); steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
env = G.LunarLanderVecEnv(n, seed=0, auto_reset=True, time_limit=1000)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); env.SetStream(stream.cuda_stream)
obs = torch.empty((n, 8), device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
acts = torch.randint(0, 4, (steps, n), dtype=torch.int32, device=dev)
for rep in range(2):   # first pass warms up
    env.Seed(0); env.ResetBatch()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    legs = []
    torch.cuda.synchronize()
    for k in range(steps):
        ev[k].record(stream)
        env.StepDevice(acts[k].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
        if rep == 1 and k % 10 == 9:
            legs.append(float((obs[:, 6:] > 0).any(1).float().mean()))
    ev[steps].record(stream); torch.cuda.synchronize()
    if rep == 1:
        ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        for k in range(0, steps, 10):
            print(json.dumps({"steps": "%d-%d" % (k, k + 9), "ms_per_step_min": min(ms[k:k + 10]), "ms_per_step_mean": sum(ms[k:k + 10]) / 10, "legs_down_frac": legs[k // 10]}))
