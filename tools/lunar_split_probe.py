"""Where does a LunarLander step's time go?  Timing probe (NOT physics): from one snapshot of 65 536 landers after 200 correct
steps, the time of ONE step under library variants built with 1 velocity iteration and / or 1 position iteration
(-DLUNAR_VEL_ITERS=1 / -DLUNAR_POS_ITERS=1, gym.net_b200/csrc/exp/).  Usage on the GPU box:
    python tools/lunar_split_probe.py snapshot      # in-tree library: 200 steps, writes /tmp/lunar_snapshot.npz
    GYMCUDA_LIB=... python tools/lunar_split_probe.py time <label>
"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G
n = 65536
env = G.LunarLanderVecEnv(n, seed=0, auto_reset=True, time_limit=1000)
dev = torch.device("cuda", 0)
obs = torch.empty((n, 8), device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
if sys.argv[1] == "snapshot":
    env.ResetBatch()
    acts = torch.randint(0, 4, (200, n), dtype=torch.int32, device=dev)
    for k in range(200):
        env.StepDevice(acts[k].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
    env.Sync()
    st, aux, t = env.GetState()
    np.savez("/tmp/lunar_snapshot.npz", st=st, aux=aux, t=t, act=acts[0].cpu().numpy())
else:
    z = np.load("/tmp/lunar_snapshot.npz")
    env.ResetBatch()
    act = torch.from_numpy(z["act"]).to(dev)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); env.SetStream(stream.cuda_stream)
    ms = []
    for rep in range(8):
        env.SetState(z["st"], z["aux"], int(z["t"]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
        e1.record(stream); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    print(json.dumps({"variant": sys.argv[2], "ms_one_step_min": min(ms[2:]), "ms_one_step_median": float(np.median(ms[2:]))}))
