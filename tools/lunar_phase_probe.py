"""Per-warp phase timing of one LunarLander step (timing probe; needs the -DLUNAR_PHASE_CLOCKS variant of the library:
GYMCUDA_LIB=gym.net_b200/csrc/exp/libgymcuda_phase.so).  From the snapshot of tools/lunar_split_probe.py: one step, then the
clock64() samples of every thread at the phase boundaries -> where the SLOWEST warps of each kernel spend their cycles."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G
n = 65536
env = G.LunarLanderVecEnv(n, seed=0, auto_reset=True, time_limit=1000)
dev = torch.device("cuda", 0)
obs = torch.empty((n, 8), device=dev); rew = torch.empty(n, device=dev); done = torch.empty(n, dtype=torch.uint8, device=dev)
z = np.load("/tmp/lunar_snapshot.npz")
env.ResetBatch()
act = torch.from_numpy(z["act"]).to(dev)
L = C.CDLL(os.environ["GYMCUDA_LIB"])
buf = np.zeros((2, 10, 65536), np.int64)
NAMES = ["pre_physics", "collide", "solve init", "velocity loop", "integrate + position loop", "store/sleep/broadphase", "post_physics"]
ORDER = [0, 1, 2, 3, 4, 5, 6, 7]   # phase marks in time order (8 / 9, around the state load / store, are not wired)
for rep in range(3):
    env.SetState(z["st"], z["aux"], int(z["t"]))
    env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr()); env.Sync()
    assert L.gymcuda_debug_lunar_phase(buf.ctypes.data_as(C.c_void_p)) == 0
print(json.dumps({"nonzero_per_phase": [[int((buf[h][k] != 0).sum()) for k in range(10)] for h in range(2)]}))
for hp, name in ((1, "contact kernel"), (0, "free-flight kernel")):
    t = buf[hp][ORDER]
    used = (t[0] != 0) & (t[-1] != 0) & (t[-1] > t[0])
    k = int(used.sum())
    if k == 0:
        continue
    idx = np.nonzero(used)[0]
    warps = k // 32
    t = t[:, idx[:warps * 32]].astype(np.float64)
    t0 = t[0].min()
    tw = t.reshape(8, warps, 32)
    wend = tw[-1].max(1) - t0; wstart = tw[0].min(1) - t0
    dur = np.maximum(np.diff(tw, axis=0), 0).max(2)          # per warp, per phase: the slowest lane
    order = np.argsort(-wend)
    out = {"kernel": name, "threads": k, "warps": warps, "kernel_cycles": float(wend.max()),
           "warp_start_cycles_p50_max": [float(np.median(wstart)), float(wstart.max())], "slowest_warps": []}
    for w in order[:4]:
        out["slowest_warps"].append({"end": float(wend[w]), "start": float(wstart[w]), "phases": {NAMES[i]: float(dur[i, w]) for i in range(7)}})
    out["median_warp_phases"] = {NAMES[i]: float(np.median(dur[i])) for i in range(7)}
    out["p90_warp_phases"] = {NAMES[i]: float(np.percentile(dur[i], 90)) for i in range(7)}
    out["median_warp_total"] = float(np.median(wend - wstart))
    print(json.dumps(out))
