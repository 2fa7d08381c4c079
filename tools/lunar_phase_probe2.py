"""Per-warp phase timing of one LunarLander step, any launch shape (one lane / three lanes per lander, wide CTAs): needs the
-DLUNAR_PHASE_CLOCKS variant of the library (GYMCUDA_LIB=...) and the snapshot of tools/lunar_split_probe.py.  Warps are
grouped by thread id; slots 8 / 9 are %globaltimer stamps, so the two kernels of the partition can be laid on one time axis."""
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
for rep in range(3):
    env.SetState(z["st"], z["aux"], int(z["t"]))
    buf[:] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    env.StepDevice(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr())
    e1.record(); env.Sync(); torch.cuda.synchronize()
    assert L.gymcuda_debug_lunar_phase(buf.ctypes.data_as(C.c_void_p)) == 0
print(json.dumps({"label": sys.argv[1] if len(sys.argv) > 1 else "", "step_ms_events": e0.elapsed_time(e1)}))
t_origin = None
for hp, name in ((1, "contact kernel"), (0, "free-flight kernel")):
    b = buf[hp]
    used = np.nonzero((b[8] != 0) & (b[9] != 0))[0]
    if len(used) == 0:
        continue
    ns0, ns1 = b[8][used], b[9][used]
    if t_origin is None:
        t_origin = min(buf[1][8][buf[1][8] != 0].min() if (buf[1][8] != 0).any() else 1 << 62, buf[0][8][buf[0][8] != 0].min())
    warp = used // 32
    ws = np.unique(warp)
    dur = {}
    tot = []
    for w in ws:
        idx = used[warp == w]
        t = b[:8, idx].astype(np.float64)
        d = np.maximum(np.diff(t, axis=0), 0).max(1)
        dur[w] = d
        tot.append(t[7].max() - t[0].min())
    D = np.array([dur[w] for w in ws]); tot = np.array(tot)
    out = {"kernel": name, "threads": int(len(used)), "warps": int(len(ws)),
           "first_start_us": float((ns0.min() - t_origin) / 1e3), "last_start_us": float((ns0.max() - t_origin) / 1e3),
           "first_end_us": float((ns1.min() - t_origin) / 1e3), "last_end_us": float((ns1.max() - t_origin) / 1e3),
           "warp_total_cycles_median_max": [float(np.median(tot)), float(tot.max())],
           "median_warp_phases": {NAMES[i]: float(np.median(D[:, i])) for i in range(7)},
           "max_warp_phases": {NAMES[i]: float(D[:, i].max()) for i in range(7)}}
    print(json.dumps(out))
