#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched CartPole hot path on N B200s (one process per GPU).

  python bench.py --gpus 1 --steps K --warmup W
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...        the reference's CPU path (CPU oracle, all host cores)

A bench "step" is ONE launch of the fused random-policy rollout kernel: `--inner` env steps of every
one of the `--num-envs` envs of this GPU (BASELINE.json configs[1]: CartPole-v1, 65 536 envs, random
policy), with the whole trajectory (obs, reward, done, action = 25 B per env step) streamed to HBM.
One step writes num_envs * inner * 25 B = 839 MB >> the 126 MB L2, so no L2 flush is needed.
The timed region is K launches right after the warm-up (one kernel timed alone: the regime of the burst copy
behind MEASURED_PEAKS.json); `roofline.sustained` repeats the measurement after >= 0.6 s of continuous launches,
next to a memset and a copy timed live in the same state.
`e2e` is the same metric, same bench step (`--inner` env steps of every env), through the host-buffer C-ABI call
gymcuda_step_many with pinned HOST buffers: the caller's actions travel in and obs / reward / done travel out inside the
timed region, every step; `e2e.per_step_call` is the one-env-step-per-call figure (gymcuda_step, the reference's Step()).
After the headline, the same process measures the other BASELINE.json configs and appends them as `envs` to the one JSON
line (SURVEY 8d configs 3-5 and the L2-busting 16 777 216-env CartPole step run), each with the roofline that binds it.
Prints exactly one JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_ROLLOUT = 25       # obs 16 + reward 4 + done 1 + action 4 written per env step
ALGO_BYTES_STATE = 32         # state read + written once per launch, per env
METRIC = "env-steps/sec"
FALLBACK_HBM_GBS = 6650.0     # B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gymcuda", choices=["gymcuda", "reference"])
    ap.add_argument("--env", default="CartPole-v1")
    ap.add_argument("--num-envs", type=int, default=65536, help="envs per GPU (weak scaling)")
    ap.add_argument("--inner", type=int, default=512, help="env steps per rollout launch")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-envs", action="store_true", help="skip the `envs` block (the other BASELINE.json configs)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "how": "nvidia-smi -lms 50 from before the warm-up, through the timed region, to >= 0.6 s of the same "
                       "rollout launches right after it (the timed region itself lasts a few ms: less than one sample)"}


def _oracle_setup(env_name, n, threads, chunk):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from helpers import KINDS
    o = O.OracleEnv(KINDS[env_name], n, seed=0, auto_reset=True, mode=O.MODE_F64)
    o.set_threads(threads)
    o.reset()
    rng = np.random.default_rng(0)
    d = o.d
    acts = (rng.integers(0, d["act_n"], (chunk, n)).astype(np.int32) if d["act_n"] > 0
            else rng.uniform(-1, 1, (chunk, n, d["act_dim"])).astype(np.float32))
    return o, acts


def oracle_cpu_rate(env_name, n, seconds, threads, chunk=32):
    """The reference's CPU path restated (oracle F64 = the C# double arithmetic of CartPoleEnv.Step),
    per-env serial loops spread over `threads` host threads (the DistributedScheduler pool shape),
    random actions pre-generated.  Returns (env-steps/s, sample text)."""
    o, acts = _oracle_setup(env_name, n, threads, chunk)
    o.step_many(acts)
    calls, t0 = 0, time.perf_counter()
    while True:
        o.step_many(acts); calls += 1
        el = time.perf_counter() - t0
        if el >= seconds or calls >= 100000:
            break
    return n * chunk * calls / el, "%s, %d envs x %d steps in %.1f s, oracle F64 port (C# double arithmetic), %d threads" % (
        env_name, n, chunk * calls, el, threads)


def workload_config(args):
    """The `config` of BOTH arms (the driver compares them): the workload, not how an arm executes it."""
    n, K = args.num_envs, args.inner
    bytes_per_step = n * K * 25 if args.env == "CartPole-v1" else None
    cfg = {"workload": "%s, %d envs per GPU, random policy, auto-reset; one step = %d env steps of every env" % (args.env, n, K),
           "num_envs_per_gpu": n, "inner_steps": K, "parallelism": "independent env shards, no collective"}
    if bytes_per_step:
        cfg["l2"] = "outputs per step (%.0f MB) exceed the 126 MB L2; no flush needed" % ((bytes_per_step + n * 32) / 1e6)
    return cfg


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is C#
    (no dotnet/mono in this image) so oracle/_ref cannot be built; the CPU oracle port (C++ restatement of
    CartPoleEnv.Step in the C#'s double arithmetic) is timed on all host cores.  One bench step = a bounded sample:
    `inner` batched steps of all num_envs envs, repeated until the step has lasted >= 3 s / steps (so the whole timed
    region is >= 3 s whatever --steps is: a 0.1 s sample disagreed by 30 % with the 10 s one in round 1)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n, inner = args.num_envs, 32
    o, acts = _oracle_setup(args.env, n, cores, inner)
    for w in range(max(1, args.warmup)):
        o.step_many(acts)
    per_step_s = max(3.0 / max(1, args.steps), 0.05)
    calls = 0
    t0 = time.perf_counter()
    for s in range(args.steps):
        ts = time.perf_counter()
        while True:
            o.step_many(acts); calls += 1
            if time.perf_counter() - ts >= per_step_s:
                break
    el = time.perf_counter() - t0
    value = n * inner * calls / el
    one_core, one_sample = oracle_cpu_rate(args.env, min(n, 4096), 2.0, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "reference_sample_per_step": "batched steps of all %d envs for >= %.2f s (%d batched steps in total)" % (n, per_step_s, inner * calls),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d envs x %d steps in %.1f s, C++ port (oracle F64) of CartPoleEnv.Step -- the C# itself cannot run here: no dotnet" % (n, inner * calls, el),
                         "one_core": {"value": one_core, "sample": one_sample}},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def measure_gather(env, torch, dist, dev, rank, world, n, od, ad, t_act, steps=200):
    import gymnet_b200 as G
    ids = [G.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    env.CommInit(ids[0], rank, world)
    handles = [None] * world
    dist.all_gather_object(handles, env.GatherCreate(rank, world))
    env.GatherOpen(handles)
    a1 = t_act[0].contiguous()
    d_obs = torch.empty((n, od), dtype=torch.float32, device=dev)
    d_rew = torch.empty((n,), dtype=torch.float32, device=dev)
    d_done = torch.empty((n,), dtype=torch.uint8, device=dev)
    out = torch.empty((world, n, od), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream()

    def timed(fn):
        for _ in range(10):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        dist.barrier(); torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) * 1e3   # us per step, max over ranks

    def plain():
        env.StepDevice(a1.data_ptr(), d_obs.data_ptr(), d_rew.data_ptr(), d_done.data_ptr())

    def nccl():
        env.StepDevice(a1.data_ptr(), d_obs.data_ptr(), d_rew.data_ptr(), d_done.data_ptr())
        env.AllGatherObs(out.data_ptr(), d_obs.data_ptr())

    def fused():
        env.StepGatherDevice(a1.data_ptr(), d_rew.data_ptr(), d_done.data_ptr())
        env.GatherWait()

    return {"unit": "us per step of %d envs per GPU, max over ranks" % n, "step_only": timed(plain),
            "step_then_ncclAllGather": timed(nccl), "step_fused_p2p_gather": timed(fused),
            "gathered_bytes_per_rank": world * n * od * 4}



# warp-instructions issued per env step of the fused rollout kernel (ncu smsp__inst_executed / env steps, profiles/):
# the FP32-issue roofline of an env whose step is arithmetic-bound = thread-instructions/s over lanes x SMSPs x clock
INSTR_PER_ENV_STEP = {"Acrobot-v1": 500.0}


def measure_envs(G, torch, dist, dev, rank, world, local_rank, peak, sm_mhz):
    """SURVEY 8d configs 3-5 + config 2's L2-busting per-launch run, measured after the headline in the same process.
    Every number is a whole-job aggregate (all ranks, max over ranks of the device time)."""
    out = {}
    stream = torch.cuda.current_stream()

    def timed(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms

    def rollout_case(name, n, K, reps, **kw):
        env = G.make(name, n, seed=0, device=local_rank, env_id_offset=rank * n, auto_reset=True, **kw)
        env.SetStream(stream.cuda_stream)
        env.ResetBatch()
        od, ad = env.obs_dim, env.act_dim
        o = torch.empty((K, n, od), dtype=torch.float32, device=dev); r = torch.empty((K, n), dtype=torch.float32, device=dev)
        d = torch.empty((K, n), dtype=torch.uint8, device=dev); a = torch.empty((K, n, ad), dtype=torch.int32 if env.act_n > 0 else torch.float32, device=dev)
        ms = timed(lambda: env.RolloutRandomDevice(K, o.data_ptr(), r.data_ptr(), d.data_ptr(), a.data_ptr()), reps)
        rate = world * n * K / (ms * 1e-3)
        b = od * 4 + 4 + 1 + ad * 4
        res = {"mode": "fused rollout, %d env steps per launch" % K, "num_envs_per_gpu": n, "env_steps_per_s": rate, "ms_per_launch": ms,
               "algorithmic_bytes_per_env_step": b, "hbm_gbs_per_gpu": rate / world * b / 1e9, "hbm_frac": rate / world * b / 1e9 / peak}
        env.Close()
        del o, r, d, a
        return res

    # config 3: Pendulum-v1 + MountainCarContinuous-v0, 262 144 envs, continuous Box actions -- HBM-write bound
    for name in ("Pendulum-v1", "MountainCarContinuous-v0"):
        res = rollout_case(name, 262144, 128, 10)
        res["bound"] = "hbm"; res["frac"] = res["hbm_frac"]
        out[name] = res
    # config 4: Acrobot-v1, 131 072 envs per GPU (1 048 576 at 8 GPUs), independent shards -- FP32-issue bound
    res = rollout_case("Acrobot-v1", 131072, 128, 10)
    lanes_per_s = 148 * 4 * 32 * (sm_mhz or 1965.0) * 1e6
    res["bound"] = "fp32_issue"
    res["instr_per_env_step"] = INSTR_PER_ENV_STEP["Acrobot-v1"]
    res["achieved_thread_instr_per_s_per_gpu"] = res["env_steps_per_s"] / world * INSTR_PER_ENV_STEP["Acrobot-v1"]
    res["peak_thread_instr_per_s_per_gpu"] = lanes_per_s
    res["frac"] = res["achieved_thread_instr_per_s_per_gpu"] / lanes_per_s
    res["note"] = "RK4 of the book dynamics: 8 sincos + ~150 flop per step; one warp-instruction per cycle and sub-partition is the ceiling (148 SMs x 4 x 32 lanes x SM clock)"
    out["Acrobot-v1"] = res
    # config 4, the other reading of SURVEY 8d: N TOTAL fixed (1 048 576 envs split over the GPUs: strong scaling)
    strong = rollout_case("Acrobot-v1", 1048576 // world, 128, 5)
    strong["mode"] += "; 1 048 576 envs in TOTAL, %d per GPU (strong scaling)" % (1048576 // world)
    strong["bound"] = "fp32_issue"
    strong["frac"] = strong["env_steps_per_s"] / world * INSTR_PER_ENV_STEP["Acrobot-v1"] / lanes_per_s
    out["Acrobot-v1 strong @1048576 total"] = strong

    # config 5: LunarLander-v2, 65 536 landers per GPU (524 288 at 8), auto-reset + done compaction, one launch group per step
    n = 65536
    env = G.make("LunarLander-v2", n, seed=0, device=local_rank, env_id_offset=rank * n, auto_reset=True, time_limit=1000)
    env.SetStream(stream.cuda_stream)
    env.ResetBatch()
    o = torch.empty((n, 8), dtype=torch.float32, device=dev); r = torch.empty((n,), dtype=torch.float32, device=dev); d = torch.empty((n,), dtype=torch.uint8, device=dev)
    acts = torch.randint(0, 4, (n,), dtype=torch.int32, device=dev)
    step = lambda: env.StepDevice(acts.data_ptr(), o.data_ptr(), r.data_ptr(), d.data_ptr())   # noqa: E731
    for _ in range(150):   # from the reset state (every lander in free flight at the top) to the steady mix of flight / touch-down / reset
        step()
    ms = timed(step, 100, warm=0)
    sb = 2 * (env.state_dim + env.aux_dim) * 4 + 4 + 32 + 4 + 1
    lunar = {"mode": "one step per launch group (partition + free-flight kernel + contact kernel), device-resident actions", "num_envs_per_gpu": n,
             "env_steps_per_s": world * n / (ms * 1e-3), "ms_per_step": ms, "bound": "fp32_latency",
             "algorithmic_bytes_per_env_step": sb, "hbm_frac": n / (ms * 1e-3) * sb / 1e9 / peak,
             "legs_down_frac": float((o[:, 6:] > 0).any(1).float().mean().item()),
             "note": "180 velocity + up to 60 position iterations of sequential impulses per lander: the step lasts as long as the dependent float32 chain of "
                     "the slowest lander in contact; HBM is idle (hbm_frac)"}
    if world > 1:
        try:
            g = measure_gather(env, torch, dist, dev, rank, world, n, 8, 1, acts.view(1, n, 1), steps=50)
            lunar["obs_allgather"] = g
            lunar["env_steps_per_s_with_ncclAllGather"] = world * n / (g["step_then_ncclAllGather"] * 1e-6)
            lunar["env_steps_per_s_with_fused_gather"] = world * n / (g["step_fused_p2p_gather"] * 1e-6)
        except Exception as ex:
            lunar["obs_allgather"] = {"error": repr(ex)[:300]}
    env.Close()
    del o, r, d, acts
    out["LunarLander-v2"] = lunar

    # config 2, per-launch mode, L2-busting: CartPole step_kernel at 16 777 216 envs (state 256 MiB >> the 126 MB L2), 41 B per env step
    n = 16777216
    env = G.make("CartPole-v1", n, seed=0, device=local_rank, env_id_offset=0, auto_reset=True)
    env.SetStream(stream.cuda_stream)
    env.ResetBatch()
    o = torch.empty((n, 4), dtype=torch.float32, device=dev); r = torch.empty((n,), dtype=torch.float32, device=dev); d = torch.empty((n,), dtype=torch.uint8, device=dev)
    acts = torch.randint(0, 2, (n,), dtype=torch.int32, device=dev)
    for _ in range(40):   # past the first episode ends (random-policy episodes last ~22 steps): the steady 4.5 % of finished envs per step
        env.StepDevice(acts.data_ptr(), o.data_ptr(), r.data_ptr(), d.data_ptr())
    ms57 = timed(lambda: env.StepDevice(acts.data_ptr(), o.data_ptr(), r.data_ptr(), d.data_ptr()), 20)
    # d_obs = GYMCUDA_NO_OBS: CartPole's observation is its state (CartPoleEnv.cs:183), read in place through gymcuda_obs_view_device
    ms = timed(lambda: env.StepDevice(acts.data_ptr(), env.NO_OBS, r.data_ptr(), d.data_ptr()), 20)
    rate, rate57 = n / (ms * 1e-3), n / (ms57 * 1e-3)
    # state 16 R + 16 W, action 4 R, reward 4 W, done 1 W = 41; a separate observation copy adds 16 W
    out["CartPole-v1 step_kernel @16777216"] = {
        "mode": "one step per launch (gymcuda_step_device, d_obs = GYMCUDA_NO_OBS: observations read in place), per GPU, steady state (4.5 % of the envs finish per step)", "num_envs_per_gpu": n,
        "env_steps_per_s_per_gpu": rate, "ms_per_launch": ms, "bound": "hbm",
        "algorithmic_bytes_per_env_step": 41, "hbm_gbs_per_gpu": rate * 41 / 1e9, "frac": rate * 41 / 1e9 / peak,
        "with_obs_copy": {"env_steps_per_s_per_gpu": rate57, "ms_per_launch": ms57, "bytes_moved_per_env_step": 57,
                          "hbm_gbs_per_gpu": rate57 * 57 / 1e9, "frac": rate57 * 57 / 1e9 / peak}}
    env.Close()
    del o, r, d, acts

    # caller-supplied actions, k steps per launch (gymcuda_step_many_device): what a device-resident learner that produces
    # blocks of actions gets instead of k per-launch steps -- CartPole, 65 536 envs, 512 steps per launch
    n, K = 65536, 512
    env = G.make("CartPole-v1", n, seed=0, device=local_rank, env_id_offset=rank * n, auto_reset=True)
    env.SetStream(stream.cuda_stream)
    env.ResetBatch()
    o = torch.empty((K, n, 4), dtype=torch.float32, device=dev); r = torch.empty((K, n), dtype=torch.float32, device=dev)
    d = torch.empty((K, n), dtype=torch.uint8, device=dev)
    acts = torch.randint(0, 2, (K, n), dtype=torch.int32, device=dev)
    ms = timed(lambda: env.StepManyDevice(K, acts.data_ptr(), o.data_ptr(), r.data_ptr(), d.data_ptr()), 10)
    rate = world * n * K / (ms * 1e-3)
    out["CartPole-v1 step_many @65536"] = {
        "mode": "gymcuda_step_many_device: %d env steps per launch with caller-supplied device-resident actions" % K, "num_envs_per_gpu": n,
        "env_steps_per_s": rate, "ms_per_launch": ms, "bound": "hbm", "algorithmic_bytes_per_env_step": 25,
        "hbm_gbs_per_gpu": rate / world * 25 / 1e9, "frac": rate / world * 25 / 1e9 / peak,
        "note": "action 4 B read + obs 16 B + reward 4 B + done 1 B written per env step; the per-launch gymcuda_step_device figure of the same batch is in profiles/"}
    env.Close()
    del o, r, d, acts

    # per-launch steps without per-launch host cost: 16 gymcuda_step_device calls captured into ONE CUDA graph
    # (gymcuda_set_device_clock makes the captured steps replayable), device-resident actions -- CartPole, 65 536 envs
    n, G16 = 65536, 16
    env = G.make("CartPole-v1", n, seed=0, device=local_rank, env_id_offset=rank * n, auto_reset=True)
    env.SetStream(stream.cuda_stream)
    env.ResetBatch()
    env.SetDeviceClock(True)
    o = torch.empty((n, 4), dtype=torch.float32, device=dev); r = torch.empty((n,), dtype=torch.float32, device=dev); d = torch.empty((n,), dtype=torch.uint8, device=dev)
    acts = torch.randint(0, 2, (n,), dtype=torch.int32, device=dev)
    step1 = lambda: env.StepDevice(acts.data_ptr(), o.data_ptr(), r.data_ptr(), d.data_ptr())   # noqa: E731
    ms_plain = timed(step1, 200)
    side = torch.cuda.Stream(device=dev)
    env.SetStream(side.cuda_stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(G16):
            step1()
    env.SetStream(stream.cuda_stream)
    ms_graph = timed(graph.replay, 50) / G16
    out["CartPole-v1 step via CUDA graph @65536"] = {
        "mode": "gymcuda_step_device captured %d times into one CUDA graph and replayed (device-resident step clock)" % G16, "num_envs_per_gpu": n,
        "us_per_step_graph": ms_graph * 1e3, "us_per_step_plain_launch": ms_plain * 1e3, "env_steps_per_s": world * n / (ms_graph * 1e-3),
        "bound": "launch_latency", "note": "one dependent chain of load -> step -> store -> counters per step; the working set lives in L2"}
    env.SetDeviceClock(False)
    env.Close()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gymnet_b200 as G

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libgymcuda has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # one process per GPU: run on the cores of that GPU's NUMA node, so that the page-locked host buffers of the
    # e2e leg (first-touched below) sit next to the GPU that writes them
    host_cpus = G.bind_host_to_device(local_rank) if world > 1 else None
    n, K = args.num_envs, args.inner
    env = G.make(args.env, n, seed=0, device=local_rank, env_id_offset=rank * n, auto_reset=True)
    od, ad = env.obs_dim, env.act_dim
    stream = torch.cuda.Stream(device=dev)   # a real (non-NULL) stream: launches and events share it
    torch.cuda.set_stream(stream)
    env.SetStream(stream.cuda_stream)
    env.ResetBatch()

    # trajectory buffers live in HBM (torch is only the allocator here)
    t_obs = torch.empty((K, n, od), dtype=torch.float32, device=dev)
    t_rew = torch.empty((K, n), dtype=torch.float32, device=dev)
    t_done = torch.empty((K, n), dtype=torch.uint8, device=dev)
    t_act = torch.empty((K, n, ad), dtype=torch.int32 if env.act_n > 0 else torch.float32, device=dev)

    def launch():
        env.RolloutRandomDevice(K, t_obs.data_ptr(), t_rew.data_ptr(), t_done.data_ptr(), t_act.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The timed region is K launches right after the warm-up: a few milliseconds of one kernel timed alone, the same
    # regime as the burst copy behind MEASURED_PEAKS.json's hbm_gbs.  nvidia-smi cannot resolve a region that short
    # (one sample per 50 ms), so the sampler runs from before the warm-up until the same launches have kept going
    # for >= 0.6 s AFTER the timed region; that tail is the sustained regime (the 1 kW power cap engages within a
    # few hundred ms of continuous 5.6 TB/s writes), timed again and reported beside a live memset / copy.
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        launch()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for s in range(args.steps):
        launch()
        evs[s + 1].record(stream)
    barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    t_pre = time.perf_counter()
    while True:
        for _ in range(16):
            launch()
        torch.cuda.synchronize()
        el = time.perf_counter() - t_pre
        # at least 0.6 s under load; on a busy 8-GPU box nvidia-smi needs longer to deliver its first rows
        if (el >= 0.6 and len(sampler.rows) >= 6) or el >= 4.0:
            break
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for s in range(args.steps):
        launch()
    s1.record(stream)
    torch.cuda.synchronize()
    sustained_ms = s0.elapsed_time(s1) / args.steps
    clocks = sampler.stop()
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = world * n * K * args.steps / (total_ms * 1e-3)

    # ---- sanity: the timed launches really produced a trajectory
    torch.cuda.synchronize()
    dsum = int(t_done.sum().item())
    assert 0 < dsum < K * n, "rollout produced no episode boundaries"

    # ---- roofline of the dominant kernel (rollout_kernel), measured live with CUDA events
    peak, peak_src = measured_peak()
    launch_bytes = n * K * (od * 4 + 4 + 1 + ad * 4) + n * ALGO_BYTES_STATE
    avg_launch_s = float(np.mean(per_launch_ms)) * 1e-3
    achieved = launch_bytes / avg_launch_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "rollout_kernel<%s>" % args.env,
                "algorithmic_bytes_per_launch": launch_bytes,
                "note": "bytes = env-steps x (obs+reward+done+action) + 32 B/env state; traffic from ncu --set full in profiles/"}
    # the same launches after >= 0.6 s of continuous load, next to what plain memset / copy sustain right then
    fill = torch.empty(launch_bytes // 4, dtype=torch.float32, device=dev)
    src = torch.empty(launch_bytes // 8, dtype=torch.float32, device=dev)
    dst = torch.empty_like(src)

    def sustained_gbs(fn, nbytes, seconds=0.25):
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(8):
                fn()
            torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(10):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        return nbytes * 10 / (a.elapsed_time(b) * 1e-3) / 1e9

    memset_gbs = sustained_gbs(lambda: fill.zero_(), fill.numel() * 4)
    copy_gbs = sustained_gbs(lambda: dst.copy_(src), src.numel() * 8)
    del fill, src, dst
    sus_gbs = launch_bytes / (sustained_ms * 1e-3) / 1e9
    roofline["sustained"] = {"ms_per_step": sustained_ms, "achieved": sus_gbs, "frac_of_peak": sus_gbs / peak,
                             "live_memset_gbs": memset_gbs, "live_copy_gbs": copy_gbs,
                             "frac_of_live_memset": sus_gbs / memset_gbs,
                             "note": "after >= 0.6 s of back-to-back launches (power cap engaged); memset = write-only "
                                     "like this kernel, copy = read + write bytes, both torch kernels timed the same way"}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("rollout_dram_bytes_per_launch")
        except Exception:
            pass

    # ---- host buffers through gymcuda_step (H2D actions, kernel, D2H obs/reward/done), one env step per call
    e2e_steps = args.e2e_steps
    h_act = torch.empty((n, ad), dtype=t_act.dtype).pin_memory()
    # obs | reward | done adjacent in one pinned block (what the C# shim pins): the library then needs one DMA
    h_out = torch.empty((n * od * 4 + n * 4 + n,), dtype=torch.uint8).pin_memory()
    h_obs = h_out[: n * od * 4].view(torch.float32).view(n, od)
    h_rew = h_out[n * od * 4: n * od * 4 + n * 4].view(torch.float32)
    h_done = h_out[n * od * 4 + n * 4:]
    rng = np.random.default_rng(rank)
    if env.act_n > 0:
        h_act.numpy()[:] = rng.integers(0, env.act_n, (n, ad))
    else:
        h_act.numpy()[:] = rng.uniform(-1, 1, (n, ad))
    L = G._native.lib()

    # the four buffer arguments are marshalled once: a C# caller pins its arrays once too, and the per-call cost of
    # building ctypes objects would otherwise be timed as part of the library
    e2e_args = (env._h, C.c_void_p(h_act.data_ptr()), C.c_void_p(h_obs.data_ptr()),
                C.c_void_p(h_rew.data_ptr()), C.c_void_p(h_done.data_ptr()))
    step_fn = L.gymcuda_step

    def e2e_step():
        if step_fn(*e2e_args) != 0:
            raise RuntimeError(L.gymcuda_last_error())

    for _ in range(5):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    per_call = {"value": world * n * e2e_steps / e2e_s, "unit": "env-steps/s",
                "h2d_bytes_per_call": n * ad * 4, "d2h_bytes_per_call": n * (od * 4 + 4 + 1),
                "api": "gymcuda_step (host buffers, pinned): one env step of every env per call", "calls": e2e_steps}

    # ---- e2e, the bench's own step: K env steps of every env per call through gymcuda_step_many with pinned HOST buffers --
    # the caller's actions [K][n] travel in, obs / reward / done [K][n] travel out, every call; inside the call the action
    # chunks, the chunk launches and the trajectory chunks are pipelined over PCIe's two directions (gymcuda.cu host_k_steps)
    K = args.inner
    alloc_error = None
    try:
        hm_act = torch.empty((K, n, ad), dtype=t_act.dtype).pin_memory()
        hm_obs = torch.empty((K, n, od), dtype=torch.float32).pin_memory()
        hm_rew = torch.empty((K, n), dtype=torch.float32).pin_memory()
        hm_done = torch.empty((K, n), dtype=torch.uint8).pin_memory()
    except Exception as ex:   # the 840 MB of pinned host memory per rank could not be had
        alloc_error = repr(ex)[:300]
    if world > 1:   # every rank takes the same path (the timed loop below holds barriers)
        flag = torch.tensor([0 if alloc_error is None else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()) and alloc_error is None:
            alloc_error = "another rank could not allocate its pinned host buffers"
    if alloc_error is None:
        if env.act_n > 0:
            hm_act.numpy()[:] = rng.integers(0, env.act_n, (K, n, ad))
        else:
            hm_act.numpy()[:] = rng.uniform(-1, 1, (K, n, ad))
        many_args = (env._h, C.c_int(K), C.c_void_p(hm_act.data_ptr()), C.c_void_p(hm_obs.data_ptr()),
                     C.c_void_p(hm_rew.data_ptr()), C.c_void_p(hm_done.data_ptr()))

        def e2e_many():
            if L.gymcuda_step_many(*many_args) != 0:
                raise RuntimeError(L.gymcuda_last_error())

        for _ in range(max(3, args.warmup)):
            e2e_many()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_many()
        barrier()
        many_s = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([many_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            many_s = float(tt.item())
        h2d, d2h = K * n * ad * 4, K * n * (od * 4 + 4 + 1)
        # the PCIe ceiling of that call, live: one bare device -> pinned-host DMA of the size of its observations
        d_src = torch.empty((K, n, od), dtype=torch.float32, device=dev)
        hm_obs.copy_(d_src, non_blocking=True); torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(); hm_obs.copy_(d_src, non_blocking=True); c1.record(); torch.cuda.synchronize()
        dma_gbs = d_src.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        del d_src
        e2e = {"value": world * n * K * args.steps / many_s, "unit": "env-steps/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "gymcuda_step_many (host buffers, pinned): %d env steps of every env per call = one bench step, caller-supplied actions" % K,
               "steps": args.steps, "ms_per_step": many_s / args.steps * 1e3,
               "pcie_gbs_per_gpu": {"h2d": h2d * args.steps / many_s / 1e9, "d2h": d2h * args.steps / many_s / 1e9,
                                    "bare_d2h_dma_live": dma_gbs, "d2h_frac_of_bare_dma": d2h * args.steps / many_s / 1e9 / dma_gbs},
               "per_step_call": per_call,
               "host_affinity": ("%d cores of the GPU's NUMA node" % len(host_cpus)) if host_cpus else "unchanged"}
        del hm_act, hm_obs, hm_rew, hm_done
        # Two host-buffer call shapes were timed; which is faster depends on where the bytes land on the host: the per-call
        # buffers (1.4 MB per rank, rewritten every call) stay in the CPU's last-level cache, the [K][n] trajectories
        # (705 MB per rank and call) go to host DRAM -- one or two ranks get the full PCIe rate for them (52 GB/s), four and
        # more share what the host's memory path takes (~80 GB/s in all on the 8-GPU boxes of this pool).  `e2e` is the
        # faster shape at this rank count; both are in the line ("pick the fastest per topology").
        if per_call["value"] > e2e["value"]:
            many = {k: e2e[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "api", "steps", "ms_per_step", "pcie_gbs_per_gpu")}
            e2e = {"value": per_call["value"], "unit": "env-steps/s", "h2d_bytes_per_step": per_call["h2d_bytes_per_call"],
                   "d2h_bytes_per_step": per_call["d2h_bytes_per_call"], "api": per_call["api"], "steps": e2e_steps,
                   "chosen": "per_step_call (faster than step_many at %d ranks on this host)" % world,
                   "per_step_call": per_call, "step_many": many, "host_affinity": e2e["host_affinity"]}
        else:
            e2e["chosen"] = "step_many (faster than per_step_call at %d rank(s) on this host)" % world
    else:   # the per-call figure stands in, and says so
        e2e = dict(per_call, h2d_bytes_per_step=per_call["h2d_bytes_per_call"], d2h_bytes_per_step=per_call["d2h_bytes_per_call"],
                   steps=e2e_steps, step_many_error=alloc_error,
                   host_affinity=("%d cores of the GPU's NUMA node" % len(host_cpus)) if host_cpus else "unchanged")

    # ---- N > 1 only, outside the headline timing: what the optional observation all-gather costs per step,
    # with NCCL after the step kernel and fused into it as NVLink peer stores (gymcuda_step_gather_device)
    gather = None
    if world > 1:
        try:
            gather = measure_gather(env, torch, dist, dev, rank, world, n, od, ad, t_act)
        except Exception as ex:   # the optional collective must never cost the headline line
            gather = {"error": repr(ex)[:300]}

    # ---- the other BASELINE.json configs (SURVEY 8d configs 3-5, config 2's 16 M-env per-launch run)
    envs = None
    if not args.no_envs:
        env.Close()
        env = None
        del t_obs, t_rew, t_done, t_act
        torch.cuda.empty_cache()
        try:
            envs = measure_envs(G, torch, dist, dev, rank, world, local_rank, peak, clocks.get("sm_mhz"))
        except Exception as ex:   # never costs the headline line
            envs = {"error": repr(ex)[:400]}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port of the reference's CPU path
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, sample = oracle_cpu_rate(args.env, n, args.cpu_seconds, cores)
        v1, sample1 = oracle_cpu_rate(args.env, min(n, 4096), 3.0, 1)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample,
               "one_core": {"value": v1, "sample": sample1}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "execution": "one step = ONE fused rollout launch (in-kernel Philox policy, state in registers, trajectory streamed to HBM)",
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "obs_allgather": gather, "envs": envs,
            "gpu_launches": args.steps, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if env is not None:
        env.Close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
