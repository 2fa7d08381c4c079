"""Small pass over every kernel for compute-sanitizer (memcheck / racecheck).
   `python tools/sanitize_probe.py lunar` runs only the LunarLander contact loop (used with GYMCUDA_LUNAR_TRIO=1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gymnet_b200 as G
rng = np.random.default_rng(0)
if len(sys.argv) > 1 and sys.argv[1] == "lunar":
    ll = G.LunarLanderVecEnv(1500, seed=2, auto_reset=True, time_limit=300); obs = ll.ResetBatch()
    for _ in range(120):
        obs, r, d = ll.StepBatch(np.full(1500, 0, np.int32))
    print("contacts", float((obs[:, 6:] > 0).any(1).mean()), "trio", os.environ.get("GYMCUDA_LUNAR_TRIO", "0"))
    ll.Close()
    sys.exit(0)
for name, n in (("CartPole-v1", 3001), ("Pendulum-v1", 1025), ("MountainCar-v0", 515), ("MountainCarContinuous-v0", 515),
                ("Acrobot-v1", 777), ("LunarLander-v2", 2500)):
    for auto in (False, True):
        env = G.make(name, n, seed=1, auto_reset=auto, episode_stats=True, done_bits=True)
        env.ResetBatch()
        env.RolloutRandom(24)
        env.RolloutRandom(5, want=("done",))
        for _ in range(6):
            a = env.SampleActions()
            env.StepBatch(a)
            env.DoneIndices()
        if env.act_n > 0:
            env.SampleActions(mask=(rng.random((n, env.act_n)) < 0.5).astype(np.uint8))
            env.Step(1)
        m = (rng.random(n) < 0.3).astype(np.uint8)
        env.ResetBatch(mask=m)
        st, ax, t = env.GetState(); env.SetState(st, ax, t); env.Observe(); env.Stats(reset=True)
        env.Close()
# the ALL_OUT rollout variant (no statistics): 8-step chunks, 32-bit row index, staged 3/6-float observation stores
for name, n in (("CartPole-v1", 3000), ("Pendulum-v1", 1028), ("Pendulum-v1", 1030), ("MountainCar-v0", 516),
                ("MountainCarContinuous-v0", 516), ("Acrobot-v1", 772), ("Acrobot-v1", 70), ("LunarLander-v2", 512)):
    env = G.make(name, n, seed=1, auto_reset=True)
    env.ResetBatch()
    env.RolloutRandom(3)
    env.RolloutRandom(29)
    env.Close()
ll = G.LunarLanderVecEnv(2500, seed=2, auto_reset=True, time_limit=300); obs = ll.ResetBatch()
for _ in range(120):      # long enough for contacts: exercises the contact partition + solver paths
    obs, r, d = ll.StepBatch(np.full(2500, 0, np.int32))
print("contacts", float((obs[:, 6:] > 0).any(1).mean()))
ll.Close()
# round-2 entry points: terminal observations, step_many (valid + rejected actions), the done list built on demand, Box.Sample,
# the device-resident clock with a captured + replayed step
import torch
for name, n in (("CartPole-v1", 3001), ("Acrobot-v1", 777), ("LunarLander-v2", 700)):
    env = G.make(name, n, seed=4, auto_reset=True, time_limit=20)
    env.ResetBatch()
    term = np.zeros((n, env.obs_dim), np.float32); env.SetTerminalObs(term)
    acts = rng.integers(0, env.act_n, (30, n)).astype(np.int32)
    env.StepMany(acts)
    if name != "CartPole-v1":
        acts[3, 5] = 99
        try:
            env.StepMany(acts[:6])
        except G.InvalidActionError:
            pass
    for _ in range(4):
        env.StepBatch(acts[0]); env.DoneIndices(); env.DoneIndices()
    env.SetTerminalObs(None)
    dev = torch.device("cuda", 0); s = torch.cuda.Stream(device=dev); env.SetStream(s.cuda_stream); env.SetDeviceClock(True)
    a = torch.zeros(n, dtype=torch.int32, device=dev); o = torch.empty((n, env.obs_dim), device=dev); r = torch.empty(n, device=dev); d = torch.empty(n, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(s):
        env.StepDevice(a.data_ptr(), o.data_ptr(), r.data_ptr(), d.data_ptr())
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        env.StepDevice(a.data_ptr(), o.data_ptr(), r.data_ptr(), d.data_ptr())
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize(); env.DoneIndices(); env.Stats(); env.SetDeviceClock(False)
    env.Close()
# last session of round 2: step_many in the one-wave 512-thread shape from an unaligned step index (head, chunks with the next chunk's
# actions prefetched, tail), a host call long enough for several pipelined chunks, rollouts through the same pipeline, no-observation-copy steps
env = G.make("CartPole-v1", 57344 + 40, seed=5, auto_reset=True); env.ResetBatch()
env.StepBatch(np.zeros(57344 + 40, np.int32))
env.StepMany(rng.integers(0, 2, (45, 57344 + 40)).astype(np.int32))
env.RolloutRandom(40)
dev = torch.device("cuda", 0)
a = torch.zeros(57344 + 40, dtype=torch.int32, device=dev); r = torch.empty(57344 + 40, device=dev); d = torch.empty(57344 + 40, dtype=torch.uint8, device=dev)
for _ in range(3):
    env.StepDevice(a.data_ptr(), env.NO_OBS, r.data_ptr(), d.data_ptr())
env.Sync(); env.ObsViewDevice(); env.Observe(); env.Close()
env = G.make("MountainCar-v0", 9000, seed=5, auto_reset=True, time_limit=11); env.ResetBatch()
acts = rng.integers(0, 3, (300, 9000)).astype(np.int32); acts[17, 100] = 5
try:
    env.StepMany(acts)        # 300 x 9000 x 17 B = 46 MB: three pipelined chunks, one rejected action
except G.InvalidActionError:
    pass
env.Close()
G.Box(np.array([-2.0, 1.5, -np.inf, -np.inf], np.float32), np.array([3.0, np.inf, -4.0, np.inf], np.float32)).SampleBatch(5000, seed=1)
print("sanitize probe done")
