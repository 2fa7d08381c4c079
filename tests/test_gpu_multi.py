"""GPU: golden fixtures through the C ABI, and the multi-GPU path (needs >= 2 GPUs: gpurun --gpus 2)."""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import gymnet_b200 as G
from gymnet_b200 import _native as N
from helpers import RTOL, STATE_SCALE, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("fixture,name", [("cartpole", "CartPole-v1"), ("pendulum", "Pendulum-v1"),
                                          ("mountaincar", "MountainCar-v0"), ("mountaincar_cont", "MountainCarContinuous-v0"),
                                          ("acrobot", "Acrobot-v1")])
def test_golden_fixtures_through_the_c_abi(fixture, name):
    g = np.load(os.path.join(GOLD, fixture + ".npz"))
    n = len(g["state"])
    env = G.make(name, n, seed=0, time_limit=-1)
    env.ResetBatch()
    aux = np.zeros((n, 3), np.int32); aux[:, 0] = g["sbd"] if "sbd" in g.files else -1
    env.SetState(g["state"], aux, 0)
    obs, rew, done = env.StepBatch(g["action"])
    st, ax, _ = env.GetState()
    assert np.array_equal(done, g["done"])                                   # bit-exact termination
    if "next_sbd" in g.files:
        assert np.array_equal(ax[:, 0], g["next_sbd"]) and np.array_equal(rew, g["reward"])
    want = g["next_state"].copy(); got = st.astype(np.float64)
    if name == "Acrobot-v1":
        d = got[:, :2] - want[:, :2]; got[:, :2] = want[:, :2] + (d + np.pi) % (2 * np.pi) - np.pi
    assert rel_err(got, want, STATE_SCALE[name]).max() <= RTOL
    env.Close()


def _ngpu():
    c = C.c_int(0)
    return c.value if N.lib().gymcuda_device_count(C.byref(c)) else c.value


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_sharded_rollout_and_nccl_allgather(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "multi_gpu_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), script, str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    gathered = np.load(tmp_path / "gathered.npy")       # [world][n][od] from ncclAllGather on rank 0
    full = G.CartPoleVecEnv(gathered.shape[0] * gathered.shape[1], seed=21, auto_reset=True)
    full.ResetBatch()
    full.RolloutRandom(50, want=())
    assert np.array_equal(gathered.reshape(-1, 4), full.Observe())
    full.Close()
    assert np.load(tmp_path / "fused_ok.npy")[0] == 1.0      # fused peer-store gather == ncclAllGather, 25 steps
    assert np.load(tmp_path / "fused_lunar_ok.npy")[0] == 1.0   # LunarLander: partitioned step + peer push == ncclAllGather, 30 steps
    gl = np.load(tmp_path / "gathered_lunar.npy")
    ll = G.LunarLanderVecEnv(gl.shape[0] * gl.shape[1], seed=5, auto_reset=True)
    ll.ResetBatch()
    for _ in range(20):
        obs, _, _ = ll.StepBatch(np.full(ll.NumberOfEnvironments, 2, np.int32))
    assert np.array_equal(gl.reshape(-1, 8), obs)
    ll.Close()


# ---------------------------------------------------------------- BASELINE.json configs 3-5 at full size:
# size-independent properties of long rollouts (the oracle covers the same paths bit-exactly at small n).
def test_config3_pendulum_and_mountaincar_continuous_full_size():
    n = 262144
    env = G.PendulumVecEnv(n, seed=0, auto_reset=True); env.ResetBatch()
    obs, rew, done, act = env.RolloutRandom(201)
    assert np.abs(obs[..., 0] ** 2 + obs[..., 1] ** 2 - 1.0).max() <= 1e-5      # (cos, sin) of one angle
    assert np.abs(obs[..., 2]).max() <= 8.0                                      # max_speed clamp
    assert rew.max() <= 0.0 and rew.min() >= -(np.pi ** 2 + 0.1 * 64 + 0.001 * 4) - 1e-4
    assert act.min() >= -2.0 and act.max() < 2.0
    assert done[199].all() and done[:199].sum() == 0 and done[200].sum() == 0    # truncation at exactly 200 steps
    env.Close()
    env = G.MountainCarContinuousVecEnv(n, seed=0, auto_reset=True); env.ResetBatch()
    obs, rew, done, act = env.RolloutRandom(64)
    assert obs[..., 0].min() >= -1.2000001 and obs[..., 0].max() <= 0.6 and np.abs(obs[..., 1]).max() <= 0.0700001
    assert np.allclose(rew[done == 0], -0.1 * act[..., 0][done == 0] ** 2, atol=1e-7)
    env.Close()


def test_config4_acrobot_shard_full_size():
    n = 131072     # 1 048 576 envs over 8 GPUs
    env = G.AcrobotVecEnv(n, seed=0, auto_reset=True, env_id_offset=3 * n); env.ResetBatch()
    obs, rew, done, act = env.RolloutRandom(100)
    assert np.abs(obs[..., 0] ** 2 + obs[..., 1] ** 2 - 1.0).max() <= 1e-5
    assert np.abs(obs[..., 2] ** 2 + obs[..., 3] ** 2 - 1.0).max() <= 1e-5
    assert np.abs(obs[..., 4]).max() <= 4 * np.pi + 1e-5 and np.abs(obs[..., 5]).max() <= 9 * np.pi + 1e-5
    assert set(np.unique(rew)) <= {-1.0, 0.0} and np.array_equal(rew == 0.0, done == 1)
    assert set(np.unique(act)) == {0, 1, 2}
    env.Close()


def test_config5_lunarlander_shard_full_size():
    n = 65536      # 524 288 envs over 8 GPUs
    env = G.LunarLanderVecEnv(n, seed=0, auto_reset=True, time_limit=300); env.ResetBatch()
    obs, rew, done, act = env.RolloutRandom(120)
    assert np.isfinite(obs).all() and np.isfinite(rew).all()
    assert set(np.unique(obs[..., 6:])) <= {0.0, 1.0}
    ended = done == 1
    assert ended.sum() > 0 and set(np.unique(rew[ended])) <= {-100.0, 100.0}    # LunarLanderEnv.cs:762-771
    assert env.Stats()["episodes"] == int(ended.sum())
    env.Close()
