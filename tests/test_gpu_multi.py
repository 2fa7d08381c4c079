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
