"""CPU: the C-ABI library loads and exports every symbol include/gymcuda.h declares; without a GPU
the product fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import pytest

import gymnet_b200 as G
from gymnet_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gymcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gymcuda_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    syms = header_symbols()
    for must in ("gymcuda_create", "gymcuda_destroy", "gymcuda_seed", "gymcuda_seed_each", "gymcuda_reset",
                 "gymcuda_reset_masked", "gymcuda_step", "gymcuda_step_device", "gymcuda_rollout_random",
                 "gymcuda_done_indices", "gymcuda_get_state", "gymcuda_set_state", "gymcuda_allgather_obs",
                 "gymcuda_sync", "gymcuda_last_error", "gymcuda_version"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(N.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), "libgymcuda.so does not export %s" % name


def test_python_binding_covers_the_header():
    assert sorted(N.SYMBOLS) == header_symbols()


def test_version_and_config_default():
    L = N.lib()
    assert L.gymcuda_version() == 111
    cfg = N.Config()
    assert L.gymcuda_config_default(C.byref(cfg), N.LUNARLANDER, 8) == 0
    assert cfg.struct_size == C.sizeof(N.Config)
    assert (cfg.gravity, cfg.wind_power, cfg.turbulence_power) == (-10.0, 15.0, 1.5)   # LunarLanderEnv.cs:351-354


def test_bad_arguments_are_reported_not_thrown():
    L = N.lib()
    assert L.gymcuda_create(None, None) == N.EINVAL
    assert b"null" in L.gymcuda_last_error()
    cfg = N.Config()
    L.gymcuda_config_default(C.byref(cfg), 99, 8)
    h = C.c_void_p()
    assert L.gymcuda_create(C.byref(cfg), C.byref(h)) == N.EINVAL
    L.gymcuda_config_default(C.byref(cfg), N.CARTPOLE, 0)
    assert L.gymcuda_create(C.byref(cfg), C.byref(h)) == N.EINVAL
    L.gymcuda_config_default(C.byref(cfg), N.CARTPOLE, 4)
    cfg.gravity = -13.0   # LunarLanderEnv.cs:396-399 throws ArgumentException
    assert L.gymcuda_create(C.byref(cfg), C.byref(h)) == N.EINVAL
    assert b"Gravity" in L.gymcuda_last_error()


def test_no_cpu_fallback():
    from conftest import HAS_GPU
    if HAS_GPU:
        pytest.skip("a GPU is present")
    with pytest.raises(G.GymCudaError) as ei:
        G.CartPoleVecEnv(4)
    assert ei.value.status == N.ECUDA


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "gym.net_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), "%s mentions the oracle" % os.path.join(dirpath, f)
