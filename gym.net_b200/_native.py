"""ctypes binding of libgymcuda.so -- the same entry points the C# shim P/Invokes
(csharp/Gym.Environments.Vector/Native.cs).

There is no fallback: if the shared library has not been built this module raises, and if no CUDA
device is present every compute call returns GYMCUDA_ECUDA (raised as GymCudaError).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
NO_OBS = 1   # GYMCUDA_NO_OBS: d_obs value that makes gymcuda_step_device write no observation copy
# GYMCUDA_LIB: development aid -- load another build of the SAME library (kernel A/B experiments under tools/)
LIB_PATH = os.environ.get("GYMCUDA_LIB") or os.path.join(HERE, "csrc", "libgymcuda.so")

OK, EINVAL, EACTION, ECUDA, ENCCL, ENOMEM, ESTATE = 0, -1, -2, -3, -4, -5, -6
STATUS_NAMES = {0: "OK", -1: "EINVAL", -2: "EACTION", -3: "ECUDA", -4: "ENCCL", -5: "ENOMEM", -6: "ESTATE"}

CARTPOLE, PENDULUM, MOUNTAINCAR, MOUNTAINCAR_CONT, ACROBOT, LUNARLANDER, LUNARLANDER_CONT = range(7)
FLAG_AUTO_RESET, FLAG_EPISODE_STATS, FLAG_DONE_BITS = 1, 2, 4


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("env_kind", C.c_int32), ("num_envs", C.c_int32), ("device", C.c_int32),
        ("seed", C.c_uint64), ("env_id_offset", C.c_uint32), ("flags", C.c_uint32), ("time_limit", C.c_int32),
        ("gravity", C.c_float), ("enable_wind", C.c_int32), ("wind_power", C.c_float),
        ("turbulence_power", C.c_float),
    ]


class SpaceInfo(C.Structure):
    _fields_ = [
        ("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("act_n", C.c_int32), ("state_dim", C.c_int32),
        ("aux_dim", C.c_int32), ("time_limit", C.c_int32),
        ("obs_low", C.c_float * 8), ("obs_high", C.c_float * 8),
        ("act_low", C.c_float * 2), ("act_high", C.c_float * 2),
    ]


class Stats(C.Structure):
    _fields_ = [("env_steps", C.c_uint64), ("episodes", C.c_uint64), ("invalid_actions", C.c_uint64),
                ("return_sum", C.c_double), ("length_sum", C.c_uint64)]


# every symbol declared in include/gymcuda.h: name -> (restype, argtypes)
_VP, _I, _U64 = C.c_void_p, C.c_int, C.c_uint64
SYMBOLS = {
    "gymcuda_version": (_I, []),
    "gymcuda_last_error": (C.c_char_p, []),
    "gymcuda_device_count": (_I, [C.POINTER(C.c_int)]),
    "gymcuda_config_default": (_I, [C.POINTER(Config), _I, _I]),
    "gymcuda_create": (_I, [C.POINTER(Config), C.POINTER(_VP)]),
    "gymcuda_destroy": (_I, [_VP]),
    "gymcuda_space": (_I, [_VP, C.POINTER(SpaceInfo)]),
    "gymcuda_num_envs": (_I, [_VP]),
    "gymcuda_seed": (_I, [_VP, _U64]),
    "gymcuda_seed_each": (_I, [_VP, _VP, _I]),
    "gymcuda_reset": (_I, [_VP, _VP]),
    "gymcuda_reset_masked": (_I, [_VP, _VP, _VP]),
    "gymcuda_step": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "gymcuda_step_device": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "gymcuda_step_broadcast": (_I, [_VP, C.c_int32, _VP, _VP, _VP]),
    "gymcuda_step_many_device": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "gymcuda_step_many": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "gymcuda_set_terminal_obs": (_I, [_VP, _VP]),
    "gymcuda_obs_view_device": (_I, [_VP, C.POINTER(C.c_void_p)]),
    "gymcuda_rollout_random_device": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "gymcuda_rollout_random": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "gymcuda_sample_actions": (_I, [_VP, _VP, _VP]),
    "gymcuda_sample_actions_device": (_I, [_VP, _VP, _VP]),
    "gymcuda_box_sample_device": (_I, [_I, _VP, _U64, _U64, _VP, _VP, _I, _I, _I, _VP]),
    "gymcuda_box_sample": (_I, [_I, _U64, _U64, _VP, _VP, _I, _I, _I, _VP]),
    "gymcuda_done_indices": (_I, [_VP, _VP, C.POINTER(C.c_int32)]),
    "gymcuda_done_indices_device": (_I, [_VP, C.POINTER(_VP), C.POINTER(_VP)]),
    "gymcuda_get_state": (_I, [_VP, _VP, _VP, C.POINTER(_U64)]),
    "gymcuda_set_state": (_I, [_VP, _VP, _VP, _U64]),
    "gymcuda_observe": (_I, [_VP, _VP]),
    "gymcuda_render_device": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "gymcuda_render": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "gymcuda_get_stats": (_I, [_VP, C.POINTER(Stats), _I]),
    "gymcuda_normalize_config": (_I, [_VP, C.c_float, C.c_float, C.c_float, C.c_float]),
    "gymcuda_normalize_device": (_I, [_VP, _VP, _VP, _VP, _I]),
    "gymcuda_normalize": (_I, [_VP, _VP, _VP, _VP, _I]),
    "gymcuda_normalize_get": (_I, [_VP, _VP, _VP, _VP, _VP]),
    "gymcuda_normalize_reset": (_I, [_VP]),
    "gymcuda_set_stream": (_I, [_VP, _VP]),
    "gymcuda_set_device_clock": (_I, [_VP, _I]),
    "gymcuda_sync": (_I, [_VP]),
    "gymcuda_host_alloc": (_I, [C.POINTER(_VP), C.c_size_t]),
    "gymcuda_host_free": (_I, [_VP]),
    "gymcuda_host_register": (_I, [_VP, C.c_size_t]),
    "gymcuda_host_unregister": (_I, [_VP]),
    "gymcuda_nccl_load": (_I, [C.c_char_p]),
    "gymcuda_nccl_unique_id": (_I, [_VP]),
    "gymcuda_comm_init": (_I, [_VP, _VP, _I, _I]),
    "gymcuda_allgather_obs": (_I, [_VP, _VP, _VP]),
    "gymcuda_gather_create": (_I, [_VP, _I, _I, _VP]),
    "gymcuda_gather_open": (_I, [_VP, _VP]),
    "gymcuda_step_gather_device": (_I, [_VP, _VP, _VP, _VP, C.POINTER(_VP)]),
    "gymcuda_gather_wait": (_I, [_VP]),
}


class GymCudaError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("gymcuda %s (%d): %s" % (STATUS_NAMES.get(status, "?"), status, message))
        self.status = status


class InvalidActionError(GymCudaError):
    """Gym.Exceptions.InvalidActionError (src/Gym/Exceptions/InvalidActionError.cs:7-10)."""


_lib = None


def lib():
    """Load libgymcuda.so (built by __graft_entry__.build() / make -C gym.net_b200/csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libgymcuda.so is missing at %s: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)   # AttributeError here = ABI mismatch; fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status == OK:
        return
    msg = lib().gymcuda_last_error().decode("utf-8", "replace")
    if status == EACTION:
        raise InvalidActionError(status, msg)
    if status == EINVAL:
        raise ValueError("gymcuda EINVAL: " + msg)   # C#: ArgumentException
    raise GymCudaError(status, msg)
