"""ctypes view of include/dmfb_b200.h.  No torch types cross this boundary: only integers,
raw device pointers and the stream handle."""
import ctypes as C
import os

from . import build as _build

DMFB_MAX_DIM = 128
DMFB_MAX_AGENTS = 32
DMFB_MAX_FOV = 19
DMFB_L2_WORDS = 12

STEP_RECORD_USAGE = 1
STEP_FREEZE_TERM = 2
STEP_AUTO_RESET = 4
STEP_SKIP_TASK_SEARCH = 8


def step_search_share(p):
    return (int(p) & 0xFF) << 8

STATUS_ILLEGAL_ACTION = 1
STATUS_SAMPLER_GAVE_UP = 2

DMFB_OBS_BASE = 0
DMFB_OBS_V01 = 1
MEDA_OBS_BASE = 0
MEDA_OBS_V01 = 1
MEDA_OBS_V02 = 2


class DmfbCfg(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("length", C.c_int32), ("n_agents", C.c_int32), ("n_blocks", C.c_int32),
        ("fov", C.c_int32), ("stall", C.c_int32), ("b_degrade", C.c_int32), ("max_step", C.c_int32),
        ("n_actions", C.c_int32), ("obs_dim", C.c_int32), ("l2_words", C.c_int32), ("obs_version", C.c_int32),
        ("per_degrade", C.c_double), ("env_base", C.c_int64),
        ("dir_x", C.c_int8 * (2 * DMFB_MAX_DIM)), ("dir_y", C.c_int8 * (2 * DMFB_MAX_DIM)),
        ("l2_row", (C.c_uint32 * DMFB_L2_WORDS) * DMFB_MAX_FOV),
        ("l2_col", (C.c_uint32 * DMFB_L2_WORDS) * DMFB_MAX_FOV),
    ]


class DmfbState(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("usage_log_cap", C.c_int32),
        ("drop", C.c_void_p), ("start", C.c_void_p), ("step_count", C.c_void_p), ("constraints", C.c_void_p),
        ("terminated", C.c_void_p), ("episode", C.c_void_p), ("usage", C.c_void_p), ("health", C.c_void_p),
        ("degrade", C.c_void_p), ("blocks", C.c_void_p), ("usage_log", C.c_void_p), ("usage_log_len", C.c_void_p),
        ("next_task", C.c_void_p), ("next_cursor", C.c_void_p), ("gen_status", C.c_void_p),
        ("health_bits", C.c_void_p),
    ]


class DmfbOut(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p), ("reward", C.c_void_p), ("reward_f64", C.c_void_p), ("team_reward", C.c_void_p),
        ("done", C.c_void_p), ("avail", C.c_void_p), ("constraints", C.c_void_p), ("success", C.c_void_p),
        ("terminated", C.c_void_p), ("padded", C.c_void_p), ("status", C.c_void_p),
    ]


class MedaCfg(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("length", C.c_int32), ("n_agents", C.c_int32), ("fov", C.c_int32),
        ("b_degrade", C.c_int32), ("obs_version", C.c_int32), ("max_step", C.c_int32), ("n_actions", C.c_int32),
        ("obs_dim", C.c_int32), ("radius", C.c_int32),
        ("per_degrade", C.c_double), ("env_base", C.c_int64),
        ("dir_x", C.c_int8 * (2 * DMFB_MAX_DIM)), ("dir_y", C.c_int8 * (2 * DMFB_MAX_DIM)),
    ]


class MedaState(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("usage_log_cap", C.c_int32),
        ("drop", C.c_void_p), ("start", C.c_void_p), ("status", C.c_void_p), ("step_count", C.c_void_p),
        ("fails", C.c_void_p), ("terminated", C.c_void_p), ("episode", C.c_void_p),
        ("usage", C.c_void_p), ("health", C.c_void_p), ("degrade", C.c_void_p),
        ("usage_log", C.c_void_p), ("usage_log_len", C.c_void_p), ("reset_list", C.c_void_p), ("reset_count", C.c_void_p),
        ("gen_status", C.c_void_p), ("health_bits", C.c_void_p),
    ]


MedaOut = DmfbOut  # same field list (include/dmfb_b200.h: meda_out_t)

# every symbol include/dmfb_b200.h declares
EXPORTS = [
    "dmfb_cfg_init", "dmfb_cfg_set_obs_version", "dmfb_flush_usage", "dmfb_sync_health_bits", "dmfb_step", "dmfb_reset", "dmfb_observe", "dmfb_global_state", "dmfb_restart",
    "meda_cfg_init", "meda_step", "meda_reset", "meda_observe", "meda_restart", "meda_set_order", "meda_flush_usage", "meda_sync_health_bits",
    "dmfb_abi_version", "dmfb_last_cuda_error", "dmfb_launch_count",
    "dmfb_host_create", "dmfb_host_destroy", "dmfb_host_reset", "dmfb_host_step", "dmfb_host_set_transfer", "dmfb_host_unpack_records",
    "dmfb_host_alloc_pinned", "dmfb_host_free_pinned",
    "meda_host_create", "meda_host_destroy", "meda_host_reset", "meda_host_step",
]

_lib = None


def lib_path():
    # DMFB_B200_LIB: an alternative build of the library (kernel experiments: tools/build_variants.py)
    return os.environ.get("DMFB_B200_LIB") or _build.LIB


def load():
    """Load libdmfb_b200.so.  There is NO fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc).  marl-dmfb_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    lib.dmfb_last_cuda_error.restype = C.c_char_p
    lib.dmfb_launch_count.restype = C.c_uint64
    lib.dmfb_cfg_init.argtypes = [C.POINTER(DmfbCfg)] + [C.c_int] * 7 + [C.c_double]
    lib.dmfb_cfg_set_obs_version.argtypes = [C.POINTER(DmfbCfg), C.c_int]
    lib.dmfb_step.argtypes = [C.POINTER(DmfbCfg), C.POINTER(DmfbState), C.c_void_p, C.c_int, C.c_void_p,
                              C.c_uint64, C.c_uint32, C.POINTER(DmfbOut), C.c_void_p]
    lib.dmfb_reset.argtypes = [C.POINTER(DmfbCfg), C.POINTER(DmfbState), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.dmfb_restart.argtypes = [C.POINTER(DmfbCfg), C.POINTER(DmfbState), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dmfb_observe.argtypes = [C.POINTER(DmfbCfg), C.POINTER(DmfbState), C.c_void_p, C.c_void_p]
    lib.dmfb_flush_usage.argtypes = [C.POINTER(DmfbCfg), C.POINTER(DmfbState), C.c_void_p]
    lib.dmfb_sync_health_bits.argtypes = [C.POINTER(DmfbCfg), C.POINTER(DmfbState), C.c_void_p]
    lib.dmfb_global_state.argtypes = [C.POINTER(DmfbCfg), C.POINTER(DmfbState), C.c_void_p, C.c_void_p]
    if hasattr(lib, "meda_cfg_init"):
        lib.meda_cfg_init.argtypes = [C.POINTER(MedaCfg)] + [C.c_int] * 5 + [C.c_double, C.c_int]
        lib.meda_step.argtypes = [C.POINTER(MedaCfg), C.POINTER(MedaState), C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_uint64, C.c_uint32, C.c_void_p, C.POINTER(MedaOut), C.c_void_p]
        lib.meda_reset.argtypes = [C.POINTER(MedaCfg), C.POINTER(MedaState), C.c_void_p, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.meda_observe.argtypes = [C.POINTER(MedaCfg), C.POINTER(MedaState), C.c_void_p, C.c_void_p, C.c_void_p]
        lib.meda_set_order.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        lib.meda_flush_usage.argtypes = [C.POINTER(MedaCfg), C.POINTER(MedaState), C.c_void_p]
        lib.meda_sync_health_bits.argtypes = [C.POINTER(MedaCfg), C.POINTER(MedaState), C.c_void_p]
        lib.meda_restart.argtypes = [C.POINTER(MedaCfg), C.POINTER(MedaState), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
    if hasattr(lib, "dmfb_host_create"):
        lib.dmfb_host_create.argtypes = [C.POINTER(DmfbCfg), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.dmfb_host_set_transfer.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.dmfb_host_unpack_records.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_size_t, C.c_int]
        lib.dmfb_host_destroy.argtypes = [C.c_void_p]
        lib.dmfb_host_destroy.restype = None
        lib.dmfb_host_reset.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.dmfb_host_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32] + [C.c_void_p] * 5
        lib.dmfb_host_alloc_pinned.restype = C.c_void_p
        lib.dmfb_host_alloc_pinned.argtypes = [C.c_size_t]
        lib.dmfb_host_free_pinned.argtypes = [C.c_void_p]
        lib.dmfb_host_free_pinned.restype = None
    if hasattr(lib, "meda_host_create"):
        lib.meda_host_create.argtypes = [C.POINTER(MedaCfg), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.meda_host_destroy.argtypes = [C.c_void_p]
        lib.meda_host_destroy.restype = None
        lib.meda_host_reset.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        lib.meda_host_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32] + [C.c_void_p] * 5
    _lib = lib
    return lib


def check(rc, what=""):
    """Map a DMFB_ERR_* code onto the exception the reference raises at the same place."""
    if rc == 0:
        return
    if rc == 1:
        raise RuntimeError("Fov is too large")                     # dmfb.py:139-140
    if rc == 2:
        raise TypeError("Too many droplets for DMFB")             # dmfb.py:144-146
    if rc == 4:
        raise ZeroDivisionError("division by zero")               # dmfb.py:446 (fov//2 == 10)
    if rc == 6:
        raise AssertionError("width >= 5 and length >= 5")        # dmfb.py:489
    if rc == 5:
        raise RuntimeError(f"CUDA error in {what}: {load().dmfb_last_cuda_error().decode()}")
    msg = load().dmfb_last_cuda_error().decode()
    raise ValueError(f"{what}: bad argument ({msg})")
