"""ctypes binding of libquadsim.so (include/quadsim.h).  There is no CPU fallback: if the CUDA library is
missing this module raises, loudly."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p

E2E, INDI = 0, 1
MODE_NORMAL, MODE_PAUSE_IF_COLLISION, MODE_PAUSE = 0, 1, 2
RESET_DEVICE, RESET_HOST = 0, 1
F_DONE, F_TRUNCATED, F_GATE_PASSED, F_GATE_COLLISION, F_GROUND, F_OUT_OF_BOUNDS = 1, 2, 4, 8, 16, 32


class QsStats(C.Structure):
    _fields_ = [("reward_sum", C.c_double), ("env_steps", C.c_uint64), ("dones", C.c_uint64),
                ("truncated", C.c_uint64), ("gates_passed", C.c_uint64), ("gate_collisions", C.c_uint64),
                ("ground_collisions", C.c_uint64), ("out_of_bounds", C.c_uint64)]


class QsStepInfo(C.Structure):
    _fields_ = [("last_done_index", C.c_int64), ("n_done", C.c_int64), ("any_truncated", C.c_int32),
                ("reserved", C.c_int32)]


F32, F64 = 0, 1


class QsTrainHyper(C.Structure):
    _fields_ = [("learning_rate", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("clip_range", C.c_float), ("vf_coef", C.c_float), ("ent_coef", C.c_float), ("max_grad_norm", C.c_float),
                ("obs_limit", C.c_float), ("act_limit", C.c_float), ("normalize_advantage", C.c_int32),
                ("reserved", C.c_int32)]

# name -> (restype, argtypes); every symbol include/quadsim.h declares
SIGNATURES = {
    "qs_state_len": (C.c_int, [C.c_int]),
    "qs_obs_len": (C.c_int, [C.c_int, C.c_int]),
    "qs_version": (C.c_char_p, []),
    "qs_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int64, C.c_int, _fp, _fp, _fp, C.c_int, C.c_int, _vp]),
    "qs_destroy": (C.c_int, [_vp]),
    "qs_last_error": (C.c_char_p, [_vp]),
    "qs_set_stream": (C.c_int, [_vp, _vp]),
    "qs_set_track_tables": (C.c_int, [_vp, _fp, _fp, _fp, _fp]),
    "qs_get_track_tables": (C.c_int, [_vp, _fp, _fp, _fp, _fp]),
    "qs_set_max_steps": (C.c_int, [_vp, C.c_int64]),
    "qs_set_dt": (C.c_int, [_vp, C.c_float]),
    "qs_set_disturbance_ranges": (C.c_int, [_vp, _dp, C.c_int, C.c_double]),
    "qs_set_residual_weights": (C.c_int, [_vp, _fp, _fp]),
    "qs_seed": (C.c_int, [_vp, C.c_uint64]),
    "qs_set_env_offset": (C.c_int, [_vp, C.c_int64]),
    "qs_set_obs_peers": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.c_int64]),
    "qs_set_obs_format": (C.c_int, [_vp, C.c_int]),
    "qs_obs_packed_bytes": (C.c_int64, [C.c_int, C.c_int64]),
    "qs_enable_stats": (C.c_int, [_vp, C.c_int]),
    "qs_get_stats": (C.c_int, [_vp, C.POINTER(QsStats), C.c_int]),
    "qs_set_state": (C.c_int, [_vp, C.c_int64, C.c_int64, _fp, _fp, _i64p, _i64p]),
    "qs_get_state": (C.c_int, [_vp, C.c_int64, C.c_int64, _fp, _fp, _i64p, _i64p]),
    "qs_observe": (C.c_int, [_vp, _vp]),
    "qs_reset_all": (C.c_int, [_vp, _vp]),
    "qs_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "qs_apply_reset": (C.c_int, [_vp, C.c_int64, _i32p, _fp, _fp, _vp]),
    "qs_step_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int]),
    "qs_step_host_ex": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(QsStepInfo)]),
    "qs_reset_all_host": (C.c_int, [_vp, _vp]),
    "qs_observe_host": (C.c_int, [_vp, _vp]),
    "qs_host_alloc": (_vp, [C.c_size_t]),
    "qs_host_free": (None, [_vp]),
    "qs_algorithmic_bytes_per_env_step": (C.c_int, [C.c_int, C.c_int]),
    "qs_launch_count": (C.c_uint64, [_vp]),
    "qs_chained_launch_count": (C.c_uint64, [_vp]),
    "qs_get_state_layout": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "qs_policy_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "qs_policy_destroy": (C.c_int, [_vp]),
    "qs_policy_last_error": (C.c_char_p, [_vp]),
    "qs_policy_set_stream": (C.c_int, [_vp, _vp]),
    "qs_policy_set_layer": (C.c_int, [_vp, C.c_int, _fp, _fp]),
    "qs_policy_set_std": (C.c_int, [_vp, _fp]),
    "qs_policy_set_activation": (C.c_int, [_vp, C.c_int]),
    "qs_policy_seed": (C.c_int, [_vp, C.c_uint64]),
    "qs_policy_set_env_offset": (C.c_int, [_vp, C.c_int64]),
    "qs_policy_set_obs_limit": (C.c_int, [_vp, C.c_float]),
    "qs_policy_forward": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int]),
    "qs_policy_forward_packed": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int]),
    "qs_policy_launch_count": (C.c_uint64, [_vp]),
    "qs_rollout": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "qs_rollout_fused_supported": (C.c_int, [_vp, _vp]),
    "qs_rollout_fused": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "qs_trainer_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _vp]),
    "qs_trainer_destroy": (C.c_int, [_vp]),
    "qs_trainer_last_error": (C.c_char_p, [_vp]),
    "qs_trainer_set_stream": (C.c_int, [_vp, _vp]),
    "qs_trainer_set_layer": (C.c_int, [_vp, C.c_int, C.c_int, _fp, _fp]),
    "qs_trainer_get_layer": (C.c_int, [_vp, C.c_int, C.c_int, _fp, _fp]),
    "qs_trainer_set_log_std": (C.c_int, [_vp, _fp]),
    "qs_trainer_get_log_std": (C.c_int, [_vp, _fp]),
    "qs_trainer_reset_optimizer": (C.c_int, [_vp]),
    "qs_trainer_minibatch": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(QsTrainHyper), C.c_int]),
    "qs_trainer_get_grad": (C.c_int, [_vp, C.c_int, C.c_int, _fp, _fp, _fp]),
    "qs_trainer_get_stats": (C.c_int, [_vp, _fp, C.c_int]),
    "qs_trainer_publish": (C.c_int, [_vp, _vp]),
    "qs_trainer_launch_count": (C.c_uint64, [_vp]),
    "qs_gae": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int, C.c_float, C.c_float, _vp]),
}

_LIB = None


class QuadsimError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Load libquadsim.so; (re)build it with nvcc when missing or stale and a compiler is present."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("QS_LIB") or _build.LIB  # QS_LIB: an experimental build of the same sources (tools/)
    if path == _build.LIB and build_if_missing and _build.is_stale():
        try:
            _build.build_library()
        except Exception as exc:  # stale-but-present library is still usable; missing one is fatal
            if not os.path.isfile(path):
                raise QuadsimError(f"libquadsim.so is missing and could not be built: {exc}") from exc
    if not os.path.isfile(path):
        raise QuadsimError(f"{path} not found: run `python -m optimal_quad_control_rl_b200.build` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype, fn.argtypes = res, args
    _LIB = lib
    return lib


def check(lib, handle, status, what):
    if status != 0:
        msg = lib.qs_last_error(handle)
        raise QuadsimError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")
