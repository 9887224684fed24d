"""ctypes binding of ``libt2d_b200.so`` (C ABI: ``include/t2d_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
CPU fallback: if the shared object is missing, loading fails loudly, and every compute entry
point needs a CUDA device.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("T2D_B200_LIB", os.path.join(_HERE, "libt2d_b200.so"))   # override: A/B builds of the kernels


class T2DError(RuntimeError):
    """A C-ABI call returned a negative code; the message is ``t2d_last_error()``."""


class Config(C.Structure):
    _fields_ = [("interval_ms", C.c_int32), ("delta_t_ms", C.c_int32), ("max_step", C.c_int32),
                ("flags", C.c_int32)]


class TypeParamsC(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "half_len", "half_wid", "radius", "lf", "lr", "steer_lo", "steer_hi", "speed_lo", "speed_hi",
        "accel_lo", "accel_hi", "mass", "mass_height", "mu", "I_z", "cf", "cr")] + [
        ("model", C.c_int32), ("shape", C.c_int32)] + [(n, C.c_float) for n in ("wheel_radius", "T_sb", "T_se", "I_yw")]


class MapTileC(C.Structure):
    """``t2d_map_tile``: the static objects and the boundary of one map (host pointers)."""
    _fields_ = [("segments", C.c_void_p), ("n_seg", C.c_int32), ("poly_start", C.c_void_p), ("n_poly", C.c_int32),
                ("bounds", C.c_void_p)]


class ControllerParamsC(C.Structure):
    """``t2d_controller_params``: one configured controller object."""
    _fields_ = [("kind", C.c_int32)] + [(n, C.c_float) for n in (
        "desired_speed", "time_headway", "min_spacing", "max_acceleration", "comfortable_deceleration", "delta",
        "target_speed", "kp", "accel_change_rate", "delta_t", "max_accel", "min_accel", "interval",
        "min_pre_aiming_distance", "pp_interval", "wheel_base")]


# name -> (restype, argtypes); every symbol include/t2d_b200.h declares
_P = C.c_void_p
SYMBOLS = {
    "t2d_version": (C.c_int, []),
    "t2d_last_error": (C.c_char_p, []),
    "t2d_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.POINTER(Config)]),
    "t2d_destroy": (C.c_int, [_P]),
    "t2d_set_config": (C.c_int, [_P, C.POINTER(Config)]),
    "t2d_set_type_table": (C.c_int, [_P, C.POINTER(TypeParamsC), C.c_int]),
    "t2d_set_map": (C.c_int, [_P, _P, C.c_int, _P, C.c_float]),
    "t2d_set_map_polygons": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P, C.c_float]),
    "t2d_set_map_table": (C.c_int, [_P, _P, C.c_int, _P, C.c_float]),
    "t2d_bind_state": (C.c_int, [_P] + [_P] * 8),
    "t2d_step": (C.c_int, [_P] + [_P] * 7),
    "t2d_step_host": (C.c_int, [_P] + [_P] * 7),
    "t2d_check_events": (C.c_int, [_P] + [_P] * 4),
    "t2d_set_ego_action": (C.c_int, [_P, _P]),
    "t2d_step_host_ego": (C.c_int, [_P] + [_P] * 8),
    "t2d_env_epilogue": (C.c_int, [_P] + [_P] * 9 + [C.c_int, _P]),
    "t2d_set_goal": (C.c_int, [_P, _P, C.c_float, C.c_int, _P, _P, _P]),
    "t2d_reset": (C.c_int, [_P, _P, _P, C.c_int] + [_P] * 7),
    "t2d_lidar_scan": (C.c_int, [_P, C.c_int, C.c_float, _P, _P, _P]),
    "t2d_set_controllers": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P]),
    "t2d_set_paths": (C.c_int, [_P, _P, _P, C.c_int]),
    "t2d_control": (C.c_int, [_P, _P, _P]),
    "t2d_exchange_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "t2d_exchange_connect": (C.c_int, [_P, _P]),
    "t2d_exchange_allgather": (C.c_int, [_P, _P, _P, _P]),
    "t2d_exchange_allgather_lagged": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "t2d_exchange_status": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "t2d_exchange_destroy": (C.c_int, [_P]),
    "t2d_physics_step": (C.c_int, [C.c_int, C.POINTER(TypeParamsC), C.c_int, C.c_int, C.c_int] + [_P] * 11),
    "t2d_bind_wheel_state": (C.c_int, [_P, _P, _P]),
    "t2d_bind_reset_wheel_pool": (C.c_int, [_P, _P, _P]),
    "t2d_debug_set_clock_buffer": (C.c_int, [_P, _P]),
    "t2d_set_prefetch": (C.c_int, [_P, C.c_int]),
    "t2d_launch_count": (C.c_int64, []),
}

_lib = None


def load():
    """Load the shared library (once) and declare its prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  tactics2d_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        msg = load().t2d_last_error()
        raise T2DError(f"t2d error {code}: {msg.decode() if msg else ''}")
