"""ctypes wrapper of oracle/c/liboracle_tick.so (TEST INFRASTRUCTURE ONLY): the float64 oracle in
plain C + OpenMP, for full-size checks and the compiled CPU figure in bench.py."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "liboracle_tick.so")
_SRC = os.path.join(_DIR, "oracle_tick.c")
_lib = None

FIELDS = ("half_len", "half_wid", "radius", "lf", "lr", "steer_lo", "steer_hi", "speed_lo", "speed_hi", "accel_lo",
          "accel_hi", "mass", "mass_height", "mu", "I_z", "cf", "cr")


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", _SO,
                               _SRC, "-lm"])
    return _SO


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def table_bytes(table: dict) -> np.ndarray:
    """Oracle table dict (column arrays) -> packed array of t2d_type_params rows."""
    n = len(table["model"])
    dt = np.dtype([(k, np.float32) for k in FIELDS] + [("model", np.int32), ("shape", np.int32)])
    arr = np.zeros(n, dtype=dt)
    for k in FIELDS:
        arr[k] = np.asarray(table[k], dtype=np.float32)
    arr["model"] = table["model"]
    arr["shape"] = table["shape"]
    return arr


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def physics(state, type_id, action, table, interval=100, delta_t=5, steer_first=False, threads=None):
    lib = load()
    tb = table_bytes(table)
    tid = np.ascontiguousarray(type_id, dtype=np.uint8)
    f = {k: np.ascontiguousarray(state[k], dtype=np.float32) for k in ("x", "y", "heading", "speed", "vx", "vy")}
    act = np.ascontiguousarray(action, dtype=np.float32)
    out = {k: np.empty(tid.shape, np.float64) for k in f}
    lib.oracle_physics(C.c_int(tid.size), _p(tb), C.c_int(len(tb)), _p(f["x"]), _p(f["y"]), _p(f["heading"]), _p(f["speed"]),
                       _p(f["vx"]), _p(f["vy"]), _p(tid), _p(act), C.c_int(interval), C.c_int(delta_t), C.c_int(int(steer_first)),
                       _p(out["x"]), _p(out["y"]), _p(out["heading"]), _p(out["speed"]), _p(out["vx"]), _p(out["vy"]))
    return out


def events(x, y, heading, type_id, table, segments=None, bounds=None):
    lib = load()
    tb = table_bytes(table)
    tid = np.ascontiguousarray(type_id, dtype=np.uint8)
    N, M = tid.shape
    assert M <= 256
    xs, ys, hs = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, heading))
    seg = None if segments is None or len(segments) == 0 else np.ascontiguousarray(segments, dtype=np.float32)
    b = None if bounds is None else np.asarray(bounds, dtype=np.float32)
    flags = np.zeros((N, M), np.uint8)
    hi = np.zeros((N, M), np.int16)
    hseg = np.zeros((N, M), np.int16)
    lib.oracle_events(C.c_int(N), C.c_int(M), _p(tb), C.c_int(len(tb)), _p(xs), _p(ys), _p(hs), _p(tid), _p(seg),
                      C.c_int(0 if seg is None else len(seg)), _p(b), _p(flags), _p(hi), _p(hseg))
    return flags, hi, hseg
