"""TEST INFRASTRUCTURE ONLY - ctypes wrapper of oracle/fpb_oracle.c (libfpb_oracle.so).

The C restatement of the reference trial, multi-threaded over trials with
OpenMP.  Used by the parity tests for full-size lattices and by bench.py as the
``cpu_baseline`` / ``--impl reference`` arm ("port").  Never imported by the
product package.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from flamingpy_b200 import _lib
from flamingpy_b200.engine import ComplexBatch, TrialBatch, config_dict

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfpb_oracle.so")
SRC_PATH = os.path.join(HERE, "fpb_oracle.c")
GCC_FLAGS = ["-O3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared"]

_i64p = C.POINTER(C.c_int64)


class Result(C.Structure):
    _fields_ = [
        ("n_trials", C.c_int64),
        ("n_complex", C.c_int32),
        ("total_qubits", C.c_int64),
        ("hom", _lib._f64p),
        ("bits", _lib._u8p),
        ("weights", _lib._f64p),
        ("parity", _lib._u8p * 2),
        ("odd_count", _lib._i32p * 2),
        ("odd_off", _i64p * 2),
        ("pair_off", _i64p * 2),
        ("odd_ids", _lib._i32p * 2),
        ("pair_len", _lib._f64p * 2),
        ("bnd_len", _lib._f64p * 2),
        ("bnd_side", _lib._u8p * 2),
    ]


_lib_handle = None


def build(force: bool = False):
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC_PATH):
        subprocess.check_call(["gcc"] + GCC_FLAGS + ["-o", LIB_PATH, SRC_PATH, "-lm"])


def load():
    global _lib_handle
    if _lib_handle is None:
        build()
        lib = C.CDLL(LIB_PATH)
        lib.fpo_run.restype = C.c_int
        lib.fpo_run.argtypes = [C.POINTER(_lib.Config), C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                C.POINTER(Result)]
        lib.fpo_free.restype = None
        lib.fpo_free.argtypes = [C.POINTER(Result)]
        lib.fpo_max_threads.restype = C.c_int
        _lib_handle = lib
    return _lib_handle


def max_threads() -> int:
    return int(load().fpo_max_threads())


def _arr(ptr, n, dtype):
    if n == 0:
        return np.empty(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def run(lat, delta, p_swap, weight_opts, x, state, stage=_lib.STAGE_GRAPH, n_threads=0, copy_out=True):
    """Run injected trials through the C oracle; returns a ``TrialBatch`` with the
    same fields as ``TrialEngine.fetch`` (or None with ``copy_out=False``, for timing)."""
    lib = load()
    wo = {"method": "uniform", "integer": False, "multiplier": 1}
    wo.update(weight_opts or {})
    cfg_d, q_off, Qtot = config_dict(lat, delta, p_swap, wo)
    cfg, keep = _lib.ctypes_config(cfg_d)
    x = np.ascontiguousarray(x, np.float64)
    state = np.ascontiguousarray(state, np.uint8)
    T = x.shape[0]
    res = Result()
    rc = lib.fpo_run(C.byref(cfg), T, x.ctypes.data, state.ctypes.data, stage, n_threads, C.byref(res))
    if rc != 0:
        raise RuntimeError("fpo_run failed")
    try:
        if not copy_out:
            return None
        batch = TrialBatch(T, stage)
        batch.hom = _arr(res.hom, T * Qtot, np.float64).reshape(T, Qtot)
        batch.bits = _arr(res.bits, T * Qtot, np.uint8).reshape(T, Qtot)
        batch.weights = _arr(res.weights, T * Qtot, np.float64).reshape(T, Qtot)
        batch.complexes = {}
        for i, ec in enumerate(lat.ec):
            ct = lat.complexes[ec]
            cb = ComplexBatch()
            if stage >= _lib.STAGE_SYNDROME:
                cb.parity = _arr(res.parity[i], T * ct.S, np.uint8).reshape(T, ct.S)
                cb.odd_count = _arr(res.odd_count[i], T, np.int32)
                cb.odd_off = _arr(res.odd_off[i], T + 1, np.int64)
                cb.pair_off = _arr(res.pair_off[i], T + 1, np.int64)
                cb.odd_ids = _arr(res.odd_ids[i], int(cb.odd_off[-1]), np.int32)
            if stage >= _lib.STAGE_GRAPH:
                cb.pair_len = _arr(res.pair_len[i], int(cb.pair_off[-1]), np.float64)
                if ct.has_boundary:
                    cb.bnd_len = _arr(res.bnd_len[i], int(cb.odd_off[-1]), np.float64)
                    cb.bnd_side = _arr(res.bnd_side[i], int(cb.odd_off[-1]), np.uint8)
                else:
                    cb.bnd_len = np.empty(0, np.float64)
                    cb.bnd_side = np.empty(0, np.uint8)
            batch.complexes[ec] = cb
        return batch
    finally:
        lib.fpo_free(C.byref(res))
        del keep
