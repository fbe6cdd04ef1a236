"""ctypes binding of the C ABI in ``include/fpb.h`` (``libfpb.so``).

The library is built in-tree by ``__graft_entry__.build()`` /
``python setup.py build_cmake``.  If it is missing this module raises
``ImportError`` - there is deliberately no CPU fallback (the reference's own
CPU path lives in the reference; ``oracle/`` is test infrastructure).
"""

from __future__ import annotations

import ctypes as C
import os

FPB_ABI_VERSION = 1
FPB_MAX_COMPLEX = 2

WEIGHT_UNIFORM, WEIGHT_BLUEPRINT = 0, 1
MODE_COMPAT, MODE_FRESH = 0, 1
STATE_SILENT, STATE_GKP, STATE_P = 0, 1, 2
STAGE_NOISE, STAGE_SYNDROME, STAGE_GRAPH = 1, 2, 3

LIB_NAME = "libfpb.so"
LIB_PATH = os.environ.get("FPB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

_i32p = C.POINTER(C.c_int32)
_u16p = C.POINTER(C.c_uint16)
_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)
_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)


class ComplexDesc(C.Structure):
    _fields_ = [
        ("n_cubes", C.c_int32),
        ("n_qubits", C.c_int32),
        ("qubit_node", _i32p),
        ("cube_qubit", _i32p),
        ("ninfo", _u16p),
        ("sbase", _i32p),
        ("d_off", C.c_int32 * 12),
        ("s_off", C.c_int32 * 12),
        ("n_slots", C.c_int32),
        ("slot_qubit", _i32p),
        ("n_low", C.c_int32),
        ("low_cube", _i32p),
        ("low_slot", _i32p),
        ("n_high", C.c_int32),
        ("high_cube", _i32p),
        ("high_slot", _i32p),
    ]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("n_nodes", C.c_int32),
        ("adj_indptr", _i32p),
        ("adj_indices", _i32p),
        ("adj_vals", _i8p),
        ("n_complex", C.c_int32),
        ("complex_", ComplexDesc * FPB_MAX_COMPLEX),
        ("delta", C.c_double),
        ("p_swap", C.c_double),
        ("weight_method", C.c_int32),
        ("weight_delta", C.c_double),
        ("weight_integer", C.c_int32),
        ("weight_multiplier", C.c_double),
        ("label_mode", C.c_int32),
        ("chain_len", C.c_int64),
        ("delta_step", C.c_double),
    ]


class Sizes(C.Structure):
    _fields_ = [
        ("n_trials", C.c_int64),
        ("n_complex", C.c_int32),
        ("total_qubits", C.c_int64),
        ("total_odd", C.c_int64 * FPB_MAX_COMPLEX),
        ("total_pairs", C.c_int64 * FPB_MAX_COMPLEX),
    ]


class Outputs(C.Structure):
    _fields_ = [
        ("hom", _f64p),
        ("bits", _u32p),
        ("weights", _f64p),
        ("state", _u8p),
        ("x", _f64p),
        ("parity", _u32p * FPB_MAX_COMPLEX),
        ("odd_count", _i32p * FPB_MAX_COMPLEX),
        ("odd_ids", _i32p * FPB_MAX_COMPLEX),
        ("pair_len", _f64p * FPB_MAX_COMPLEX),
        ("bnd_len", _f64p * FPB_MAX_COMPLEX),
        ("bnd_side", _u8p * FPB_MAX_COMPLEX),
    ]


class Timings(C.Structure):
    _fields_ = [
        ("gen_ms", C.c_float),
        ("noise_ms", C.c_float),
        ("syndrome_ms", C.c_float),
        ("graph_ms", C.c_float),
        ("total_ms", C.c_float),
        ("launches", C.c_int32),
        ("graph_ctas", C.c_int32),
        ("graph_warps_per_cta", C.c_int32),
        ("graph_smem_bytes", C.c_int32),
    ]


EXPORTS = {
    "fpb_last_error": (C.c_char_p, []),
    "fpb_abi_version": (C.c_int, []),
    "fpb_device_count": (C.c_int, []),
    "fpb_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "fpb_destroy": (C.c_int, [C.c_void_p]),
    "fpb_run_injected": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "fpb_run_rng": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int]),
    "fpb_knn": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "fpb_run_bits": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int]),
    "fpb_run_iid": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_double, C.c_int]),
    "fpb_get_sizes": (C.c_int, [C.c_void_p, C.POINTER(Sizes)]),
    "fpb_fetch": (C.c_int, [C.c_void_p, C.POINTER(Outputs), C.c_int]),
    "fpb_device_outputs": (C.c_int, [C.c_void_p, C.POINTER(Outputs)]),
    "fpb_get_timings": (C.c_int, [C.c_void_p, C.POINTER(Timings)]),
    "fpb_synchronize": (C.c_int, [C.c_void_p]),
    "fpb_build_table": (C.c_int, [C.c_void_p]),
    "fpb_mark": (C.c_int, [C.c_void_p, C.c_int]),
    "fpb_elapsed_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "fpb_paths_toggle": (C.c_int, [C.c_void_p, C.c_int64, _i32p, _i32p, _i32p, _i32p, _u32p]),
}

_lib = None


class FpbError(RuntimeError):
    """An fpb_* call failed (message from fpb_last_error), the analogue of the
    ``RuntimeError`` the reference's pybind11 module raises
    (``flamingpy/cpp/lemonpy.cpp:53-54``)."""


def load():
    """Load ``libfpb.so`` (once) and declare every exported prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python setup.py build_cmake --inplace`. There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.fpb_abi_version() != FPB_ABI_VERSION:
        raise ImportError("libfpb.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().fpb_last_error()
        raise FpbError(f"fpb error {rc}: {msg.decode() if msg else '?'}")
