"""Bindings of the C ABI in ``include/fpb.h`` (``libfpb.so``).

Two interchangeable bindings with one interface (``open_handle``):

* the pybind11 module ``flamingpy_b200.cpp.fpbpy`` - the reference's native pattern
  (``flamingpy/cpp/lemonpy.cpp:50-84``, imported behind ``try/except ImportError`` like
  ``flamingpy/__init__.py:28-34``) - used whenever it is importable;
* plain ``ctypes`` on ``libfpb.so`` otherwise (or with ``FPB_BINDING=ctypes``).

Both are built in-tree by ``__graft_entry__.build()`` / ``python setup.py build_cmake``.  If
``libfpb.so`` is missing this module raises ``ImportError`` - there is deliberately no CPU
fallback (the reference's own CPU path lives in the reference; ``oracle/`` is test infrastructure).
``FPB_LIB`` names another build of the same ABI explicitly; the test suite uses it to run the GPU
tests against the CPU emulation build of the kernel sources (``tests/emu``), nothing else sets it.
"""

from __future__ import annotations

import ctypes as C
import os

FPB_ABI_VERSION = 2
FPB_MAX_COMPLEX = 2

WEIGHT_UNIFORM, WEIGHT_BLUEPRINT = 0, 1
MODE_COMPAT, MODE_FRESH = 0, 1
STATE_SILENT, STATE_GKP, STATE_P = 0, 1, 2
STAGE_NOISE, STAGE_SYNDROME, STAGE_GRAPH = 1, 2, 3
LENGTH_FLOAT, LENGTH_RX_TRUNC = 0, 1

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
        ("length_mode", C.c_int32),
        ("reserve_sms", C.c_int32),
        ("reserved_i", C.c_int32 * 6),
        ("reserved_d", C.c_double * 4),
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
    "fpb_run_weighted": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "fpb_paths_list": (C.c_int, [C.c_void_p, C.c_int64, _i32p, _i32p, _i32p, _i32p, C.c_int32, _i32p, _i32p]),
    "fpb_knn_len": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fpb_pairs_at": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, _i32p, _i32p, _i32p, _f64p]),
    "fpb_pairs_of_trial": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, _f64p, C.c_int64]),
    "fpb_price": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int64), _i32p, C.POINTER(C.c_int64), _f64p, _i32p, _u8p,
                            C.c_int64, _i32p, _i32p, _i32p, _f64p, C.POINTER(C.c_int64)]),
    "fpb_set_macro": (C.c_int, [C.c_void_p, _i32p, _i8p]),
    "fpb_run_macro_injected": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "fpb_run_macro_rng": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int]),
    "fpb_fetch_macro": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
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
    "fpb_uf_set_graph": (C.c_int, [C.c_void_p, C.c_int, C.c_int32, _i32p, _i32p]),
    "fpb_uf_decode": (C.c_int, [C.c_void_p, C.c_int, _u32p, _u32p, _i32p]),
    "fpb_set_planes": (C.c_int, [C.c_void_p, C.c_int32, _i32p, _i32p]),
    "fpb_logical_check": (C.c_int, [C.c_void_p, _u8p, C.POINTER(C.c_int64)]),
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


# ---------------------------------------------------------------------------------------------
# One handle interface over both bindings
# ---------------------------------------------------------------------------------------------
_PER_COMPLEX = ("parity", "odd_count", "odd_ids", "pair_len", "bnd_len", "bnd_side")
_OUT_TYPES = {"hom": _f64p, "bits": _u32p, "weights": _f64p, "state": _u8p, "x": _f64p, "parity": _u32p,
              "odd_count": _i32p, "odd_ids": _i32p, "pair_len": _f64p, "bnd_len": _f64p, "bnd_side": _u8p}


class SizesView:
    """Result sizes of the last run (``fpb_sizes``) with attribute access."""

    __slots__ = ("n_trials", "n_complex", "total_qubits", "total_odd", "total_pairs")

    def __init__(self, d):
        self.n_trials = int(d["n_trials"])
        self.n_complex = int(d["n_complex"])
        self.total_qubits = int(d["total_qubits"])
        self.total_odd = [int(v) for v in d["total_odd"]]
        self.total_pairs = [int(v) for v in d["total_pairs"]]


def ctypes_config(cfg: dict):
    """``fpb_config`` as a ctypes struct from a config dict.  Returns (struct, keep-alive list)."""
    import numpy as np

    keep = []

    def arr(a, dt, typ):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data_as(typ)

    c = Config()
    c.abi_version = FPB_ABI_VERSION
    c.device = int(cfg.get("device", 0))
    c.n_nodes = int(cfg["n_nodes"])
    c.adj_indptr = arr(cfg["adj_indptr"], np.int32, _i32p)
    c.adj_indices = arr(cfg["adj_indices"], np.int32, _i32p)
    c.adj_vals = arr(cfg["adj_vals"], np.int8, _i8p)
    cxs = cfg["complexes"]
    if not 1 <= len(cxs) <= FPB_MAX_COMPLEX:
        raise FpbError("config: 1 or 2 complexes")
    c.n_complex = len(cxs)
    for i, d in enumerate(cxs):
        x = c.complex_[i]
        x.n_cubes, x.n_qubits, x.n_slots = int(d["n_cubes"]), int(d["n_qubits"]), int(d["n_slots"])
        x.qubit_node = arr(d["qubit_node"], np.int32, _i32p)
        x.cube_qubit = arr(d["cube_qubit"], np.int32, _i32p)
        x.ninfo = arr(d["ninfo"], np.uint16, _u16p)
        x.sbase = arr(d["sbase"], np.int32, _i32p)
        for j, v in enumerate(np.asarray(d["d_off"], np.int32).ravel()):
            x.d_off[j] = int(v)
        for j, v in enumerate(np.asarray(d["s_off"], np.int32).ravel()):
            x.s_off[j] = int(v)
        x.slot_qubit = arr(d["slot_qubit"], np.int32, _i32p)
        x.n_low, x.n_high = len(d["low_cube"]), len(d["high_cube"])
        x.low_cube = arr(d["low_cube"], np.int32, _i32p)
        x.low_slot = arr(d["low_slot"], np.int32, _i32p)
        x.high_cube = arr(d["high_cube"], np.int32, _i32p)
        x.high_slot = arr(d["high_slot"], np.int32, _i32p)
    c.delta = float(cfg["delta"])
    c.p_swap = float(cfg.get("p_swap", 0.0))
    c.weight_method = int(cfg.get("weight_method", WEIGHT_UNIFORM))
    c.weight_delta = float(cfg.get("weight_delta", c.delta))
    c.weight_integer = int(cfg.get("weight_integer", 0))
    c.weight_multiplier = float(cfg.get("weight_multiplier", 1.0))
    c.label_mode = int(cfg.get("label_mode", MODE_COMPAT))
    c.chain_len = int(cfg.get("chain_len", 0))
    c.delta_step = float(cfg.get("delta_step", 0.0))
    c.length_mode = int(cfg.get("length_mode", LENGTH_FLOAT))
    c.reserve_sms = int(cfg.get("reserve_sms", 0))
    return c, keep


class CtypesHandle:
    """``fpb_handle`` driven through ctypes; same methods as ``fpbpy.Handle``."""

    def __init__(self, cfg: dict):
        lib = load()
        self._lib = lib
        c, keep = ctypes_config(cfg)
        self._cubes = [c.complex_[i].n_cubes for i in range(c.n_complex)]
        self._has_bnd = [c.complex_[i].n_low > 0 for i in range(c.n_complex)]
        self._qtot = sum(c.complex_[i].n_qubits for i in range(c.n_complex))
        self._n_nodes = c.n_nodes
        self._bit_words = (self._qtot + 31) // 32
        h = C.c_void_p()
        check(lib.fpb_create(C.byref(c), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.fpb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def address(self):
        return self._h.value

    def run_injected(self, x, state, stage=STAGE_GRAPH):
        check(self._lib.fpb_run_injected(self._h, x.shape[0], x.ctypes.data, state.ctypes.data, 0, stage))

    def run_rng(self, seed, first_trial, n_trials, stage=STAGE_GRAPH):
        check(self._lib.fpb_run_rng(self._h, seed, first_trial, n_trials, stage))

    def run_bits(self, bits, stage=STAGE_GRAPH):
        check(self._lib.fpb_run_bits(self._h, bits.shape[0], bits.ctypes.data, 0, stage))

    def run_weighted(self, weights, bits, stage=STAGE_GRAPH):
        check(self._lib.fpb_run_weighted(self._h, weights.shape[0], weights.ctypes.data, bits.ctypes.data, 0, stage))

    def set_macro(self, partner, sign):
        check(self._lib.fpb_set_macro(self._h, partner.ctypes.data_as(_i32p), sign.ctypes.data_as(_i8p)))

    def run_macro_injected(self, state, mean, z, stage=STAGE_GRAPH):
        check(self._lib.fpb_run_macro_injected(self._h, state.shape[0], state.ctypes.data, mean.ctypes.data,
                                               z.ctypes.data, 0, stage))

    def run_macro_rng(self, seed, first_trial, n_trials, stage=STAGE_GRAPH):
        check(self._lib.fpb_run_macro_rng(self._h, seed, first_trial, n_trials, stage))

    def fetch_macro(self):
        import numpy as np

        T = self.sizes()["n_trials"]
        rs, bv = np.empty((T, self._n_nodes), np.uint8), np.empty((T, self._n_nodes), np.uint8)
        pp, oc = np.empty((T, self._n_nodes), np.float64), np.empty((T, self._n_nodes), np.float64)
        check(self._lib.fpb_fetch_macro(self._h, rs.ctypes.data, bv.ctypes.data, pp.ctypes.data, oc.ctypes.data, 0))
        return rs, bv, pp, oc

    def run_iid(self, seed, first_trial, n_trials, p_err, stage=STAGE_GRAPH):
        check(self._lib.fpb_run_iid(self._h, seed, first_trial, n_trials, float(p_err), stage))

    def sizes(self):
        s = Sizes()
        check(self._lib.fpb_get_sizes(self._h, C.byref(s)))
        return {"n_trials": s.n_trials, "n_complex": s.n_complex, "total_qubits": s.total_qubits,
                "total_odd": list(s.total_odd), "total_pairs": list(s.total_pairs)}

    def fetch_into(self, bufs: dict):
        s = self.sizes()
        T = s["n_trials"]
        out = Outputs()

        def dest(key, field, count, index=None):
            a = bufs.get(key)
            if a is None:
                return
            if a.size < count:
                raise FpbError(f"buffer {key!r} too small: {a.size} < {count}")
            if not (a.flags.c_contiguous and a.flags.writeable):
                raise FpbError(f"buffer {key!r} must be C-contiguous and writable")
            p = a.ctypes.data_as(_OUT_TYPES[field])
            if index is None:
                setattr(out, field, p)
            else:
                getattr(out, field)[index] = p

        dest("hom", "hom", T * self._qtot)
        dest("weights", "weights", T * self._qtot)
        dest("bits", "bits", T * self._bit_words)
        dest("state", "state", T * self._n_nodes)
        dest("x", "x", T * 2 * self._n_nodes)
        for i, S in enumerate(self._cubes):
            dest(f"parity{i}", "parity", T * ((S + 31) // 32), i)
            dest(f"odd_count{i}", "odd_count", T, i)
            dest(f"odd_ids{i}", "odd_ids", s["total_odd"][i], i)
            dest(f"pair_len{i}", "pair_len", s["total_pairs"][i], i)
            if self._has_bnd[i]:
                dest(f"bnd_len{i}", "bnd_len", s["total_odd"][i], i)
                dest(f"bnd_side{i}", "bnd_side", s["total_odd"][i], i)
        check(self._lib.fpb_fetch(self._h, C.byref(out), 0))

    def pair_lengths(self, complex_index=0):
        import numpy as np

        out = np.empty(self.sizes()["total_pairs"][complex_index], np.float64)
        self.fetch_into({f"pair_len{complex_index}": out})
        return out

    def device_outputs(self):
        o = Outputs()
        check(self._lib.fpb_device_outputs(self._h, C.byref(o)))
        addr = lambda p: C.cast(p, C.c_void_p).value or 0  # noqa: E731
        d = {k: addr(getattr(o, k)) for k in ("hom", "bits", "weights", "state", "x")}
        for i in range(len(self._cubes)):
            for f in _PER_COMPLEX:
                d[f"{f}{i}"] = addr(getattr(o, f)[i])
        return d

    def timings(self):
        t = Timings()
        check(self._lib.fpb_get_timings(self._h, C.byref(t)))
        return {name: getattr(t, name) for name, _ in Timings._fields_}

    def synchronize(self):
        check(self._lib.fpb_synchronize(self._h))

    def build_table(self):
        check(self._lib.fpb_build_table(self._h))

    def mark(self, which):
        check(self._lib.fpb_mark(self._h, which))

    def elapsed_ms(self):
        ms = C.c_float()
        check(self._lib.fpb_elapsed_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def knn(self, complex_index, k, out):
        check(self._lib.fpb_knn(self._h, complex_index, k, out.ctypes.data, 0))

    def paths_toggle(self, trial, cx, src, dst, flips):
        p = lambda a: a.ctypes.data_as(_i32p)  # noqa: E731
        check(self._lib.fpb_paths_toggle(self._h, len(trial), p(trial), p(cx), p(src), p(dst),
                                         flips.ctypes.data_as(_u32p)))


    def knn_len(self, complex_index, k):
        import numpy as np

        s = self.sizes()
        n = s["total_odd"][complex_index]
        ids = np.empty((n, k), np.int32)
        lens = np.empty((n, k), np.float64)
        wmax = np.empty(s["n_trials"], np.float64)
        check(self._lib.fpb_knn_len(self._h, complex_index, k, ids.ctypes.data, lens.ctypes.data, wmax.ctypes.data))
        return ids, lens, wmax

    def knn_len_into(self, complex_index, k, ids, lens, wmax):
        s = self.sizes()
        need = s["total_odd"][complex_index] * k
        if ids.size < need or lens.size < need or wmax.size < s["n_trials"]:
            raise FpbError("knn_len_into: buffers too small")
        check(self._lib.fpb_knn_len(self._h, complex_index, k, ids.ctypes.data, lens.ctypes.data, wmax.ctypes.data))

    def pairs_at(self, complex_index, trial, a, b):
        import numpy as np

        out = np.empty(len(trial), np.float64)
        p = lambda x: x.ctypes.data_as(_i32p)  # noqa: E731
        check(self._lib.fpb_pairs_at(self._h, complex_index, len(trial), p(trial), p(a), p(b),
                                     out.ctypes.data_as(_f64p)))
        return out

    def pairs_of_trial(self, complex_index, trial, n_pairs):
        import numpy as np

        out = np.empty(n_pairs, np.float64)
        check(self._lib.fpb_pairs_of_trial(self._h, complex_index, trial, out.ctypes.data_as(_f64p), n_pairs))
        return out

    def price(self, complex_index, y, top, ztop, scale, wshift, active, cap=65536):
        import numpy as np

        f = lambda x: x.ctypes.data_as(_f64p)  # noqa: E731
        i64 = C.POINTER(C.c_int64)
        topp = top.ctypes.data_as(_i32p) if top is not None else None
        ztp = ztop.ctypes.data_as(i64) if top is not None else None
        while True:
            vt, va, vb = (np.empty(cap, np.int32) for _ in range(3))
            vl = np.empty(cap, np.float64)
            found = C.c_int64()
            check(self._lib.fpb_price(self._h, complex_index, y.ctypes.data_as(i64), topp, ztp, f(scale),
                                      wshift.ctypes.data_as(_i32p), active.ctypes.data_as(_u8p), cap,
                                      vt.ctypes.data_as(_i32p), va.ctypes.data_as(_i32p), vb.ctypes.data_as(_i32p),
                                      f(vl), C.byref(found)))
            n = int(found.value)
            if n <= cap:
                return vt[:n], va[:n], vb[:n], vl[:n]
            cap = n + n // 8 + 64

    def paths_list(self, trial, cx, src, dst, max_len):
        import numpy as np

        n = len(trial)
        qubits = np.empty((n, max_len), np.int32)
        length = np.empty(n, np.int32)
        p = lambda a: a.ctypes.data_as(_i32p)  # noqa: E731
        check(self._lib.fpb_paths_list(self._h, n, p(trial), p(cx), p(src), p(dst), max_len, p(qubits), p(length)))
        return qubits, length


# ---------------------------------------------------------------------------------------------
# Device Union-Find decoder: plain C-ABI calls on the handle's address, the same for both bindings
# ---------------------------------------------------------------------------------------------
def uf_set_graph(handle, complex_index: int, n_nodes: int, edge_a, edge_b):
    """``fpb_uf_set_graph`` on a handle of either binding (int32 arrays, one entry per syndrome qubit)."""
    check(load().fpb_uf_set_graph(C.c_void_p(handle.address()), complex_index, int(n_nodes),
                                  edge_a.ctypes.data_as(_i32p), edge_b.ctypes.data_as(_i32p)))


def uf_decode(handle, use_erasures: bool, flips, grown=None) -> int:
    """``fpb_uf_decode``: fills ``flips`` (and ``grown`` if given) uint32[T, bit_words]; returns the number of
    stuck items."""
    stuck = C.c_int32()
    check(load().fpb_uf_decode(C.c_void_p(handle.address()), int(bool(use_erasures)),
                               flips.ctypes.data_as(_u32p) if flips is not None else None,
                               grown.ctypes.data_as(_u32p) if grown is not None else None, C.byref(stuck)))
    return int(stuck.value)


def set_planes(handle, plane_off, plane_qubits):
    """``fpb_set_planes``: int32 offsets [n_planes + 1] and global qubit ids of the checked correlation surfaces."""
    check(load().fpb_set_planes(C.c_void_p(handle.address()), len(plane_off) - 1, plane_off.ctypes.data_as(_i32p),
                                plane_qubits.ctypes.data_as(_i32p)))


def logical_check(handle, fail=None) -> int:
    """``fpb_logical_check``: fills ``fail`` uint8[T] if given; returns the number of failed trials."""
    n = C.c_int64()
    check(load().fpb_logical_check(C.c_void_p(handle.address()), fail.ctypes.data_as(_u8p) if fail is not None else None,
                                   C.byref(n)))
    return int(n.value)


_fpbpy = None
_fpbpy_tried = False


def pybind_module():
    """``flamingpy_b200.cpp.fpbpy`` if it is built and importable, else None (import guard in the
    style of ``flamingpy/__init__.py:28-34``)."""
    global _fpbpy, _fpbpy_tried
    if not _fpbpy_tried:
        _fpbpy_tried = True
        try:
            load()  # libfpb.so first: a missing library must raise the clear ImportError above
            from .cpp import fpbpy

            if fpbpy.abi_version() != FPB_ABI_VERSION:
                raise ImportError("fpbpy was built against another ABI version; rebuild it")
            fpbpy.set_error_class(FpbError)
            _fpbpy = fpbpy
        except ImportError as exc:
            import warnings

            warnings.warn(f"Failed to import the fpbpy pybind11 module ({exc}); using ctypes.", ImportWarning)
    return _fpbpy


def binding_name() -> str:
    """Which binding ``open_handle`` uses: "pybind11" or "ctypes"."""
    if os.environ.get("FPB_BINDING", "").lower() == "ctypes":
        return "ctypes"
    return "pybind11" if pybind_module() is not None else "ctypes"


def open_handle(cfg: dict):
    """Create an ``fpb_handle`` from a config dict (the fields of ``fpb_config`` by name, with
    ``complexes`` a list of dicts of the fields of ``fpb_complex_desc``)."""
    load()
    if binding_name() == "pybind11":
        return pybind_module().Handle(cfg)
    return CtypesHandle(cfg)
