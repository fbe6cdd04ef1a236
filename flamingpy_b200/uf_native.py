"""ctypes binding of the host-side Union-Find + peeling decoder (``csrc/uf.cpp`` -> ``libfpb_uf.so``,
C ABI in ``include/fpb_uf.h``), the batch replacement of the reference's pure-Python
``flamingpy/decoders/unionfind`` package, and the static graph it runs on."""

from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import numpy as np

LIB_PATH = os.environ.get("FPB_UF_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfpb_uf.so")
_lib = None
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run __graft_entry__.build()")
        lib = ctypes.CDLL(LIB_PATH)
        lib.fpb_uf_batch.restype = ctypes.c_int
        lib.fpb_uf_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _i32p, _i32p, ctypes.c_int, _u8p, _u8p,
                                     _u8p, _u8p, ctypes.c_int]
        lib.fpb_uf_max_threads.restype = ctypes.c_int
        lib.fpb_uf_erasures.restype = ctypes.c_longlong
        lib.fpb_uf_erasures.argtypes = [ctypes.c_int, ctypes.c_int, _i32p, _i32p, _i32p, ctypes.c_int, _u8p, _u8p,
                                        ctypes.c_int]
        lib.fpb_uf_check_correction.restype = ctypes.c_longlong
        u32p = ctypes.POINTER(ctypes.c_uint32)
        lib.fpb_uf_check_correction.argtypes = [ctypes.c_int, ctypes.c_int, _u8p, u32p, _u8p, u32p, ctypes.c_int, _i32p, _i32p,
                                                _u8p, ctypes.c_int]
        _lib = lib
    return _lib


def check_correction(bits: np.ndarray, flips: np.ndarray, plane_off: np.ndarray, plane_qubits: np.ndarray,
                     n_qubits: Optional[int] = None, threads: int = 0) -> np.ndarray:
    """bool[T]: some plane has odd parity of bit value XOR recovery.  ``bits`` and ``flips`` each uint8[T, Q] or the
    packed uint32[T, ceil(Q/32)] words of the device; planes as int32 offsets + global qubit ids.  ``n_qubits`` is
    needed when both come packed."""
    u32p = ctypes.POINTER(ctypes.c_uint32)

    def split(a, name):
        packed = a.dtype == np.uint32
        a = np.ascontiguousarray(a, np.uint32 if packed else np.uint8)
        if a.ndim != 2:
            raise ValueError(f"{name} must be [trials, qubits] bytes or [trials, ceil(qubits / 32)] words")
        return a, packed

    bits, bp = split(bits, "bits")
    flips, fp = split(flips, "flips")
    T = flips.shape[0]
    Q = int(n_qubits) if n_qubits is not None else (flips.shape[1] if not fp else (bits.shape[1] if not bp else -1))
    if Q < 0:
        raise ValueError("n_qubits is needed when bits and flips are both packed")
    for a, packed, name in ((bits, bp, "bits"), (flips, fp, "flips")):
        if a.shape != (T, (Q + 31) // 32 if packed else Q):
            raise ValueError(f"{name} has shape {a.shape}, expected {T} trials of {Q} qubits")
    if len(plane_qubits) and int(plane_qubits.max()) >= Q:
        raise ValueError("plane qubit out of range")
    fail = np.empty(T, np.uint8)
    load().fpb_uf_check_correction(T, Q, None if bp else bits.ctypes.data_as(_u8p), bits.ctypes.data_as(u32p) if bp else None,
                                   None if fp else flips.ctypes.data_as(_u8p), flips.ctypes.data_as(u32p) if fp else None,
                                   len(plane_off) - 1, plane_off.ctypes.data_as(_i32p), plane_qubits.ctypes.data_as(_i32p),
                                   fail.ctypes.data_as(_u8p), int(threads))
    return fail.astype(bool)


def uf_graph(ct) -> Tuple[int, np.ndarray, np.ndarray]:
    """The stabilizer graph of one complex without its 'low' / 'high' connectors, as the decoder's
    edge list: ``(n_nodes, edge_a, edge_b)`` with one entry per syndrome qubit (local id into
    ``ct.qubits``).  Nodes ``[0, S)`` are the cubes; a qubit on a boundary face joins its cube to a
    boundary point of its own (boundary points are leaves of the reference's graph,
    ``surface_code.py:477-525``), numbered from ``S`` in qubit order.  A qubit with no edge (the
    dropped half of a doubly shared face, ``stabilizer_graph.py:295-298``) has ``edge_a = -1``."""
    cached = getattr(ct, "_uf_graph", None)
    if cached is not None:
        return cached
    S, Q = ct.S, ct.Q
    ea = np.full(Q, -1, np.int32)
    eb = np.full(Q, -1, np.int32)
    is_bnd = np.zeros(Q, bool)
    for f in range(6):
        q = ct.cube_qubit[:, f]
        nb = ct.nbr[:, f]
        for s in np.nonzero(q >= 0)[0]:
            qq, o = int(q[s]), int(nb[s])
            if o < 0:
                continue
            if o >= S:  # LOW / HIGH side: the qubit itself is the boundary point
                ea[qq], is_bnd[qq] = s, True
            elif ea[qq] < 0:
                ea[qq], eb[qq] = min(s, o), max(s, o)
    bq = np.nonzero(is_bnd)[0]
    eb[bq] = S + np.arange(len(bq), dtype=np.int32)
    out = (S + len(bq), ea, eb)
    try:
        ct._uf_graph = out
    except AttributeError:  # pragma: no cover - frozen tables
        pass
    return out


def decode_batch(ct, parity: np.ndarray, erased: Optional[np.ndarray] = None, threads: int = 0, want_grown: bool = False):
    """Recovery of every trial: ``parity`` uint8[T, S] (1 = unsatisfied cube), ``erased`` uint8[T, Q] or
    None; returns uint8[T, Q] bit flips (and the fully grown edges when ``want_grown``)."""
    n_nodes, ea, eb = uf_graph(ct)
    parity = np.ascontiguousarray(parity, np.uint8)
    T = parity.shape[0]
    if parity.shape != (T, ct.S):
        raise ValueError("parity must be [trials, cubes]")
    if erased is not None:
        erased = np.ascontiguousarray(erased, np.uint8)
        if erased.shape != (T, ct.Q):
            raise ValueError("erased must be [trials, qubits]")
    flips = np.zeros((T, ct.Q), np.uint8)
    grown = np.zeros((T, ct.Q), np.uint8) if want_grown else None
    rc = load().fpb_uf_batch(n_nodes, ct.S, ct.Q, ea.ctypes.data_as(_i32p), eb.ctypes.data_as(_i32p), T,
                             parity.ctypes.data_as(_u8p), erased.ctypes.data_as(_u8p) if erased is not None else None,
                             flips.ctypes.data_as(_u8p), grown.ctypes.data_as(_u8p) if want_grown else None, int(threads))
    if rc:
        raise ValueError(f"trial {-rc - 1}: an odd cluster can neither grow nor reach a boundary point "
                         "(odd number of unsatisfied stabilizers on a complex without boundary)")
    return (flips, grown) if want_grown else flips


def erasures(lat, qubit_nodes: np.ndarray, state: np.ndarray, threads: int = 0):
    """Erased edges uint8[T, Q] (None when there is none) for the syndrome qubits ``qubit_nodes``
    (lattice node indices) from the state labels uint8[T, N]: two or more p-squeezed neighbours."""
    state = np.ascontiguousarray(np.atleast_2d(state), np.uint8)
    qn = np.ascontiguousarray(qubit_nodes, np.int32)
    indptr = np.ascontiguousarray(lat.adj_indptr, np.int32)
    indices = np.ascontiguousarray(lat.adj_indices, np.int32)
    out = np.empty((state.shape[0], len(qn)), np.uint8)
    n = load().fpb_uf_erasures(lat.N, len(qn), qn.ctypes.data_as(_i32p), indptr.ctypes.data_as(_i32p),
                               indices.ctypes.data_as(_i32p), state.shape[0], state.ctypes.data_as(_u8p),
                               out.ctypes.data_as(_u8p), int(threads))
    return out if n else None
