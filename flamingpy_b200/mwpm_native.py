"""ctypes binding of the native minimum-weight perfect matcher (``csrc/mwpm.cpp`` ->
``libfpb_mwpm.so``), the host-side replacement of the reference's LEMON plugin
(``flamingpy/cpp/lemonpy.cpp``, ``decoders/mwpm/lemon.py:25-40``)."""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

LIB_PATH = os.environ.get("FPB_MWPM_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfpb_mwpm.so")
_lib = None

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: run __graft_entry__.build()")
        lib = C.CDLL(LIB_PATH)
        lib.fpb_mwpm.restype = C.c_int
        lib.fpb_mwpm.argtypes = [C.c_int, _f64p, _f64p, _i32p, _i32p]
        lib.fpb_mwpm_batch.restype = C.c_int
        lib.fpb_mwpm_batch.argtypes = [C.c_int, _i32p, _f64p, _f64p, _i64p, _i32p, _i32p, _i32p, C.c_int]
        lib.fpb_mwpm_batch_cand.restype = C.c_int
        lib.fpb_mwpm_batch_cand.argtypes = [C.c_int, _i32p, _f64p, _f64p, _i64p, _i32p, C.c_int, _i32p, _i32p, _i32p,
                                            C.c_int]
        _u8p = C.POINTER(C.c_uint8)
        lib.fpb_mwpm_sparse_begin.restype = C.c_void_p
        lib.fpb_mwpm_sparse_begin.argtypes = [C.c_int, _i32p, _i64p, _i32p, _f64p, C.c_int, _f64p, _f64p, C.c_int]
        lib.fpb_mwpm_sparse_requests.restype = C.c_longlong
        lib.fpb_mwpm_sparse_requests.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, C.c_longlong]
        lib.fpb_mwpm_sparse_start.restype = C.c_int
        lib.fpb_mwpm_sparse_start.argtypes = [C.c_void_p, _f64p]
        lib.fpb_mwpm_sparse_solve.restype = C.c_int
        lib.fpb_mwpm_sparse_solve.argtypes = [C.c_void_p, _i64p, _i32p, _i64p, _f64p, _i32p, _u8p]
        lib.fpb_mwpm_sparse_apply.restype = C.c_int
        lib.fpb_mwpm_sparse_apply.argtypes = [C.c_void_p, C.c_longlong, _i32p, _i32p, _i32p, _f64p, C.c_int]
        lib.fpb_mwpm_sparse_finish.restype = C.c_int
        lib.fpb_mwpm_sparse_finish.argtypes = [C.c_void_p, _i32p, _i32p, _i32p, _u8p]
        lib.fpb_mwpm_sparse_free.restype = None
        lib.fpb_mwpm_sparse_free.argtypes = [C.c_void_p]
        lib.fpb_mwpm_max_threads.restype = C.c_int
        lib.fpb_mwpm_configure.restype = None
        lib.fpb_mwpm_configure.argtypes = [C.c_int, C.c_int]
        lib.fpb_mwpm_stats.restype = None
        lib.fpb_mwpm_stats.argtypes = [_i64p, C.c_int]
        _lib = lib
    return _lib


MODES = {"auto": 0, "dense": 1, "sparse": 2}
SPARSE_MIN_K = 48  # graphs from this many odd cubes on go to the sparse solver (csrc/mwpm.cpp)
CAND_K = 12        # candidate partners per cube requested from the device (= the solver's default knn)


def configure(mode: str = "auto", knn: int = 0) -> None:
    """Choose the driver: ``"auto"`` (sparse candidate graph + pricing from 48 cubes on, every edge
    of the complete graph below), ``"dense"`` (the complete graph for every K) or ``"sparse"``;
    ``knn`` = nearest partners per cube in the first candidate graph (0 keeps the current value).
    Each returns a minimum-weight perfect matching of the same complete graph; only the choice among
    ties can differ."""
    load().fpb_mwpm_configure(MODES[mode], int(knn))


def stats(reset: bool = False) -> dict:
    out = np.zeros(8, np.int64)
    load().fpb_mwpm_stats(_p(out, _i64p), int(reset))
    d = dict(zip(("sparse_solves", "retries", "pricing_rounds", "priced_edges"), out[:4].tolist()))
    d.update(zip(("scan_ms", "build_ms", "stages_ms", "pricing_ms"), (out[4:] / 1e6).round(1).tolist()))  # summed over threads
    return d


def _p(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


def default_threads() -> int:
    """Host cores this process may use for matching: its CPU affinity, shared evenly between the
    ranks of a one-process-per-GPU job on this node (``LOCAL_WORLD_SIZE``, set by torchrun)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    return max(1, n // local)


def solve(K: int, pair_len: np.ndarray, bnd_len: Optional[np.ndarray]):
    """Matched queries [(u, v)] with v == -1 for "to its connector" (one matching graph)."""
    if K == 0:
        return []
    lib = load()
    pl = np.ascontiguousarray(pair_len, np.float64)
    bl = None if bnd_len is None else np.ascontiguousarray(bnd_len, np.float64)
    src = np.empty(max(K, 1), np.int32)
    dst = np.empty(max(K, 1), np.int32)
    n = lib.fpb_mwpm(K, _p(pl, _f64p), _p(bl, _f64p), _p(src, _i32p), _p(dst, _i32p))
    if n < 0:
        raise ValueError(f"native matcher failed ({n}): no perfect matching or weights out of range")
    return [(int(src[i]), int(dst[i])) for i in range(n)]


def min_weight_perfect_matching(K, pair_len, bnd_len):
    """mate dict in the node numbering of ``matching.solve_matching`` (virtual node of u is K+u)."""
    mate = {}
    for u, v in solve(K, pair_len, bnd_len):
        if v < 0:
            mate[u] = K + u
        else:
            mate[u], mate[v] = v, u
    return mate


def solve_batch(odd_count: np.ndarray, pair_len: np.ndarray, bnd_len: Optional[np.ndarray], threads: int = 0,
                cand: Optional[np.ndarray] = None):
    """All trials of a batch at once (OpenMP over trials).  Returns (trial, src, dst) int32 arrays
    of the matched queries, ready for ``TrialEngine.paths_toggle``.  ``cand``: int32[sum K, k]
    candidate partners from ``TrialEngine.knn`` (optional; the result is the same exact optimum)."""
    lib = load()
    T = len(odd_count)
    oc = np.ascontiguousarray(odd_count, np.int32)
    q_off = np.concatenate([[0], np.cumsum(oc.astype(np.int64))]).astype(np.int64)
    total = int(q_off[-1])
    src = np.empty(max(total, 1), np.int32)
    dst = np.empty(max(total, 1), np.int32)
    nq = np.zeros(T, np.int32)
    pl = np.ascontiguousarray(pair_len, np.float64)
    bl = None if (bnd_len is None or len(bnd_len) == 0) else np.ascontiguousarray(bnd_len, np.float64)
    if threads <= 0:
        threads = default_threads()
    if cand is not None and len(cand):
        cand = np.ascontiguousarray(cand, np.int32)
        if cand.ndim != 2 or cand.shape[0] != total:
            raise ValueError("cand must have one row per odd cube")
        rc = lib.fpb_mwpm_batch_cand(T, _p(oc, _i32p), _p(pl, _f64p), _p(bl, _f64p), _p(q_off, _i64p),
                                     _p(cand, _i32p), cand.shape[1], _p(src, _i32p), _p(dst, _i32p), _p(nq, _i32p),
                                     threads)
    else:
        rc = lib.fpb_mwpm_batch(T, _p(oc, _i32p), _p(pl, _f64p), _p(bl, _f64p), _p(q_off, _i64p), _p(src, _i32p),
                                _p(dst, _i32p), _p(nq, _i32p), threads)
    if rc != 0:
        raise ValueError("native matcher failed on at least one trial")
    keep = np.concatenate([np.arange(q_off[t], q_off[t] + nq[t]) for t in range(T)]) if T else np.empty(0, np.int64)
    keep = keep.astype(np.int64)
    trial = np.repeat(np.arange(T, dtype=np.int32), nq)
    return trial, src[keep], dst[keep]


SPARSE_MAX_ROUNDS = 3  # warm restarts per trial before it is finished from its full triangle instead


def solve_batch_sparse(eng, complex_index: int, odd_count: np.ndarray, bnd_len: Optional[np.ndarray],
                       threads: int = 0, stats: Optional[dict] = None, max_rounds: int = 0, knn=None):
    """All trials of the engine's last graph-stage run, matched exactly **without fetching the
    K(K-1)/2 lengths**: the device hands over each cube's nearest partners with their reduced
    lengths (``fpb_knn_len``), the lengths of the few pairs the solver asks for (``fpb_pairs_at``) and,
    after every solve, the pairs that violate dual feasibility under the solver's duals
    (``fpb_price``, the pricing pass run where the lengths live).  Same result as ``solve_batch``
    (a minimum-weight perfect matching of the complete graph; ties may be broken differently).
    Returns (trial, src, dst) int32 arrays for ``TrialEngine.paths_toggle``."""
    import time

    lib = load()
    T = len(odd_count)
    oc = np.ascontiguousarray(odd_count, np.int32)
    q_off = np.concatenate([[0], np.cumsum(oc.astype(np.int64))]).astype(np.int64)
    total = int(q_off[-1])
    if threads <= 0:
        threads = default_threads()
    if total == 0:
        e = np.empty(0, np.int32)
        return e, e.copy(), e.copy()
    tm = {"knn_ms": 0.0, "begin_ms": 0.0, "request_ms": 0.0, "solve_ms": 0.0, "price_ms": 0.0, "apply_ms": 0.0,
          "fallback_ms": 0.0}
    clock = [time.perf_counter()]

    def lap(key):
        now = time.perf_counter()
        tm[key] += (now - clock[0]) * 1e3
        clock[0] = now

    # knn: (ids, lens, wmax) already fetched by the caller (e.g. overlapped with the previous batch's matching)
    cand, cand_len, wmax = knn if knn is not None else eng.knn_len(complex_index, CAND_K)
    lap("knn_ms")
    bl = None if (bnd_len is None or len(bnd_len) == 0) else np.ascontiguousarray(bnd_len, np.float64)
    ctx = lib.fpb_mwpm_sparse_begin(T, _p(oc, _i32p), _p(q_off, _i64p), _p(cand, _i32p), _p(cand_len, _f64p), CAND_K,
                                    _p(bl, _f64p), _p(wmax, _f64p), threads)
    u8 = C.POINTER(C.c_uint8)
    lap("begin_ms")
    try:
        n_req = int(lib.fpb_mwpm_sparse_requests(ctx, None, None, None, 0))
        rt, ra, rb = (np.empty(max(n_req, 1), np.int32) for _ in range(3))
        lib.fpb_mwpm_sparse_requests(ctx, _p(rt, _i32p), _p(ra, _i32p), _p(rb, _i32p), n_req)
        req_len = eng.pairs_at(complex_index, rt[:n_req], ra[:n_req], rb[:n_req]) if n_req else np.empty(1)
        lap("request_ms")
        lib.fpb_mwpm_sparse_start(ctx, _p(np.ascontiguousarray(req_len, np.float64), _f64p))
        y = np.zeros(total, np.int64)
        top = np.zeros(total, np.int32)
        ztop = np.zeros(total, np.int64)
        scale = np.zeros(T, np.float64)
        wshift = np.zeros(T, np.int32)
        active = np.zeros(T, np.uint8)
        rounds = flagged = 0
        while True:
            n_active = lib.fpb_mwpm_sparse_solve(ctx, _p(y, _i64p), _p(top, _i32p), _p(ztop, _i64p), _p(scale, _f64p),
                                                 _p(wshift, _i32p), active.ctypes.data_as(u8))
            lap("solve_ms")
            if n_active == 0:
                break
            vt, va, vb, vl = eng.price(complex_index, y, scale, wshift, active, cap=max(65536, 8 * total), top=top,
                                       ztop=ztop)
            vt, va, vb = (np.ascontiguousarray(v, np.int32) for v in (vt, va, vb))
            vl = np.ascontiguousarray(vl, np.float64)
            rounds += 1
            flagged += len(vt)
            lap("price_ms")
            lib.fpb_mwpm_sparse_apply(ctx, len(vt), _p(vt, _i32p), _p(va, _i32p), _p(vb, _i32p), _p(vl, _f64p),
                                      max_rounds or SPARSE_MAX_ROUNDS)
            lap("apply_ms")
        src = np.empty(max(total, 1), np.int32)
        dst = np.empty(max(total, 1), np.int32)
        nq = np.zeros(T, np.int32)
        failed = np.zeros(T, np.uint8)
        n_failed = lib.fpb_mwpm_sparse_finish(ctx, _p(src, _i32p), _p(dst, _i32p), _p(nq, _i32p),
                                              failed.ctypes.data_as(u8))
    finally:
        lib.fpb_mwpm_sparse_free(ctx)
    if n_failed:
        # still violated after max_rounds warm restarts (or a candidate graph without a perfect
        # matching): these few trials are finished from their full triangles, in parallel
        ft = np.nonzero(failed)[0]
        tris = [eng.pairs_of_trial(complex_index, int(t), int(oc[t])) for t in ft]
        fb = None if bl is None else np.concatenate([bl[q_off[t]: q_off[t + 1]] for t in ft])
        r_t, r_s, r_d = solve_batch(oc[ft], np.concatenate(tris), fb, threads)
        for i, t in enumerate(ft):
            sel = r_t == i
            nq[t] = int(sel.sum())
            src[q_off[t]: q_off[t] + nq[t]] = r_s[sel]
            dst[q_off[t]: q_off[t] + nq[t]] = r_d[sel]
    lap("fallback_ms")
    if stats is not None:
        stats.update(pricing_rounds=rounds, flagged_pairs=flagged, requested_pairs=n_req, fallback_trials=int(n_failed),
                     **{k: round(v, 1) for k, v in tm.items()})
    keep = np.concatenate([np.arange(q_off[t], q_off[t] + nq[t]) for t in range(T)]).astype(np.int64)
    trial = np.repeat(np.arange(T, dtype=np.int32), nq)
    return trial, src[keep], dst[keep]
