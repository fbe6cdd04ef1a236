"""Batched trial engine: the Python face of the C ABI (``include/fpb.h``).

One ``TrialEngine`` = one code instance + noise parameters + weight options on
one GPU.  It runs batches of independent Monte-Carlo trials through
noise -> bits -> weights -> syndrome -> matching graph on the device and hands
back NumPy arrays.  The reference-compatible classes in ``codes.py``,
``noise.py``, ``decoder.py`` and ``simulations.py`` are thin layers over this.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from .lattice import LatticeTables


def _ptr(a: np.ndarray, typ):
    return a.ctypes.data_as(typ)


def build_config(lat: LatticeTables, delta: float, p_swap: float, wo: dict, mode: str = "compat",
                 chain_len: int = 0, device: int = 0, delta_step: float = 0.0):
    """Fill an ``fpb_config`` (include/fpb.h) from lattice tables and options.
    Returns (config, keep-alive list of arrays, qubit offsets per complex, Qtot)."""
    keepalive = []

    def keep(a, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        keepalive.append(a)
        return a

    cfg = _lib.Config()
    cfg.abi_version = _lib.FPB_ABI_VERSION
    cfg.device = device
    cfg.n_nodes = lat.N
    cfg.adj_indptr = _ptr(keep(lat.adj_indptr, np.int32), _lib._i32p)
    cfg.adj_indices = _ptr(keep(lat.adj_indices, np.int32), _lib._i32p)
    cfg.adj_vals = _ptr(keep(lat.adj_vals, np.int8), _lib._i8p)
    cfg.n_complex = len(lat.ec)
    q_off = {}
    off = 0
    for i, ec in enumerate(lat.ec):
        ct = lat.complexes[ec]
        d = cfg.complex_[i]
        d.n_cubes = ct.S
        d.n_qubits = ct.Q
        d.qubit_node = _ptr(keep(ct.qubits, np.int32), _lib._i32p)
        d.cube_qubit = _ptr(keep(ct.cube_qubit, np.int32), _lib._i32p)
        d.ninfo = _ptr(keep(ct.ninfo, np.uint16), _lib._u16p)
        d.sbase = _ptr(keep(ct.sbase, np.int32), _lib._i32p)
        for j, v in enumerate(np.asarray(ct.d_off, np.int32).ravel()):
            d.d_off[j] = int(v)
        for j, v in enumerate(np.asarray(ct.s_off, np.int32).ravel()):
            d.s_off[j] = int(v)
        d.n_slots = ct.n_slots
        d.slot_qubit = _ptr(keep(ct.slot_qubit, np.int32), _lib._i32p)
        d.n_low = len(ct.low_seed_cube)
        d.low_cube = _ptr(keep(ct.low_seed_cube, np.int32), _lib._i32p)
        d.low_slot = _ptr(keep(ct.low_seed_slot, np.int32), _lib._i32p)
        d.n_high = len(ct.high_seed_cube)
        d.high_cube = _ptr(keep(ct.high_seed_cube, np.int32), _lib._i32p)
        d.high_slot = _ptr(keep(ct.high_seed_slot, np.int32), _lib._i32p)
        q_off[ec] = off
        off += ct.Q
    cfg.delta = float(delta)
    cfg.p_swap = float(p_swap or 0.0)
    cfg.weight_method = _lib.WEIGHT_BLUEPRINT if wo["method"] == "blueprint" else _lib.WEIGHT_UNIFORM
    cfg.weight_delta = float(wo.get("delta") or delta)
    cfg.weight_integer = int(bool(wo.get("integer", False)))
    cfg.weight_multiplier = float(wo.get("multiplier", 1))
    cfg.label_mode = _lib.MODE_FRESH if mode == "fresh" else _lib.MODE_COMPAT
    cfg.chain_len = int(chain_len)
    cfg.delta_step = float(delta_step)
    return cfg, keepalive, q_off, off


@dataclass
class ComplexBatch:
    """Per-complex results of a batch (ragged arrays are concatenated over
    trials with ``*_off`` offsets)."""

    parity: Optional[np.ndarray] = None  # uint8[T, S]
    odd_count: Optional[np.ndarray] = None  # int32[T]
    odd_off: Optional[np.ndarray] = None  # int64[T + 1]
    odd_ids: Optional[np.ndarray] = None  # int32[sum K]
    pair_off: Optional[np.ndarray] = None  # int64[T + 1]
    pair_len: Optional[np.ndarray] = None  # float64[sum K(K-1)/2]
    bnd_len: Optional[np.ndarray] = None  # float64[sum K]
    bnd_side: Optional[np.ndarray] = None  # uint8[sum K]

    def odd(self, t: int) -> np.ndarray:
        return self.odd_ids[self.odd_off[t] : self.odd_off[t + 1]]

    def pairs(self, t: int) -> np.ndarray:
        return self.pair_len[self.pair_off[t] : self.pair_off[t + 1]]

    def boundary(self, t: int):
        sl = slice(self.odd_off[t], self.odd_off[t + 1])
        return self.bnd_len[sl], self.bnd_side[sl]


@dataclass
class TrialBatch:
    n_trials: int
    stage: int
    hom: Optional[np.ndarray] = None  # float64[T, Qtot]
    bits: Optional[np.ndarray] = None  # uint8[T, Qtot]
    weights: Optional[np.ndarray] = None  # float64[T, Qtot]
    state: Optional[np.ndarray] = None  # uint8[T, N]
    x: Optional[np.ndarray] = None  # float64[T, 2N]
    complexes: Optional[Dict[str, ComplexBatch]] = None
    timings: Optional[dict] = None


class TrialEngine:
    """Owns an ``fpb_handle``.  ``weight_opts`` follows the reference's
    ``weight_options`` dict (``decoders/decoder.py:31-56``): ``method``
    ("uniform" | "blueprint"), ``integer``, ``multiplier``, ``delta``."""

    def __init__(
        self,
        lat: LatticeTables,
        delta: float,
        p_swap=None,
        weight_opts: Optional[dict] = None,
        mode: str = "compat",
        chain_len: int = 0,
        device: int = 0,
        delta_step: float = 0.0,
    ):
        lib = _lib.load()
        self._lib = lib
        self.lat = lat
        self.delta = float(delta)
        self.p_swap = float(p_swap or 0.0)
        wo = {"method": "uniform", "integer": False, "multiplier": 1}
        wo.update(weight_opts or {})
        if wo["method"] not in ("uniform", "blueprint"):
            raise ValueError(f"unknown weight method {wo['method']!r}")
        if mode not in ("compat", "fresh"):
            raise ValueError("mode must be 'compat' or 'fresh'")
        self.weight_opts = wo
        self.mode = mode
        self.device = device
        self.ec = list(lat.ec)
        cfg, self._keep, self.q_off, self.Qtot = build_config(
            lat, self.delta, self.p_swap, wo, mode, chain_len, device, delta_step)
        self.bit_words = (self.Qtot + 31) // 32
        handle = C.c_void_p()
        _lib.check(lib.fpb_create(C.byref(cfg), C.byref(handle)))
        self._h = handle
        self._keep = []
        self.uses_table = False

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.fpb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ runs
    def run_injected(self, x: np.ndarray, state: np.ndarray, stage: int = _lib.STAGE_GRAPH,
                     fetch: bool = True) -> TrialBatch:
        """Run trials on injected draws ``x`` [T, 2N] and labels ``state`` [T, N]."""
        x = np.ascontiguousarray(x, np.float64)
        state = np.ascontiguousarray(state, np.uint8)
        if x.ndim != 2 or x.shape[1] != 2 * self.lat.N:
            raise ValueError(f"x must have shape [T, {2 * self.lat.N}]")
        if state.shape != (x.shape[0], self.lat.N):
            raise ValueError(f"state must have shape [T, {self.lat.N}]")
        T = x.shape[0]
        _lib.check(self._lib.fpb_run_injected(self._h, T, x.ctypes.data, state.ctypes.data, 0, stage))
        return self.fetch(stage) if fetch else TrialBatch(T, stage, timings=self.timings())

    def run_rng(self, seed: int, first_trial: int, n_trials: int, stage: int = _lib.STAGE_GRAPH,
                fetch: bool = True, want_draws: bool = False) -> TrialBatch:
        """Run trials ``[first_trial, first_trial + n_trials)`` of the Philox stream ``seed``."""
        _lib.check(self._lib.fpb_run_rng(self._h, seed & (2**64 - 1), first_trial, n_trials, stage))
        return self.fetch(stage, want_draws=want_draws) if fetch else TrialBatch(
            n_trials, stage, timings=self.timings())

    def run_bits(self, bits, stage: int = _lib.STAGE_GRAPH, fetch: bool = True) -> TrialBatch:
        """Trials given directly as bit values of the syndrome qubits - the i.i.d. Pauli-noise model
        (``flamingpy/noise/iid.py``) or any other bit-level sampler.  ``bits``: uint8[T, Qtot] (0/1)
        or already packed uint32[T, bit_words] (LSB first).  Needs uniform weights."""
        bits = np.asarray(bits)
        if bits.dtype != np.uint32:
            if bits.ndim != 2 or bits.shape[1] != self.Qtot:
                raise ValueError(f"bits must have shape (n_trials, {self.Qtot})")
            pad = np.zeros((bits.shape[0], self.bit_words * 32), np.uint8)
            pad[:, : self.Qtot] = bits != 0
            bits = np.packbits(pad, axis=1, bitorder="little").view(np.uint32)
        if bits.ndim != 2 or bits.shape[1] != self.bit_words:
            raise ValueError(f"packed bits must have shape (n_trials, {self.bit_words})")
        bits = np.ascontiguousarray(bits)
        n_trials = bits.shape[0]
        _lib.check(self._lib.fpb_run_bits(self._h, n_trials, bits.ctypes.data, 0, stage))
        return self.fetch(stage) if fetch else TrialBatch(n_trials, stage, timings=self.timings())

    def run_iid(self, seed: int, first_trial: int, n_trials: int, p_err: float, stage: int = _lib.STAGE_GRAPH,
                fetch: bool = True) -> TrialBatch:
        """Trials ``[first_trial, first_trial + n_trials)`` of i.i.d. Z errors with probability
        ``p_err`` per qubit, drawn on the device (Philox stream ``seed``).  Needs uniform weights."""
        _lib.check(self._lib.fpb_run_iid(self._h, seed & (2**64 - 1), first_trial, n_trials, float(p_err), stage))
        return self.fetch(stage) if fetch else TrialBatch(n_trials, stage, timings=self.timings())

    def sizes(self) -> _lib.Sizes:
        s = _lib.Sizes()
        _lib.check(self._lib.fpb_get_sizes(self._h, C.byref(s)))
        return s

    def timings(self) -> dict:
        t = _lib.Timings()
        _lib.check(self._lib.fpb_get_timings(self._h, C.byref(t)))
        return {name: getattr(t, name) for name, _ in _lib.Timings._fields_}

    def build_table(self):
        """Precompute the all-pairs distance table for trial-independent weights (uniform, or
        blueprint with ``p_swap == 1``); later graph-stage runs gather from it."""
        _lib.check(self._lib.fpb_build_table(self._h))
        self.uses_table = True

    def mark(self, which: int):
        """Record timing mark 0 / 1 on the handle's stream."""
        _lib.check(self._lib.fpb_mark(self._h, which))

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        _lib.check(self._lib.fpb_elapsed_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def device_outputs(self) -> _lib.Outputs:
        o = _lib.Outputs()
        _lib.check(self._lib.fpb_device_outputs(self._h, C.byref(o)))
        return o

    def device_tensors(self) -> dict:
        """Zero-copy ``torch`` views of the last run's result buffers on the GPU (valid until the
        next run of this engine).  PyTorch is used for plumbing only: the tensors alias the
        handle's own device memory through ``__cuda_array_interface__``."""
        import torch

        o = self.device_outputs()
        s = self.sizes()
        T = int(s.n_trials)

        class _View:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False),
                                                 "version": 2, "strides": None}

        def view(ptr, shape, typestr):
            addr = C.cast(ptr, C.c_void_p).value
            if not addr or 0 in shape:
                return None
            return torch.as_tensor(_View(addr, shape, typestr), device=torch.device("cuda", self.device))

        out = {"hom": view(o.hom, (T, self.Qtot), "<f8"), "weights": view(o.weights, (T, self.Qtot), "<f8"),
               "bits": view(o.bits, (T, self.bit_words), "<i4")}
        for i, ec in enumerate(self.ec):
            ct = self.lat.complexes[ec]
            out[ec] = {
                "parity": view(o.parity[i], (T, (ct.S + 31) // 32), "<i4"),
                "odd_count": view(o.odd_count[i], (T,), "<i4"),
                "odd_ids": view(o.odd_ids[i], (int(s.total_odd[i]),), "<i4"),
                "pair_len": view(o.pair_len[i], (int(s.total_pairs[i]),), "<f8"),
                "bnd_len": view(o.bnd_len[i], (int(s.total_odd[i]),), "<f8"),
            }
        return out

    def fetch(self, stage: int = _lib.STAGE_GRAPH, want_draws: bool = False, want_noise: bool = True) -> TrialBatch:
        """Copy the last run's results to host arrays."""
        s = self.sizes()
        T = int(s.n_trials)
        N, Q = self.lat.N, self.Qtot
        out = _lib.Outputs()
        batch = TrialBatch(T, stage)
        packed_bits = None
        if want_noise:
            batch.hom = np.empty((T, Q), np.float64)
            batch.weights = np.empty((T, Q), np.float64)
            packed_bits = np.empty((T, self.bit_words), np.uint32)
            out.hom = _ptr(batch.hom, _lib._f64p)
            out.weights = _ptr(batch.weights, _lib._f64p)
            out.bits = _ptr(packed_bits, _lib._u32p)
        if want_draws:
            batch.state = np.empty((T, N), np.uint8)
            batch.x = np.empty((T, 2 * N), np.float64)
            out.state = _ptr(batch.state, _lib._u8p)
            out.x = _ptr(batch.x, _lib._f64p)
        packed_par: List[Optional[np.ndarray]] = []
        batch.complexes = {}
        for i, ec in enumerate(self.ec):
            ct = self.lat.complexes[ec]
            cb = ComplexBatch()
            pp = None
            if stage >= _lib.STAGE_SYNDROME:
                pw = (ct.S + 31) // 32
                pp = np.empty((T, pw), np.uint32)
                cb.odd_count = np.empty(T, np.int32)
                cb.odd_ids = np.empty(int(s.total_odd[i]), np.int32)
                out.parity[i] = _ptr(pp, _lib._u32p)
                out.odd_count[i] = _ptr(cb.odd_count, _lib._i32p)
                out.odd_ids[i] = _ptr(cb.odd_ids, _lib._i32p)
            if stage >= _lib.STAGE_GRAPH:
                cb.pair_len = np.empty(int(s.total_pairs[i]), np.float64)
                out.pair_len[i] = _ptr(cb.pair_len, _lib._f64p)
                if ct.has_boundary:
                    cb.bnd_len = np.empty(int(s.total_odd[i]), np.float64)
                    cb.bnd_side = np.empty(int(s.total_odd[i]), np.uint8)
                    out.bnd_len[i] = _ptr(cb.bnd_len, _lib._f64p)
                    out.bnd_side[i] = _ptr(cb.bnd_side, _lib._u8p)
                else:
                    cb.bnd_len = np.empty(0, np.float64)
                    cb.bnd_side = np.empty(0, np.uint8)
            packed_par.append(pp)
            batch.complexes[ec] = cb
        _lib.check(self._lib.fpb_fetch(self._h, C.byref(out), 0))
        if packed_bits is not None:
            batch.bits = unpack_bits(packed_bits, Q)
        for i, ec in enumerate(self.ec):
            cb = batch.complexes[ec]
            if packed_par[i] is not None:
                cb.parity = unpack_bits(packed_par[i], self.lat.complexes[ec].S)
                K = cb.odd_count.astype(np.int64)
                cb.odd_off = np.concatenate([[0], np.cumsum(K)]).astype(np.int64)
                cb.pair_off = np.concatenate([[0], np.cumsum(K * (K - 1) // 2)]).astype(np.int64)
        batch.timings = self.timings()
        return batch

    def fetch_into(self, bufs: dict) -> _lib.Sizes:
        """Copy the last run's results into caller-owned (e.g. pinned) NumPy
        buffers without allocating: keys ``hom``, ``weights``, ``bits`` (packed
        uint32) and per complex index ``i``: ``parity{i}`` (packed), ``odd_count{i}``,
        ``odd_ids{i}``, ``pair_len{i}``, ``bnd_len{i}``, ``bnd_side{i}``.  Buffers must be
        at least as large as the sizes returned by ``sizes()``."""
        s = self.sizes()
        T = int(s.n_trials)
        out = _lib.Outputs()

        def need(key, n, typ):
            a = bufs.get(key)
            if a is None:
                return None
            if a.size < n:
                raise ValueError(f"buffer {key!r} too small: {a.size} < {n}")
            return _ptr(a, typ)

        out.hom = need("hom", T * self.Qtot, _lib._f64p)
        out.weights = need("weights", T * self.Qtot, _lib._f64p)
        out.bits = need("bits", T * self.bit_words, _lib._u32p)
        for i, ec in enumerate(self.ec):
            ct = self.lat.complexes[ec]
            out.parity[i] = need(f"parity{i}", T * ((ct.S + 31) // 32), _lib._u32p)
            out.odd_count[i] = need(f"odd_count{i}", T, _lib._i32p)
            out.odd_ids[i] = need(f"odd_ids{i}", int(s.total_odd[i]), _lib._i32p)
            out.pair_len[i] = need(f"pair_len{i}", int(s.total_pairs[i]), _lib._f64p)
            if ct.has_boundary:
                out.bnd_len[i] = need(f"bnd_len{i}", int(s.total_odd[i]), _lib._f64p)
                out.bnd_side[i] = need(f"bnd_side{i}", int(s.total_odd[i]), _lib._u8p)
        _lib.check(self._lib.fpb_fetch(self._h, C.byref(out), 0))
        return s

    def fetch_decode(self, bufs: Optional[dict] = None) -> "TrialBatch":
        """What the host decoder needs of the last run - bits, odd counts, pair and boundary
        lengths - copied into reusable pinned buffers (``bufs``, grown on demand and owned by
        the caller; the returned arrays are views into them, valid until the next call)."""
        import torch

        bufs = {} if bufs is None else bufs
        s = self.sizes()
        T = int(s.n_trials)

        def room(key, n, dtype):
            a = bufs.get(key)
            if a is None or a.size < n:
                cap = max(int(n * 1.3) + 1024, 1)
                t = torch.empty(cap, dtype=dtype, pin_memory=torch.cuda.is_available())
                a = bufs[key] = t.numpy().view(np.uint32) if key == "bits" else t.numpy()
                bufs["_keep_" + key] = t
            return a

        room("bits", T * self.bit_words, torch.int32)
        for i, ec in enumerate(self.ec):
            room(f"odd_count{i}", T, torch.int32)
            room(f"pair_len{i}", int(s.total_pairs[i]), torch.float64)
            if self.lat.complexes[ec].has_boundary:
                room(f"bnd_len{i}", int(s.total_odd[i]), torch.float64)
        self.fetch_into({k: v for k, v in bufs.items() if not k.startswith("_keep_")})
        batch = TrialBatch(T, _lib.STAGE_GRAPH)
        batch.bits = unpack_bits(bufs["bits"][: T * self.bit_words].reshape(T, self.bit_words), self.Qtot)
        batch.complexes = {}
        for i, ec in enumerate(self.ec):
            cb = ComplexBatch()
            cb.odd_count = bufs[f"odd_count{i}"][:T]
            k64 = cb.odd_count.astype(np.int64)
            cb.odd_off = np.concatenate([[0], np.cumsum(k64)])
            cb.pair_off = np.concatenate([[0], np.cumsum(k64 * (k64 - 1) // 2)])
            cb.pair_len = bufs[f"pair_len{i}"][: int(s.total_pairs[i])]
            cb.bnd_len = (bufs[f"bnd_len{i}"][: int(s.total_odd[i])] if self.lat.complexes[ec].has_boundary
                          else np.empty(0, np.float64))
            batch.complexes[ec] = cb
        batch.timings = self.timings()
        return batch

    def knn(self, complex_index: int, k: int = 12, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Candidate partners for the host matcher from the last graph-stage run: int32[sum K, k],
        for every odd cube of every trial its ``k`` nearest other odd cubes of that trial (indices
        into the trial's odd list, nearest first, -1 padded), computed on the device."""
        s = self.sizes()
        n = int(s.total_odd[complex_index])
        if out is None or out.size < n * k:
            out = np.empty(max(n * k, 1), np.int32)
        if n:
            _lib.check(self._lib.fpb_knn(self._h, complex_index, k, out.ctypes.data, 0))
        return out[: n * k].reshape(n, k)

    def paths_toggle(self, trial, cx, src, dst) -> np.ndarray:
        """XOR-toggle the qubits on the shortest paths of the given (trial,
        complex, odd-index src, odd-index dst | -1) queries of the last run.
        Returns uint8[T, Qtot] flips."""
        s = self.sizes()
        T = int(s.n_trials)
        trial = np.ascontiguousarray(trial, np.int32)
        cx = np.ascontiguousarray(cx, np.int32)
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        n = len(trial)
        flips = np.zeros((T, self.bit_words), np.uint32)
        _lib.check(self._lib.fpb_paths_toggle(
            self._h, n, _ptr(trial, _lib._i32p), _ptr(cx, _lib._i32p), _ptr(src, _lib._i32p),
            _ptr(dst, _lib._i32p), _ptr(flips, _lib._u32p)))
        return unpack_bits(flips, self.Qtot)


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    """uint32[T, W] (LSB-first) -> uint8[T, n]."""
    T = words.shape[0]
    b = np.unpackbits(words.view(np.uint8).reshape(T, -1), axis=1, bitorder="little")
    return np.ascontiguousarray(b[:, :n])
