"""Batched trial engine: the Python face of the C ABI (``include/fpb.h``).

One ``TrialEngine`` = one code instance + noise parameters + weight options on
one GPU.  It runs batches of independent Monte-Carlo trials through
noise -> bits -> weights -> syndrome -> matching graph on the device and hands
back NumPy arrays.  The reference-compatible classes in ``codes.py``,
``noise.py``, ``decoder.py`` and ``simulations.py`` are thin layers over this.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import _lib
from .lattice import LatticeTables


def config_dict(lat: LatticeTables, delta: float, p_swap: float, wo: dict, mode: str = "compat",
                chain_len: int = 0, device: int = 0, delta_step: float = 0.0, length_mode: int = _lib.LENGTH_FLOAT,
                reserve_sms: int = 0):
    """The fields of ``fpb_config`` (include/fpb.h) by name, from lattice tables and options -
    what ``_lib.open_handle`` (pybind11 ``fpbpy.Handle`` or ctypes) takes.
    Returns (config dict, qubit offsets per complex, Qtot)."""
    complexes = []
    q_off = {}
    off = 0
    for ec in lat.ec:
        ct = lat.complexes[ec]
        complexes.append({
            "n_cubes": ct.S, "n_qubits": ct.Q, "n_slots": ct.n_slots,
            "qubit_node": np.ascontiguousarray(ct.qubits, np.int32),
            "cube_qubit": np.ascontiguousarray(ct.cube_qubit, np.int32).ravel(),
            "ninfo": np.ascontiguousarray(ct.ninfo, np.uint16),
            "sbase": np.ascontiguousarray(ct.sbase, np.int32),
            "d_off": np.ascontiguousarray(ct.d_off, np.int32).ravel(),
            "s_off": np.ascontiguousarray(ct.s_off, np.int32).ravel(),
            "slot_qubit": np.ascontiguousarray(ct.slot_qubit, np.int32),
            "low_cube": np.ascontiguousarray(ct.low_seed_cube, np.int32),
            "low_slot": np.ascontiguousarray(ct.low_seed_slot, np.int32),
            "high_cube": np.ascontiguousarray(ct.high_seed_cube, np.int32),
            "high_slot": np.ascontiguousarray(ct.high_seed_slot, np.int32),
        })
        q_off[ec] = off
        off += ct.Q
    cfg = {
        "device": int(device), "n_nodes": lat.N,
        "adj_indptr": np.ascontiguousarray(lat.adj_indptr, np.int32),
        "adj_indices": np.ascontiguousarray(lat.adj_indices, np.int32),
        "adj_vals": np.ascontiguousarray(lat.adj_vals, np.int8),
        "complexes": complexes,
        "delta": float(delta), "p_swap": float(p_swap or 0.0),
        "weight_method": _lib.WEIGHT_BLUEPRINT if wo["method"] == "blueprint" else _lib.WEIGHT_UNIFORM,
        "weight_delta": float(wo.get("delta") or delta),
        "weight_integer": int(bool(wo.get("integer", False))),
        "weight_multiplier": float(wo.get("multiplier", 1)),
        "label_mode": _lib.MODE_FRESH if mode == "fresh" else _lib.MODE_COMPAT,
        "chain_len": int(chain_len), "delta_step": float(delta_step), "length_mode": int(length_mode),
        "reserve_sms": int(reserve_sms),
    }
    return cfg, q_off, off


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
        lengths: str = "float",
        reserve_sms: int = 0,
    ):
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
        if lengths not in ("float", "rustworkx"):
            raise ValueError("lengths must be 'float' (networkx path sums) or 'rustworkx' (sum of int(weight))")
        self.lengths = lengths
        cfg, self.q_off, self.Qtot = config_dict(
            lat, self.delta, self.p_swap, wo, mode, chain_len, device, delta_step,
            _lib.LENGTH_RX_TRUNC if lengths == "rustworkx" else _lib.LENGTH_FLOAT, reserve_sms)
        self.bit_words = (self.Qtot + 31) // 32
        self._h = _lib.open_handle(cfg)  # pybind11 fpbpy.Handle, or ctypes when the module is absent
        self.binding = _lib.binding_name()
        self.uses_table = False
        self.run_id = 0  # bumped by every run: consumers of "the last run" check it

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "_h", None) is not None:
            self._h.close()
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
        self.run_id += 1
        self._h.run_injected(x, state, stage)
        return self.fetch(stage) if fetch else TrialBatch(T, stage, timings=self.timings())

    def run_rng(self, seed: int, first_trial: int, n_trials: int, stage: int = _lib.STAGE_GRAPH,
                fetch: bool = True, want_draws: bool = False) -> TrialBatch:
        """Run trials ``[first_trial, first_trial + n_trials)`` of the Philox stream ``seed``."""
        self.run_id += 1
        self._h.run_rng(seed & (2**64 - 1), first_trial, n_trials, stage)
        return self.fetch(stage, want_draws=want_draws) if fetch else TrialBatch(
            n_trials, stage, timings=self.timings())

    def _packed_bits(self, bits) -> np.ndarray:
        """uint8[T, Qtot] (0/1) or already packed uint32[T, bit_words] -> packed, C-contiguous."""
        bits = np.asarray(bits)
        if bits.dtype != np.uint32:
            if bits.ndim != 2 or bits.shape[1] != self.Qtot:
                raise ValueError(f"bits must have shape (n_trials, {self.Qtot})")
            pad = np.zeros((bits.shape[0], self.bit_words * 32), np.uint8)
            pad[:, : self.Qtot] = bits != 0
            bits = np.packbits(pad, axis=1, bitorder="little").view(np.uint32)
        if bits.ndim != 2 or bits.shape[1] != self.bit_words:
            raise ValueError(f"packed bits must have shape (n_trials, {self.bit_words})")
        return np.ascontiguousarray(bits)

    def run_bits(self, bits, stage: int = _lib.STAGE_GRAPH, fetch: bool = True) -> TrialBatch:
        """Trials given directly as bit values of the syndrome qubits - the i.i.d. Pauli-noise model
        (``flamingpy/noise/iid.py``) or any other bit-level sampler.  ``bits``: uint8[T, Qtot] (0/1)
        or already packed uint32[T, bit_words] (LSB first).  Needs uniform weights."""
        bits = self._packed_bits(bits)
        n_trials = bits.shape[0]
        self.run_id += 1
        self._h.run_bits(bits, stage)
        return self.fetch(stage) if fetch else TrialBatch(n_trials, stage, timings=self.timings())

    def run_weighted(self, weights, bits, stage: int = _lib.STAGE_GRAPH, fetch: bool = True) -> TrialBatch:
        """Trials given as the node attributes the reference's matching-graph builder reads: the
        ``weight`` (float64[T, Qtot]) and ``bit_val`` (uint8[T, Qtot], or packed uint32[T, bit_words])
        of every syndrome qubit, whatever produced them (``mwpm/matching.py:97-173``)."""
        weights = np.ascontiguousarray(weights, np.float64)
        if weights.ndim != 2 or weights.shape[1] != self.Qtot:
            raise ValueError(f"weights must have shape (n_trials, {self.Qtot})")
        bits = self._packed_bits(bits)
        if bits.shape[0] != weights.shape[0]:
            raise ValueError("weights and bits differ in the number of trials")
        self.run_id += 1
        self._h.run_weighted(weights, bits, stage)
        return self.fetch(stage) if fetch else TrialBatch(weights.shape[0], stage, timings=self.timings())

    # -- macronode layer (CVMacroLayer) ------------------------------------------------------------
    def _ensure_macro(self):
        if not getattr(self, "_macro_ready", False):
            from .lattice import macro_tables

            partner, sign = macro_tables(self.lat)
            self._h.set_macro(np.ascontiguousarray(partner, np.int32), np.ascontiguousarray(sign, np.int8))
            self._macro_ready = True

    def run_macro_injected(self, state, mean, z, stage: int = _lib.STAGE_GRAPH) -> dict:
        """``CVMacroLayer`` trials on injected draws: ``state`` uint8[T, 4N] micronode labels, ``mean``
        float32[T, 4N] initial q means, ``z`` float64[T, 4N] standard normals (stars, then planets) -
        see ``noise.MacroSampler``.  Returns the reduced lattice (``fetch_macro``); the usual outputs
        (weights from the precomputed probabilities, bits, parities, matching graph) follow from
        ``fetch``."""
        self._ensure_macro()
        state = np.ascontiguousarray(state, np.uint8)
        mean = np.ascontiguousarray(mean, np.float32)
        z = np.ascontiguousarray(z, np.float64)
        M = 4 * self.lat.N
        if state.ndim != 2 or state.shape[1] != M or mean.shape != state.shape or z.shape != state.shape:
            raise ValueError(f"state, mean and z must have shape [T, {M}]")
        self.run_id += 1
        self._h.run_macro_injected(state, mean, z, stage)
        return self.fetch_macro()

    def run_macro_rng(self, seed: int, first_trial: int, n_trials: int, stage: int = _lib.STAGE_GRAPH,
                      fetch: bool = True):
        """``CVMacroLayer`` trials ``[first_trial, first_trial + n_trials)`` of the Philox stream ``seed``."""
        self._ensure_macro()
        self.run_id += 1
        self._h.run_macro_rng(seed & (2**64 - 1), first_trial, n_trials, stage)
        return self.fetch(stage) if fetch else TrialBatch(n_trials, stage, timings=self.timings())

    def fetch_syndrome(self, want_state: bool = False, packed_bits: bool = False):
        """What the Union-Find decoder consumes from the last run (stage >= SYNDROME): the bit values
        uint8[T, Qtot] (``packed_bits``: the device's uint32[T, ceil(Qtot/32)] words, LSB first, which
        ``decoder.logical_failures`` reads directly), per complex the cube parities uint8[T, S], and - on
        request - the state labels uint8[T, N] (1 GKP, 2 p) that decide which edges are erased."""
        s = self.sizes()
        T = int(s.n_trials)
        bufs = {"bits": np.empty((T, self.bit_words), np.uint32)}
        if want_state:
            bufs["state"] = np.empty((T, self.lat.N), np.uint8)
        for i, ec in enumerate(self.ec):
            bufs[f"parity{i}"] = np.empty((T, (self.lat.complexes[ec].S + 31) // 32), np.uint32)
        self._h.fetch_into(bufs)
        parity = [unpack_bits(bufs[f"parity{i}"], self.lat.complexes[ec].S) for i, ec in enumerate(self.ec)]
        return (bufs["bits"] if packed_bits else unpack_bits(bufs["bits"], self.Qtot)), parity, bufs.get("state")

    def logical_check(self, want_trials: bool = True):
        """``check_correction`` of every trial of the last run on the device (``fpb_logical_check``): bit values
        XOR the recovery the last ``uf_decode`` / ``paths_toggle`` left in HBM, over the lattice's correlation
        surfaces.  Returns bool[T] (True = logical failure), or only their number with ``want_trials=False``."""
        if not getattr(self, "_planes_set", False):
            off, q, qoff = [0], [], 0
            for ec in self.ec:
                ct = self.lat.complexes[ec]
                for plane in ct.planes:
                    q.append(np.asarray(plane, np.int64) + qoff)
                    off.append(off[-1] + len(plane))
                qoff += ct.Q
            _lib.set_planes(self._h, np.asarray(off, np.int32),
                            np.ascontiguousarray(np.concatenate(q) if q else np.zeros(0), np.int32))
            self._planes_set = True
        if not want_trials:
            return _lib.logical_check(self._h)
        fail = np.zeros(int(self.sizes().n_trials), np.uint8)
        _lib.logical_check(self._h, fail)
        return fail.astype(bool)

    def uf_decode(self, use_erasures: bool = True, want_grown: bool = False, fetch: bool = True):
        """Union-Find + peeling recovery of the last run (stage >= SYNDROME) computed on the device
        (``fpb_uf_decode``, ``csrc/fpb_uf.cuh``): bit flips uint8[T, Qtot].  ``use_erasures``: start from the
        erased edges the state labels of the run define (two or more p-squeezed neighbours); runs without
        labels (bit-level, i.i.d.) have none.  ``want_grown``: also the fully grown edges uint8[T, Qtot] (the
        reference's Support table when the growth stops).  ``fetch=False``: the recovery stays on the device
        (for ``logical_check``) and None is returned.  Raises ``ValueError`` like the host decoder when an odd
        cluster can neither grow nor reach a boundary point."""
        from . import uf_native

        if not getattr(self, "_uf_graph_set", False):
            for i, ec in enumerate(self.ec):
                n_nodes, ea, eb = uf_native.uf_graph(self.lat.complexes[ec])
                _lib.uf_set_graph(self._h, i, n_nodes, np.ascontiguousarray(ea, np.int32), np.ascontiguousarray(eb, np.int32))
            self._uf_graph_set = True
        T = int(self.sizes().n_trials)
        flips = np.zeros((T, self.bit_words), np.uint32) if fetch or want_grown else None
        grown = np.zeros((T, self.bit_words), np.uint32) if want_grown else None
        stuck = _lib.uf_decode(self._h, use_erasures, flips, grown)
        if stuck:
            raise ValueError(f"{stuck} (trial, complex) item(s): an odd cluster can neither grow nor reach a boundary point "
                             "(odd number of unsatisfied stabilizers on a complex without boundary)")
        if want_grown:
            return unpack_bits(flips, self.Qtot), unpack_bits(grown, self.Qtot)
        return unpack_bits(flips, self.Qtot) if fetch else None

    def fetch_macro(self) -> dict:
        """Reduced lattice of the last macronode run, each [T, N]: ``red_state`` (1 GKP, 2 p),
        ``bit_val``, ``p_phase_cond``, ``outcome`` (effective homodyne value)."""
        rs, bv, pp, oc = self._h.fetch_macro()
        return {"red_state": rs, "bit_val": bv, "p_phase_cond": pp, "outcome": oc}

    def run_iid(self, seed: int, first_trial: int, n_trials: int, p_err: float, stage: int = _lib.STAGE_GRAPH,
                fetch: bool = True) -> TrialBatch:
        """Trials ``[first_trial, first_trial + n_trials)`` of i.i.d. Z errors with probability
        ``p_err`` per qubit, drawn on the device (Philox stream ``seed``).  Needs uniform weights."""
        self.run_id += 1
        self._h.run_iid(seed & (2**64 - 1), first_trial, n_trials, float(p_err), stage)
        return self.fetch(stage) if fetch else TrialBatch(n_trials, stage, timings=self.timings())

    def sizes(self) -> _lib.SizesView:
        return _lib.SizesView(self._h.sizes())

    def timings(self) -> dict:
        return dict(self._h.timings())

    def synchronize(self):
        """Wait for everything queued on the handle's stream."""
        self._h.synchronize()

    def build_table(self):
        """Precompute the all-pairs distance table for trial-independent weights (uniform, or
        blueprint with ``p_swap == 1``); later graph-stage runs gather from it."""
        self._h.build_table()
        self.uses_table = True

    def mark(self, which: int):
        """Record timing mark 0 / 1 on the handle's stream."""
        self._h.mark(which)

    def elapsed_ms(self) -> float:
        return float(self._h.elapsed_ms())

    def device_outputs(self) -> dict:
        """Raw device addresses of the last run's result buffers (``fpb_device_outputs``), keyed
        like ``fetch_into``'s buffers."""
        return dict(self._h.device_outputs())

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

        def view(addr, shape, typestr):
            if not addr or 0 in shape:
                return None
            return torch.as_tensor(_View(addr, shape, typestr), device=torch.device("cuda", self.device))

        out = {"hom": view(o["hom"], (T, self.Qtot), "<f8"), "weights": view(o["weights"], (T, self.Qtot), "<f8"),
               "bits": view(o["bits"], (T, self.bit_words), "<i4")}
        for i, ec in enumerate(self.ec):
            ct = self.lat.complexes[ec]
            out[ec] = {
                "parity": view(o[f"parity{i}"], (T, (ct.S + 31) // 32), "<i4"),
                "odd_count": view(o[f"odd_count{i}"], (T,), "<i4"),
                "odd_ids": view(o[f"odd_ids{i}"], (int(s.total_odd[i]),), "<i4"),
                "pair_len": view(o[f"pair_len{i}"], (int(s.total_pairs[i]),), "<f8"),
                "bnd_len": view(o[f"bnd_len{i}"], (int(s.total_odd[i]),), "<f8"),
            }
        return out

    def fetch(self, stage: int = _lib.STAGE_GRAPH, want_draws: bool = False, want_noise: bool = True) -> TrialBatch:
        """Copy the last run's results to host arrays."""
        s = self.sizes()
        T = int(s.n_trials)
        N, Q = self.lat.N, self.Qtot
        bufs = {}
        batch = TrialBatch(T, stage)
        if want_noise:
            batch.hom = bufs["hom"] = np.empty((T, Q), np.float64)
            batch.weights = bufs["weights"] = np.empty((T, Q), np.float64)
            bufs["bits"] = np.empty((T, self.bit_words), np.uint32)
        if want_draws:
            batch.state = bufs["state"] = np.empty((T, N), np.uint8)
            batch.x = bufs["x"] = np.empty((T, 2 * N), np.float64)
        batch.complexes = {}
        for i, ec in enumerate(self.ec):
            ct = self.lat.complexes[ec]
            cb = ComplexBatch()
            if stage >= _lib.STAGE_SYNDROME:
                bufs[f"parity{i}"] = np.empty((T, (ct.S + 31) // 32), np.uint32)
                cb.odd_count = bufs[f"odd_count{i}"] = np.empty(T, np.int32)
                cb.odd_ids = bufs[f"odd_ids{i}"] = np.empty(int(s.total_odd[i]), np.int32)
            if stage >= _lib.STAGE_GRAPH:
                cb.pair_len = bufs[f"pair_len{i}"] = np.empty(int(s.total_pairs[i]), np.float64)
                if ct.has_boundary:
                    cb.bnd_len = bufs[f"bnd_len{i}"] = np.empty(int(s.total_odd[i]), np.float64)
                    cb.bnd_side = bufs[f"bnd_side{i}"] = np.empty(int(s.total_odd[i]), np.uint8)
                else:
                    cb.bnd_len = np.empty(0, np.float64)
                    cb.bnd_side = np.empty(0, np.uint8)
            batch.complexes[ec] = cb
        self._h.fetch_into(bufs)
        if want_noise:
            batch.bits = unpack_bits(bufs["bits"], Q)
        for i, ec in enumerate(self.ec):
            cb = batch.complexes[ec]
            if stage >= _lib.STAGE_SYNDROME:
                cb.parity = unpack_bits(bufs[f"parity{i}"], self.lat.complexes[ec].S)
                K = cb.odd_count.astype(np.int64)
                cb.odd_off = np.concatenate([[0], np.cumsum(K)]).astype(np.int64)
                cb.pair_off = np.concatenate([[0], np.cumsum(K * (K - 1) // 2)]).astype(np.int64)
        batch.timings = self.timings()
        return batch

    def fetch_into(self, bufs: dict) -> _lib.SizesView:
        """Copy the last run's results into caller-owned (e.g. pinned) NumPy
        buffers without allocating: keys ``hom``, ``weights``, ``bits`` (packed
        uint32) and per complex index ``i``: ``parity{i}`` (packed), ``odd_count{i}``,
        ``odd_ids{i}``, ``pair_len{i}``, ``bnd_len{i}``, ``bnd_side{i}``.  Buffers must be
        at least as large as the sizes returned by ``sizes()``."""
        s = self.sizes()
        self._h.fetch_into({k: v for k, v in bufs.items() if isinstance(v, np.ndarray)})
        return s

    def fetch_decode(self, bufs: Optional[dict] = None, want_pairs: bool = True, packed_bits: bool = False) -> "TrialBatch":
        """What the host decoder needs of the last run - bits, odd counts, pair and boundary
        lengths - copied into reusable pinned buffers (``bufs``, grown on demand and owned by
        the caller; the returned arrays are views into them, valid until the next call).
        ``want_pairs=False`` (sparse hand-off) leaves the K(K-1)/2 pair lengths on the device.
        ``packed_bits``: ``batch.bits`` stays the device's uint32[T, ceil(Qtot/32)] words (a view;
        ``decoder.logical_failures`` reads them as they are) instead of being unpacked to bytes."""
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
            if want_pairs:
                room(f"pair_len{i}", int(s.total_pairs[i]), torch.float64)
            if self.lat.complexes[ec].has_boundary:
                room(f"bnd_len{i}", int(s.total_odd[i]), torch.float64)
        self.fetch_into({k: v for k, v in bufs.items()
                         if not k.startswith("_keep_") and (want_pairs or not k.startswith("pair_len"))})
        batch = TrialBatch(T, _lib.STAGE_GRAPH)
        words = bufs["bits"][: T * self.bit_words].reshape(T, self.bit_words)
        batch.bits = words if packed_bits else unpack_bits(words, self.Qtot)
        batch.complexes = {}
        for i, ec in enumerate(self.ec):
            cb = ComplexBatch()
            cb.odd_count = bufs[f"odd_count{i}"][:T]
            k64 = cb.odd_count.astype(np.int64)
            cb.odd_off = np.concatenate([[0], np.cumsum(k64)])
            cb.pair_off = np.concatenate([[0], np.cumsum(k64 * (k64 - 1) // 2)])
            cb.pair_len = bufs[f"pair_len{i}"][: int(s.total_pairs[i])] if want_pairs else None
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
            self._h.knn(complex_index, k, out)
        return out[: n * k].reshape(n, k)

    # -- sparse hand-off to the host matcher (fpb_knn_len / fpb_pairs_at / fpb_price) ----------------
    def knn_len(self, complex_index: int, k: int = 12):
        """Candidate partners with their reduced weights ``min(d_ab, b_a + b_b)``: (ids int32[sum K, k],
        lens float64[sum K, k], wmax float64[T]).  The K(K-1)/2 lengths stay on the device."""
        return self._h.knn_len(int(complex_index), int(k))

    def knn_len_into(self, complex_index: int, k: int, ids: np.ndarray, lens: np.ndarray, wmax: np.ndarray):
        """``knn_len`` into caller-owned (e.g. pinned) buffers: no allocation, direct D2H copies."""
        self._h.knn_len_into(int(complex_index), int(k), ids, lens, wmax)

    def knn_len_pinned(self, complex_index: int, k: int, bufs: dict):
        """``knn_len`` into reusable pinned buffers kept in ``bufs`` (grown on demand, owned by the caller): the
        returned arrays are views, valid until the next call with the same ``bufs``.  Saves a fresh 20 MB
        allocation and a pageable device-to-host copy per d = 13 batch."""
        import torch

        s = self.sizes()
        n, T = int(s.total_odd[complex_index]), int(s.n_trials)

        def room(key, count, dtype):
            a = bufs.get(key)
            if a is None or a.size < count:
                t = torch.empty(max(int(count * 1.3) + 1024, 1), dtype=dtype, pin_memory=torch.cuda.is_available())
                a = bufs[key] = t.numpy()
                bufs["_keep_" + key] = t
            return a

        ids = room(f"ids{complex_index}", n * k, torch.int32)
        lens = room(f"lens{complex_index}", n * k, torch.float64)
        wmax = room(f"wmax{complex_index}", T, torch.float64)
        self._h.knn_len_into(int(complex_index), int(k), ids, lens, wmax)
        return ids[: n * k].reshape(n, k), lens[: n * k].reshape(n, k), wmax[:T]

    def pairs_at(self, complex_index: int, trial, a, b) -> np.ndarray:
        """Shortest-path lengths of the requested pairs (trial, odd index a, odd index b) of the last run."""
        c = lambda x: np.ascontiguousarray(x, np.int32)  # noqa: E731
        return self._h.pairs_at(int(complex_index), c(trial), c(a), c(b))

    def pairs_of_trial(self, complex_index: int, trial: int, K: int) -> np.ndarray:
        """The packed triangle of all K(K-1)/2 lengths of one trial of the last run."""
        return self._h.pairs_of_trial(int(complex_index), int(trial), int(K) * (int(K) - 1) // 2)

    def price(self, complex_index: int, y, scale, wshift, active, cap: int = 65536, top=None, ztop=None):
        """Pricing of the complete graph under the matcher's vertex duals on the device, in the
        matcher's integer arithmetic: the pairs with ``y_a + y_b + ((4 * int64(w_ab * scale_t + 0.5)) <<
        wshift_t) < 0`` as (trial, a, b, w_ab) arrays.  ``top`` / ``ztop``: each vertex's outermost
        blossom (-1: none) and twice its dual, added for two vertices of one blossom."""
        if top is not None:
            top, ztop = np.ascontiguousarray(top, np.int32), np.ascontiguousarray(ztop, np.int64)
        return self._h.price(int(complex_index), np.ascontiguousarray(y, np.int64), top, ztop,
                             np.ascontiguousarray(scale, np.float64), np.ascontiguousarray(wshift, np.int32),
                             np.ascontiguousarray(active, np.uint8), int(cap))

    def paths_list(self, trial, cx, src, dst, max_len: int = 0):
        """Ordered correction paths of the given queries (same arguments as ``paths_toggle``):
        a list with one int32 array per query holding the local qubit ids crossed (qubit order of the
        query's complex), walking from ``dst`` back to ``src``, or from ``src`` out to its nearer
        connector (``dst == -1``, last entry = the boundary qubit)."""
        trial = np.ascontiguousarray(trial, np.int32)
        cx = np.ascontiguousarray(cx, np.int32)
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        if max_len <= 0:
            max_len = max(self.lat.complexes[e].S for e in self.ec) + 1  # a simple path visits a cube once
        qubits, length = self._h.paths_list(trial, cx, src, dst, int(max_len))
        if np.any(length < 0):
            raise ValueError("a path is longer than max_len")
        return [qubits[i, : length[i]].copy() for i in range(len(trial))]

    def paths_toggle(self, trial, cx, src, dst, packed: bool = False) -> np.ndarray:
        """XOR-toggle the qubits on the shortest paths of the given (trial,
        complex, odd-index src, odd-index dst | -1) queries of the last run.
        Returns uint8[T, Qtot] flips (``packed``: the uint32[T, ceil(Qtot/32)] words as copied)."""
        s = self.sizes()
        T = int(s.n_trials)
        trial = np.ascontiguousarray(trial, np.int32)
        cx = np.ascontiguousarray(cx, np.int32)
        src = np.ascontiguousarray(src, np.int32)
        dst = np.ascontiguousarray(dst, np.int32)
        flips = np.zeros((T, self.bit_words), np.uint32)
        self._h.paths_toggle(trial, cx, src, dst, flips)
        return flips if packed else unpack_bits(flips, self.Qtot)


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    """uint32[T, W] (LSB-first) -> uint8[T, n]."""
    T = words.shape[0]
    b = np.unpackbits(words.view(np.uint8).reshape(T, -1), axis=1, bitorder="little")
    return np.ascontiguousarray(b[:, :n])
