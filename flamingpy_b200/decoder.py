"""Reference-compatible decoding front end (``flamingpy/decoders/decoder.py``).

``correct(code, decoder="MWPM", weight_options=..., decoder_opts=...)`` keeps the
reference's signature and result.  Weight assignment, syndrome evaluation and
matching-graph construction run on the GPU for the noise realisation the last
``CVLayer.apply_noise`` call produced; the minimum-weight perfect matching is
solved on the host (north star: "the final MWPM solve still hands the GPU-built
graph to the reference's own matcher"), then the matched pairs' correction
paths are extracted on the GPU and the logical check is evaluated.
"""

from __future__ import annotations

import numpy as np

from . import _lib
from .matching import GpuMatchingGraph, solve_matching


def _require_trial(code):
    if code._trial is None:
        raise RuntimeError("no noise realisation: call the noise model's apply_noise(rng) first")
    return code._trial


def _run_trial(code, wo, stage, lengths="float"):
    """Run the last noise realisation (one trial) through the device pipeline up to ``stage``.
    CVLayer realisations are (x, state, delta, p_swap); bit-level ones (IidNoise) are
    ("bits", bit values of all nodes) and need uniform weights."""
    trial = _require_trial(code)
    if isinstance(trial[0], str) and trial[0] == "macro":
        # CVMacroLayer realisation: (labels, initial means, normals) of the micronodes; blueprint weights
        # come from the reduced nodes' p_phase_cond (decoder.py:67-68)
        if wo.get("method") == "blueprint" and not wo.get("prob_precomputed"):
            raise ValueError("CVMacroLayer leaves no homodyne values on the reduced lattice: use weight_options "
                             "{'method': 'blueprint', 'prob_precomputed': True} (or 'uniform')")
        (state, mean, z), delta, p_swap = trial[1], trial[2], trial[3]
        eng = code.engine(delta, p_swap, wo, lengths=lengths)
        eng.run_macro_injected(state[None, :], mean[None, :], z[None, :], stage=stage)
        return eng, eng.fetch(stage)
    if wo.get("prob_precomputed"):
        raise ValueError("prob_precomputed weights need the p_phase_cond attribute that CVMacroLayer computes")
    if isinstance(trial[0], str) and trial[0] == "bits":
        if wo.get("method", "uniform") != "uniform":
            raise ValueError("IidNoise has no homodyne outcomes: use weight_options {'method': 'uniform'}")
        eng = code.engine(1.0, 0.0, wo, lengths=lengths)
        qn = np.concatenate([code.lattice.complexes[e].qubits for e in code.ec])
        return eng, eng.run_bits(trial[1][qn][None, :], stage=stage)
    x, state, delta, p_swap = trial
    eng = code.engine(delta, p_swap, wo, lengths=lengths)
    return eng, eng.run_injected(x[None, :], state[None, :], stage=stage)


def _trial_state(code, eng=None):
    """State labels (uint8[N], 2 = p-squeezed) of the last noise realisation, or None when the noise
    model has none (IidNoise).  CVMacroLayer: the labels of the reduced lattice, computed by the
    device pipeline (``eng`` must have run the trial)."""
    trial = _require_trial(code)
    if isinstance(trial[0], str):
        if trial[0] == "macro":
            if eng is None:
                eng, _ = _run_trial(code, {"method": "uniform"}, _lib.STAGE_NOISE)
            return eng.fetch_macro()["red_state"][0]
        return None
    return trial[1]


def uf_erasures(code, state):
    """Erased edges of the Union-Find decoder: syndrome qubits with two or more p-squeezed
    neighbours (weight -1 of ``assign_weights(code, "UF")``, ``decoders/decoder.py:96-114``).
    ``state`` uint8[T, N] (2 = p) -> uint8[T, Qtot]; None (no labels) -> None."""
    if state is None:
        return None
    from . import uf_native

    qn = code.__dict__.get("_uf_qubit_nodes")
    if qn is None:
        qn = code.__dict__["_uf_qubit_nodes"] = np.ascontiguousarray(
            np.concatenate([code.lattice.complexes[e].qubits for e in code.ec]), np.int32)
    return uf_native.erasures(code.lattice, qn, state)


def uf_decode_batch(code, parity, erased=None, threads=0):
    """Union-Find + peeling recovery of a batch: ``parity`` = per complex uint8[T, S], ``erased``
    uint8[T, Qtot] or None -> bit flips uint8[T, Qtot] (host decoder, ``csrc/uf.cpp``)."""
    from . import uf_native

    out, off = [], 0
    for i, e in enumerate(code.ec):
        ct = code.lattice.complexes[e]
        er = None if erased is None else erased[:, off : off + ct.Q]
        out.append(uf_native.decode_batch(ct, parity[i], er, threads=threads))
        off += ct.Q
    return np.concatenate(out, axis=1)


def uf_decoder(code, ec, draw=False, drawing_opts=None):
    """``uf_decoder`` (``decoders/unionfind/algos.py:314-334``): Union-Find + peeling on the code's
    current ``bit_val`` / ``weight`` attributes (weight -1 = erased edge); returns the set of qubits
    (coordinates) whose bit values the recovery flips."""
    if draw:
        raise NotImplementedError("drawing is out of scope for the accelerated path")
    from . import uf_native

    g = code.graph
    g._ensure_arrays()
    ct = code.lattice.complexes[ec]
    bits = g.bit_val[ct.qubits].astype(np.uint8)
    parity = np.zeros(ct.S, np.uint8)
    for f in range(6):
        q = ct.cube_qubit[:, f]
        parity[q >= 0] ^= bits[q[q >= 0]]
    erased = (g.weight[ct.qubits] == -1).astype(np.uint8)
    flips = uf_native.decode_batch(ct, parity[None, :], erased[None, :] if erased.any() else None)[0]
    return {tuple(int(v) for v in code.lattice.coords[ct.qubits[q]]) for q in np.nonzero(flips)[0]}


def _assign_weights_uf(code, wo):
    """The ``decoder == "UF"`` branch of ``assign_weights`` (``decoders/decoder.py:94-114``)."""
    if wo.get("method") == "blueprint":
        raise Exception("Incompatible decoder & weight options combination.")  # noqa: TRY002 - the reference's
    if wo.get("method") != "uniform":
        raise ValueError(f"unknown weight method {wo.get('method')!r}")
    er = uf_erasures(code, _trial_state(code))
    g = code.graph
    g._ensure_arrays()
    off = 0
    for e in code.ec:
        ct = code.lattice.complexes[e]
        w = np.full(ct.Q, 2.0)
        if er is not None:
            w[er[0, off : off + ct.Q] == 1] = -1.0
        g.weight[ct.qubits] = w
        off += ct.Q


def assign_weights(code, decoder="MWPM", **kwargs):
    """``assign_weights`` (``decoders/decoder.py:31-119``): writes the ``weight`` attribute of every
    syndrome qubit - for the MWPM decoder on the device (uniform / blueprint), for the Union-Find
    decoder (uniform only: 2 per edge, -1 for an erased one) on the host from the state labels."""
    wo = {"method": "uniform", "integer": False, "multiplier": 1}
    wo.update(kwargs)
    if decoder == "UF":
        return _assign_weights_uf(code, wo)
    if decoder != "MWPM":
        raise ValueError(f"unknown decoder {decoder!r}")
    eng, batch = _run_trial(code, wo, _lib.STAGE_NOISE)
    g = code.graph
    g._ensure_arrays()
    off = 0
    for e in code.ec:
        ct = code.lattice.complexes[e]
        w = batch.weights[0, off : off + ct.Q]
        g.weight[ct.qubits] = w
        off += ct.Q
    return eng


def build_match_graph(code, ec, matching_backend="native", weight_options=None):
    """``build_match_graph(code, ec, matching_backend)`` (``decoders/mwpm/algos.py:23-47``): the
    matching graph of the code's current ``weight`` / ``bit_val`` attributes, built on the GPU.

    ``matching_backend``: a string picks the host matcher of ``GpuMatchingGraph`` ("native";
    "networkx"; the reference's compiled backends "rustworkx" / "lemon" map to the native one), or -
    like in the reference - any class deriving from the ``MatchingGraph`` contract with an
    ``(ec, code)`` constructor, e.g. ``GpuMatchingGraph`` itself.  ``weight_options`` (not in the
    reference's signature) runs ``assign_weights`` first; without it the weights must already be
    on the code, as in the reference's ``correct`` (``decoder.py:270-273``)."""
    if weight_options is not None:
        assign_weights(code, "MWPM", **weight_options)
    if isinstance(matching_backend, str):
        backend = "native" if matching_backend in ("rustworkx", "retworkx", "lemon") else matching_backend
        if backend not in ("native", "networkx"):
            raise ValueError(f"unknown matching backend {matching_backend!r}")
        return GpuMatchingGraph(ec, code, backend=backend)
    return matching_backend(ec, code)


def mwpm_decoder(code, ec, backend="native", draw=False, drawing_opts=None):
    """``mwpm_decoder`` (``decoders/mwpm/algos.py:50-96``): build the matching graph, solve the
    matching, and return the set of qubits (coordinates) whose bit values the recovery flips -
    the XOR of the ``common_vertex`` of every edge on every matched path."""
    if draw:
        raise NotImplementedError("drawing is out of scope for the accelerated path")
    stab_graph = getattr(code, ec + "_stab_graph")
    matching_graph = build_match_graph(code, ec, backend)
    virtual = set(matching_graph.virtual_points)
    qubits_to_flip = set()
    for match in matching_graph.min_weight_perfect_matching():
        if match[0] in virtual and match[1] in virtual:
            continue
        path = matching_graph.edge_path(match)
        for a, b in zip(path[:-1], path[1:]):
            qubits_to_flip ^= {stab_graph.edge_data(a, b)["common_vertex"]}
    return qubits_to_flip


def recovery(qubits_to_flip, code, ec, sanity_check=False):
    """``recovery`` (``decoders/decoder.py:122-148``)."""
    g = code.graph
    for q in qubits_to_flip:
        g.bit_val[g.to_indices[tuple(q)]] ^= 1
    if sanity_check:
        odd = list(getattr(code, ec + "_stab_graph").odd_parity_stabilizers())
        if odd:
            print("Unsatisfied " + ec + " stabilizers:", odd)
        else:
            print(ec.capitalize() + " recovery succeeded - no unsatisfied stabilizers.")


def check_correction(code, sanity_check=False):
    """``check_correction`` (``decoders/decoder.py:151-224``): parity of the corrected
    bits over the first correlation surface of each checked direction."""
    g = code.graph
    ec_checks, truth_dicts = [], []
    for ec in code.ec:
        ct = code.lattice.complexes[ec]
        truth = {"x": [], "y": [], "z": []}
        qcoords = code.lattice.coords[ct.qubits]
        minimum = 0 if ec == "primal" else 1
        for name in ct.plane_names:
            a = "xyz".index(name)
            maximum = 2 * code.dims[a] if sanity_check else minimum + 2
            for sheet in range(minimum, maximum, 2):
                sel = ct.qubits[qcoords[:, a] == sheet]
                parity = int(np.bitwise_xor.reduce(g.bit_val[sel].astype(np.int64))) if len(sel) else 0
                truth[name].append(bool(1 - parity))
        if sanity_check:
            print(ec.capitalize() + " error correction check --", truth)
        ec_checks.append(bool(np.all([truth[n][0] for n in ct.plane_names])))
        truth_dicts.append(truth)
    if sanity_check:
        return ec_checks, truth_dicts
    return ec_checks


def correct(code, decoder="MWPM", weight_options=None, sanity_check=False, decoder_opts=None, draw=False,
            drawing_opts=None):
    """``correct`` (``decoders/decoder.py:227-288``).  Returns True if error correction
    succeeded.  ``decoder_opts["backend"]`` picks the host matcher.  The default is the compiled one
    ("native": exact C++ blossom matcher), like the reference's default is its compiled backend
    (``decoder.py:277``); the reference's "rustworkx" and "lemon" are not installable offline and
    map to it (same total weight, possibly another choice among tied matchings).  "networkx"
    reproduces the reference's networkx verdicts trial by trial, ties included, and is slow.
    ``decoder_opts["lengths"]``: "float" (default) - path lengths are the sums of the float weights, as
    with the reference's networkx graphs; "rustworkx" - the sums of ``int(weight)`` along the
    float-shortest paths, which is what the reference's default rustworkx graphs report
    (``stabilizer_graph.py:486-500``) and therefore what its default ``correct`` matches on.
    With ``decoder="UF"``, ``decoder_opts["backend"]`` is "native" (host decoder, default) or "device"."""
    if decoder not in ("MWPM", "UF"):
        raise KeyError(decoder)  # the reference's decoder_dict lookup (decoder.py:279-286)
    if draw:
        raise NotImplementedError("drawing is out of scope for the accelerated path")
    weight_options = dict(weight_options or {})
    if decoder == "UF":
        return _correct_uf(code, weight_options, sanity_check, (decoder_opts or {}).get("backend", "native"))
    opts = {"backend": "native"}
    opts.update(decoder_opts or {})
    if opts["backend"] in ("rustworkx", "retworkx", "lemon"):
        opts["backend"] = "native"
    wo = {"method": "uniform", "integer": False, "multiplier": 1}
    wo.update(weight_options)
    eng, batch = _run_trial(code, wo, _lib.STAGE_GRAPH, lengths=opts.get("lengths", "float"))
    is_cv = not isinstance(code._trial[0], str)  # CVLayer: homodyne values exist on the lattice
    g = code.graph
    g._ensure_arrays()
    off = 0
    for e in code.ec:
        ct = code.lattice.complexes[e]
        g.weight[ct.qubits] = batch.weights[0, off : off + ct.Q]
        if is_cv:
            g.hom_val_p[ct.qubits] = batch.hom[0, off : off + ct.Q]
        g.bit_val[ct.qubits] = batch.bits[0, off : off + ct.Q]
        off += ct.Q
    flips = decode_batch(code, eng, batch, backend=opts["backend"])
    off = 0
    for e in code.ec:
        ct = code.lattice.complexes[e]
        g.bit_val[ct.qubits] ^= flips[0, off : off + ct.Q].astype(np.int8)
        off += ct.Q
        if sanity_check:
            recovery([], code, e, sanity_check=True)
    result = check_correction(code, sanity_check=sanity_check)
    if sanity_check:
        return bool(np.all(result[0]))
    return bool(np.all(result))


def _correct_uf(code, weight_options, sanity_check, backend="native"):
    """``correct(code, "UF", ...)``: noise -> bits -> syndrome on the device, Union-Find + peeling on
    the host (``decoder_opts["backend"] = "device"``: by ``k_uf`` on the GPU), logical check."""
    if backend not in ("native", "device"):
        raise ValueError(f"unknown Union-Find backend {backend!r}")
    wo = {"method": "uniform", "integer": False, "multiplier": 1}
    wo.update(weight_options)
    if wo.get("method") == "blueprint":
        raise Exception("Incompatible decoder & weight options combination.")  # noqa: TRY002 - the reference's
    eng, batch = _run_trial(code, wo, _lib.STAGE_SYNDROME)
    er = uf_erasures(code, _trial_state(code, eng))
    g = code.graph
    g._ensure_arrays()
    off = 0
    for e in code.ec:
        ct = code.lattice.complexes[e]
        w = np.full(ct.Q, 2.0)
        if er is not None:
            w[er[0, off : off + ct.Q] == 1] = -1.0
        g.weight[ct.qubits] = w
        g.bit_val[ct.qubits] = batch.bits[0, off : off + ct.Q]
        off += ct.Q
    if backend == "device":
        flips = eng.uf_decode(use_erasures=er is not None)
    else:
        flips = uf_decode_batch(code, [batch.complexes[e].parity for e in code.ec], er)
    off = 0
    for e in code.ec:
        ct = code.lattice.complexes[e]
        g.bit_val[ct.qubits] ^= flips[0, off : off + ct.Q].astype(np.int8)
        off += ct.Q
        if sanity_check:
            recovery([], code, e, sanity_check=True)
    result = check_correction(code, sanity_check=sanity_check)
    return bool(np.all(result[0] if sanity_check else result))


def match_batch(code, eng, batch, backend="native", cands=None, handoff="full", stats=None, threads=0):
    """Host matching of every trial of ``batch``: (trial, complex, src, dst) int32 arrays of the
    matched queries (dst = -1: to the connector).  ``backend="native"`` solves all trials with the
    C++ matcher (OpenMP over trials); ``cands`` = per-complex candidate lists (``candidates``).
    ``handoff="sparse"`` (native backend only): the pair lengths stay on the device - the matcher
    gets candidate lists with lengths and prices the complete graph through ``TrialEngine.price``
    (``mwpm_native.solve_batch_sparse``); ``batch`` must be the engine's last run and ``cands`` then
    holds per complex the ``TrialEngine.knn_len`` triple (or None: fetched here)."""
    trial, cx, src, dst = [], [], [], []
    for ci, e in enumerate(code.ec):
        ct = code.lattice.complexes[e]
        cb = batch.complexes[e]
        if backend == "native" and handoff == "sparse":
            from . import mwpm_native

            st = {} if stats is not None else None
            t_, s_, d_ = mwpm_native.solve_batch_sparse(eng, ci, cb.odd_count, cb.bnd_len if ct.has_boundary else None,
                                                        stats=st, knn=cands[ci] if cands is not None else None,
                                                        threads=threads)
            if stats is not None:
                for k, v in st.items():
                    stats[k] = stats.get(k, 0) + v
            trial.append(t_)
            cx.append(np.full(len(t_), ci, np.int32))
            src.append(s_)
            dst.append(d_)
            continue
        if backend == "native":
            from . import mwpm_native

            cand = cands[ci] if cands is not None else candidates(eng, batch, ci, e)
            t_, s_, d_ = mwpm_native.solve_batch(cb.odd_count, cb.pair_len, cb.bnd_len if ct.has_boundary else None,
                                                 threads=threads, cand=cand)
            trial.append(t_)
            cx.append(np.full(len(t_), ci, np.int32))
            src.append(s_)
            dst.append(d_)
            continue
        tt, ss, dd = [], [], []
        for t in range(batch.n_trials):
            K = int(cb.odd_count[t])
            if K == 0:
                continue
            bl = cb.boundary(t)[0] if ct.has_boundary else None
            for u, v in solve_matching(K, cb.pairs(t), bl, backend=backend):
                tt.append(t)
                ss.append(u)
                dd.append(v)
        trial.append(np.array(tt, np.int32))
        cx.append(np.full(len(tt), ci, np.int32))
        src.append(np.array(ss, np.int32))
        dst.append(np.array(dd, np.int32))
    cat = lambda parts: np.concatenate(parts) if parts else np.empty(0, np.int32)  # noqa: E731
    return cat(trial), cat(cx), cat(src), cat(dst)


def candidates(eng, batch, ci, e):
    """Large graphs: the device picks each cube's nearest partners (``TrialEngine.knn``), the host
    solver starts from that sparse graph and still prices every pair, so the matching stays exact.
    Must be called while ``batch`` is the engine's last run; None for batches of small graphs."""
    from . import mwpm_native

    cb = batch.complexes[e]
    if len(cb.odd_count) and int(cb.odd_count.max()) >= mwpm_native.SPARSE_MIN_K and batch.n_trials == eng.sizes().n_trials:
        return eng.knn(ci, mwpm_native.CAND_K)
    return None


def decode_batch(code, eng, batch, backend="native"):
    """Host matching + device path extraction for every trial of ``batch``.
    Returns the recovery as uint8[T, Qtot] bit flips."""
    return eng.paths_toggle(*match_batch(code, eng, batch, backend=backend))


def check_planes(code):
    """The checked correlation surfaces of all complexes as (int32 offsets [n_planes + 1], int32 global qubit ids)."""
    cached = code.__dict__.get("_check_planes")
    if cached is None:
        off, q, qoff = [0], [], 0
        for e in code.ec:
            ct = code.lattice.complexes[e]
            for plane in ct.planes:
                q.append(np.asarray(plane, np.int64) + qoff)
                off.append(off[-1] + len(plane))
            qoff += ct.Q
        cached = code.__dict__["_check_planes"] = (
            np.asarray(off, np.int32), np.ascontiguousarray(np.concatenate(q) if q else np.zeros(0), np.int32))
    return cached


def logical_failures(code, batch, flips):
    """``check_correction`` over a batch (``decoders/decoder.py:151-224``): bool[T], True where any checked
    correlation surface of any complex has odd parity after the recovery.  ``batch.bits`` and ``flips`` may
    each be unpacked uint8[T, Qtot] or the packed words uint32[T, ceil(Qtot/32)] as they come off the
    device; only the planes' qubits are touched (``fpb_uf_check_correction``, host threads over trials)."""
    from . import uf_native

    off, q = check_planes(code)
    Qtot = sum(code.lattice.complexes[e].Q for e in code.ec)
    return uf_native.check_correction(batch.bits, flips, off, q, n_qubits=Qtot)
