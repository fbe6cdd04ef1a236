"""TEST INFRASTRUCTURE ONLY - run the live reference and dump per-trial arrays.

``capture_trial`` drives the reference's own ``CVLayer.apply_noise``,
``assign_weights`` and ``NxMatchingGraph`` (and optionally ``correct``) on a
reference ``SurfaceCode`` and returns everything in the index order of
``flamingpy_b200.lattice`` so the restatement (``oracle/restate.py``) and the
CUDA path can be compared on the same injected noise.  Used by the parity tests
and by ``oracle/gen_golden.py``; needs ``/root/reference``.
"""

from __future__ import annotations

import copy

import numpy as np

from . import ref_shim


class RecordingRng:
    """Pass-through wrapper around ``numpy.random.Generator`` that keeps the
    arrays returned by ``normal`` (the raw quadrature draws)."""

    def __init__(self, rng):
        self._rng = rng
        self.normals = []

    def normal(self, *a, **k):
        out = self._rng.normal(*a, **k)
        self.normals.append(np.array(out, copy=True))
        return out

    def __getattr__(self, name):
        return getattr(self._rng, name)


def capture_trial(code, layer, lat, rng, weight_opts, run_correct=False, want_paths=False):
    """One reference trial.  Returns a dict of arrays (see keys below)."""
    ref_shim.load()
    from flamingpy.decoders.decoder import assign_weights, correct
    from flamingpy.decoders.mwpm.matching import NxMatchingGraph

    rec = RecordingRng(rng)
    layer.apply_noise(rec)
    N = lat.N
    x = rec.normals[-1]
    state = np.zeros(N, np.uint8)
    state[np.asarray(layer.states["GKP"], int)] = 1
    state[np.asarray(layer.states["p"], int)] = 2
    sigma = np.array(layer._init_covs, copy=True)
    pts = [tuple(int(v) for v in c) for c in lat.coords]
    nodes = code.graph.nodes
    label_p = np.array([nodes[p].get("state") == "p" for p in pts])

    assign_weights(code, "MWPM", **weight_opts)
    out = {"x": x, "state": state, "sigma": sigma, "label_p": label_p, "complex": {}}
    for ec in code.ec:
        ct = lat.complexes[ec]
        qpts = [pts[i] for i in ct.qubits]
        hom = np.array([nodes[p]["hom_val_p"] for p in qpts], np.float64)
        bits = np.array([int(nodes[p]["bit_val"]) for p in qpts], np.uint8)
        w = np.array([float(nodes[p]["weight"]) for p in qpts], np.float64)
        stabs = getattr(code, ec + "_stabilizers")
        parity = np.array([s.parity for s in stabs], np.uint8)
        mg = NxMatchingGraph(ec, code)
        pos = {id(s): i for i, s in enumerate(stabs)}
        odd = np.nonzero(parity)[0].astype(np.int32)
        K = len(odd)
        pair_len = np.empty(K * (K - 1) // 2, np.float64)
        pair_paths = {}
        k = 0
        for a in range(K - 1):
            for b in range(a + 1, K):
                e = (stabs[odd[a]], stabs[odd[b]])
                pair_len[k] = mg.edge_weight(e)
                if want_paths:
                    pair_paths[(a, b)] = [pos[id(s)] for s in mg.edge_path(e)]
                k += 1
        bnd_len = np.empty(0, np.float64)
        bnd_side = np.empty(0, np.uint8)
        n_edges = mg.graph.number_of_edges()
        if ct.has_boundary and K:
            sg = getattr(code, ec + "_stab_graph")
            low = set(sg.low_bound_points)
            bnd_len = np.empty(K, np.float64)
            bnd_side = np.empty(K, np.uint8)
            for i, vp in enumerate(mg.virtual_points):
                bnd_len[i] = mg.edge_weight((stabs[odd[i]], vp))
                bnd_side[i] = 0 if vp[0] in low else 1
        rec_c = {
            "hom": hom, "bits": bits, "weights": w, "parity": parity, "odd": odd,
            "pair_len": pair_len, "bnd_len": bnd_len, "bnd_side": bnd_side,
            "n_edges": n_edges, "pair_paths": pair_paths,
        }
        out["complex"][ec] = rec_c
    if run_correct:
        out["success"] = bool(
            correct(code, "MWPM", weight_options=weight_opts, decoder_opts={"backend": "networkx"})
        )
    return out


def make_layer(code, delta, p_swap):
    ref_shim.load()
    from flamingpy.noise import CVLayer

    return CVLayer(code, delta=delta, p_swap=p_swap)


def capture_iid_trial(code, lat, rng, error_probability, run_correct=True):
    """One reference trial of ``IidNoise`` (``flamingpy/noise/iid.py``) decoded with uniform
    weights: the bit of every node in ``flamingpy_b200.lattice`` index order, per complex the
    parities / odd cubes / matching-graph lengths of ``NxMatchingGraph``, and ``correct``'s verdict."""
    ref_shim.load()
    from flamingpy.decoders.decoder import assign_weights, correct
    from flamingpy.decoders.mwpm.matching import NxMatchingGraph
    from flamingpy.noise import IidNoise

    IidNoise(code, error_probability).apply_noise(rng)
    pts = [tuple(int(v) for v in c) for c in lat.coords]
    nodes = code.graph.nodes
    bit_val = np.array([int(nodes[p]["bit_val"]) for p in pts], np.uint8)
    wo = {"method": "uniform", "integer": False, "multiplier": 1}
    assign_weights(code, "MWPM", **wo)
    out = {"bit_val": bit_val, "complex": {}}
    for ec in code.ec:
        ct = lat.complexes[ec]
        stabs = getattr(code, ec + "_stabilizers")
        parity = np.array([s.parity for s in stabs], np.uint8)
        mg = NxMatchingGraph(ec, code)
        odd = np.nonzero(parity)[0].astype(np.int32)
        K = len(odd)
        pair_len = np.empty(K * (K - 1) // 2, np.float64)
        k = 0
        for a in range(K - 1):
            for b in range(a + 1, K):
                pair_len[k] = mg.edge_weight((stabs[odd[a]], stabs[odd[b]]))
                k += 1
        bnd_len = np.empty(0, np.float64)
        bnd_side = np.empty(0, np.uint8)
        if ct.has_boundary and K:
            low = set(getattr(code, ec + "_stab_graph").low_bound_points)
            bnd_len = np.empty(K, np.float64)
            bnd_side = np.empty(K, np.uint8)
            for i, vp in enumerate(mg.virtual_points):
                bnd_len[i] = mg.edge_weight((stabs[odd[i]], vp))
                bnd_side[i] = 0 if vp[0] in low else 1
        out["complex"][ec] = {"parity": parity, "odd": odd, "pair_len": pair_len, "bnd_len": bnd_len,
                              "bnd_side": bnd_side}
    if run_correct:
        out["success"] = bool(correct(code, "MWPM", weight_options=wo, decoder_opts={"backend": "networkx"}))
    return out
