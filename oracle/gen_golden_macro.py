"""TEST INFRASTRUCTURE ONLY - generate tests/golden/macro_*.npz from the live reference's CVMacroLayer.

    python -m oracle.gen_golden_macro        (build container only: needs /root/reference)

The unmodified reference class runs with ``thewalrus.symplectic.beam_splitter`` / ``expand`` replaced
by their closed forms (``oracle/ref_shim.py``; thewalrus is not installable offline, and the fixed
four-splitter is all the reference uses them for).  Each file holds a chain of consecutive trials on
one reused reference layer: the random draws to inject (micronode labels, float32 initial means,
standard normals - replayed from the same NumPy seed by ``flamingpy_b200.noise.MacroSampler``, whose
call order the comparison with the reference's outputs pins) and everything the reference computed:
reduced state labels, bit values, ``p_phase_cond``, the weights of ``assign_weights(prob_precomputed=
True)``, parities, odd cubes, matching-graph lengths and the verdict of ``correct``.
"""

import os

import numpy as np

from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.noise import MacroSampler
from oracle import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, d, ec, boundaries, delta, p_swap, weight opts, trials, seed
CASES = [
    ("macro_d3_open_primal_ps025", 3, "primal", "open", 0.09, 0.25, {"method": "blueprint", "prob_precomputed": True}, 8, 201),
    ("macro_d3_open_dual_ps0", 3, "dual", "open", 0.1, 0, {"method": "blueprint", "prob_precomputed": True}, 6, 202),
    ("macro_d4_periodic_primal_ps02", 4, "primal", "periodic", 0.1, 0.2, {"method": "blueprint", "prob_precomputed": True}, 4, 203),
    ("macro_d3_toric_dual_ps05", 3, "dual", "toric", 0.12, 0.5, {"method": "blueprint", "prob_precomputed": True}, 4, 204),
    ("macro_d5_open_primal_ps01", 5, "primal", "open", 0.09, 0.1, {"method": "blueprint", "prob_precomputed": True}, 3, 205),
    ("macro_d234_open_primal_ps1", (2, 3, 4), "primal", "open", 0.1, 1, {"method": "blueprint", "prob_precomputed": True}, 3, 206),
    ("macro_d3_open_primal_int", 3, "primal", "open", 0.09, 0.25,
     {"method": "blueprint", "prob_precomputed": True, "integer": True, "multiplier": 100}, 4, 207),
    ("macro_d3_periodic_both_uniform", 3, "both", "periodic", 0.11, 0.15, {"method": "uniform"}, 4, 208),
]


def capture_macro_trial(code, layer, lat, rng, weight_opts, run_correct=True):
    ref_shim.load()
    from flamingpy.decoders.decoder import assign_weights, correct
    from flamingpy.decoders.mwpm.matching import NxMatchingGraph

    layer.apply_noise(rng)
    pts = [tuple(int(v) for v in c) for c in lat.coords]
    nodes = code.graph.nodes
    out = {
        "red_state": np.array([2 if nodes[p]["state"] == "p" else 1 for p in pts], np.uint8),
        "bit_val": np.array([int(nodes[p]["bit_val"]) for p in pts], np.uint8),
        "p_phase_cond": np.array([float(nodes[p]["p_phase_cond"]) for p in pts], np.float64),
        "complex": {},
    }
    assign_weights(code, "MWPM", **weight_opts)
    for ec in code.ec:
        ct = lat.complexes[ec]
        qpts = [pts[i] for i in ct.qubits]
        w = np.array([float(nodes[p]["weight"]) for p in qpts], np.float64)
        stabs = getattr(code, ec + "_stabilizers")
        parity = np.array([s.parity for s in stabs], np.uint8)
        mg = NxMatchingGraph(ec, code)
        odd = np.nonzero(parity)[0].astype(np.int32)
        K = len(odd)
        pair_len = np.array([mg.edge_weight((stabs[odd[a]], stabs[odd[b]])) for a in range(K - 1) for b in range(a + 1, K)],
                            np.float64)
        bnd_len = np.empty(0, np.float64)
        bnd_side = np.empty(0, np.uint8)
        if ct.has_boundary and K:
            low = set(getattr(code, ec + "_stab_graph").low_bound_points)
            bnd_len = np.array([mg.edge_weight((stabs[odd[i]], vp)) for i, vp in enumerate(mg.virtual_points)], np.float64)
            bnd_side = np.array([0 if vp[0] in low else 1 for vp in mg.virtual_points], np.uint8)
        out["complex"][ec] = {"weights": w, "parity": parity, "odd": odd, "pair_len": pair_len, "bnd_len": bnd_len,
                              "bnd_side": bnd_side}
    if run_correct:
        out["success"] = bool(correct(code, "MWPM", weight_options=weight_opts, decoder_opts={"backend": "networkx"}))
    return out


def generate(case):
    name, d, ec, b, delta, p_swap, wopts, trials, seed = case
    ref_shim.load()
    from flamingpy.noise import CVMacroLayer

    code = ref_shim.make_code(d, ec, b)
    layer = CVMacroLayer(code, delta=delta, p_swap=p_swap)
    lat = build_lattice(d, ec, b)
    sampler = MacroSampler(4 * lat.N, p_swap)
    rng_ref, rng_in = np.random.default_rng(seed), np.random.default_rng(seed)
    data = {"dims": np.array(lat.dims), "ec": np.array(ec), "boundaries": np.array(b), "delta": np.array(delta),
            "p_swap": np.array(float(p_swap)), "trials": np.array(trials), "seed": np.array(seed),
            "method": np.array(wopts["method"]), "integer": np.array(bool(wopts.get("integer", False))),
            "multiplier": np.array(wopts.get("multiplier", 1)),
            "prob_precomputed": np.array(bool(wopts.get("prob_precomputed", False)))}
    ins = {"state": [], "mean": [], "z": []}
    red = {"red_state": [], "bit_val": [], "p_phase_cond": []}
    succ = []
    per = {e: {k: [] for k in ("weights", "parity", "odd", "pair_len", "bnd_len", "bnd_side")} for e in lat.ec}
    for _ in range(trials):
        r = capture_macro_trial(code, layer, lat, rng_ref, wopts)
        st, mean, z = sampler.next(rng_in)
        ins["state"].append(st), ins["mean"].append(mean), ins["z"].append(z)
        for k in red:
            red[k].append(r[k])
        succ.append(r["success"])
        for e in lat.ec:
            for k in per[e]:
                per[e][k].append(r["complex"][e][k])
    for k, v in {**ins, **red}.items():
        data[k] = np.stack(v)
    data["success"] = np.array(succ)
    for e in lat.ec:
        for k in ("weights", "parity"):
            data[f"{e}_{k}"] = np.stack(per[e][k])
        for k in ("odd", "pair_len", "bnd_len", "bnd_side"):
            arrs = per[e][k]
            data[f"{e}_{k}_off"] = np.cumsum([0] + [len(a) for a in arrs]).astype(np.int64)
            data[f"{e}_{k}"] = np.concatenate(arrs) if arrs else np.empty(0)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **data)
    return name


if __name__ == "__main__":
    import sys

    for c in CASES:
        if len(sys.argv) > 1 and c[0] not in sys.argv[1:]:
            continue
        print("wrote", generate(c))
