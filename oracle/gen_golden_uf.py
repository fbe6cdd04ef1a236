"""TEST INFRASTRUCTURE ONLY - golden vectors of the reference's Union-Find decoder.

Runs the unmodified reference (``flamingpy/decoders/unionfind``; rustworkx's PyGraph calls are
answered by the stand-in in ``oracle/ref_shim.py``) on injected CVLayer noise and records, per trial
and complex, in ``flamingpy_b200.lattice`` index order: the noise draws and state labels, bit values,
stabilizer parities, the ``assign_weights(code, "UF")`` weights (2, or -1 = erased), the Support
table when the clusters stopped growing (fully grown edges - deterministic), the reference's own
recovery set (one of several valid ones: it depends on rustworkx's spanning tree and on Python set
order) and ``correct``'s verdict.  Needs /root/reference; writes tests/golden/uf_*.npz.

    python -m oracle.gen_golden_uf [names...]
"""

from __future__ import annotations

import os

import numpy as np

from flamingpy_b200.lattice import build_lattice
from oracle import capture, ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, d, ec, boundaries, delta, p_swap, trials, seed
CASES = [
    ("uf_d3_open_primal_ps025", 3, "primal", "open", 0.09, 0.25, 32, 301),
    ("uf_d4_toric_primal_ps05", 4, "primal", "toric", 0.1, 0.5, 12, 302),
    ("uf_d5_open_primal_ps03", 5, "primal", "open", 0.09, 0.3, 12, 303),
    ("uf_d3_periodic_both_ps0", 3, "both", "periodic", 0.12, 0, 12, 304),
    ("uf_d5_periodic_dual_ps02", 5, "dual", "periodic", 0.1, 0.2, 8, 305),
    ("uf_d2_periodic_both_ps0", 2, "both", "periodic", 0.12, 0, 8, 306),
    ("uf_d234_open_primal_ps01", (2, 3, 4), "primal", "open", 0.12, 0.1, 8, 307),
    ("uf_d4_open_dual_ps06", 4, "dual", "open", 0.09, 0.6, 10, 308),
    # every node p-squeezed: every edge erased, nothing to grow; odd syndromes end at a boundary point
    ("uf_d3_open_primal_ps1", 3, "primal", "open", 0.1, 1, 12, 309),
    ("uf_d223_toric_dual_ps1", (2, 2, 3), "dual", "toric", 0.1, 1, 8, 310),
]


def generate(case):
    name, d, ec, b, delta, p_swap, trials, seed = case
    ref_shim.load()
    from flamingpy.decoders.decoder import assign_weights, check_correction, recovery
    from flamingpy.decoders.unionfind import algos
    from flamingpy.decoders.unionfind.uf_classes import Boundary, Support

    code = ref_shim.make_code(d, ec, b)
    layer = capture.make_layer(code, delta, p_swap)
    lat = build_lattice(d, ec, b)
    index = lat.coord_index()
    rng = np.random.default_rng(seed)
    pts = [tuple(int(v) for v in c) for c in lat.coords]
    nodes = code.graph.nodes
    data = {"dims": np.array(lat.dims), "ec": np.array(ec), "boundaries": np.array(b), "delta": np.array(delta),
            "p_swap": np.array(float(p_swap)), "trials": np.array(trials), "seed": np.array(seed)}
    xs, states, succ = [], [], []
    per = {e: {k: [] for k in ("bits", "parity", "weights", "grown", "ref_flips")} for e in lat.ec}
    for _ in range(trials):
        rec = capture.RecordingRng(rng)
        layer.apply_noise(rec)
        xs.append(rec.normals[-1])
        state = np.zeros(lat.N, np.uint8)
        state[np.asarray(layer.states["GKP"], int)] = 1
        state[np.asarray(layer.states["p"], int)] = 2
        states.append(state)
        assign_weights(code, "UF", method="uniform")
        for e in lat.ec:
            ct = lat.complexes[e]
            local = {int(n): i for i, n in enumerate(ct.qubits)}
            qpts = [pts[i] for i in ct.qubits]
            per[e]["bits"].append(np.array([int(nodes[p]["bit_val"]) for p in qpts], np.uint8))
            per[e]["weights"].append(np.array([int(nodes[p]["weight"]) for p in qpts], np.int8))
            per[e]["parity"].append(np.array([s.parity for s in getattr(code, e + "_stabilizers")], np.uint8))
            # uf_decode (algos.py:229-257) with the Support table copied before span_forest empties it
            sg = getattr(code, e + "_stab_graph")
            node_dict, cluster_trees, odd_clusters = algos.initialize_cluster_trees(sg)
            support = Support(sg)
            boundary = {c.node: Boundary(c, support, sg) for c in cluster_trees}
            algos.union_find(odd_clusters, boundary, sg, support, node_dict)
            grown = np.zeros(ct.Q, np.uint8)
            for edge, status in support.status.items():
                a, bb = tuple(edge)
                if status == "grown":
                    grown[local[index[tuple(sg.edge_data(a, bb)["common_vertex"])]]] = 1
            per[e]["grown"].append(grown)
            forest, parity_dict = algos.obtain_spanning_forest(sg, support)
            rset = algos.peeling(forest, parity_dict)
            flips = np.zeros(ct.Q, np.uint8)
            for q in rset:
                flips[local[index[tuple(q)]]] ^= 1
            per[e]["ref_flips"].append(flips)
            recovery(rset, code, e)
            assert not list(sg.odd_parity_stabilizers())
        succ.append(bool(np.all(check_correction(code))))
    data["x"] = np.stack(xs)
    data["state"] = np.stack(states)
    data["success"] = np.array(succ)
    for e in lat.ec:
        for k, v in per[e].items():
            data[f"{e}_{k}"] = np.stack(v)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **data)
    return name, float(np.mean(succ))


if __name__ == "__main__":
    import sys

    for c in CASES:
        if len(sys.argv) > 1 and c[0] not in sys.argv[1:]:
            continue
        print("wrote", *generate(c))
