"""TEST INFRASTRUCTURE ONLY - logical-failure rates of the Union-Find decoder measured on the LIVE reference
(build container only; rustworkx's PyGraph calls answered by the stand-in in oracle/ref_shim.py).

    python -m oracle.gen_failure_rates_uf_ref     # writes tests/golden/failure_rates_uf_reference.json

Eight processes stand in for the reference's mpi4py ranks; each runs the reference's loop body
(flamingpy/simulations.py:46-59: `noise.apply_noise(rng)`, `correct(code, "UF", {"method": "uniform"})`) on one
reused noise layer ("compat" rates, SURVEY.md finding 0.3).  Beside the reference's count the host decoder of
this repository (csrc/uf.cpp) decodes the very same syndromes and erasures: `failures_same_noise` differs from
`failures` only where a grown cluster holds a cycle or two boundary points, i.e. where the outcome hangs on the
spanning tree picked (the reference's own choice comes from rustworkx and from Python set order).
"""

import json
import os
import sys
import time
from multiprocessing import Pool

# d, boundaries, delta, p_swap, trials per process
CASES = [
    (3, "open", 0.09, 0.25, 2000),
    (3, "open", 0.1, 0.0, 2000),
    (5, "open", 0.09, 0.2, 400),
    (4, "periodic", 0.09, 0.2, 500),
]


def work(args):
    d, b, delta, ps, n, seed = args
    import numpy as np

    from flamingpy_b200 import uf_native
    from flamingpy_b200.lattice import build_lattice
    from oracle import ref_shim

    ref_shim.load()
    from flamingpy.decoders.decoder import assign_weights, correct
    from flamingpy.noise import CVLayer

    code = ref_shim.make_code(d, "primal", b)
    lat = build_lattice(d, "primal", b)
    ct = lat.complexes["primal"]
    pts = [tuple(int(v) for v in lat.coords[i]) for i in ct.qubits]
    nodes = code.graph.nodes
    layer = CVLayer(code, delta=delta, p_swap=ps)
    rng = np.random.default_rng(seed)
    fails = mine = 0
    for _ in range(n):
        layer.apply_noise(rng)
        assign_weights(code, "UF", method="uniform")
        bits = np.array([int(nodes[p]["bit_val"]) for p in pts], np.uint8)
        erased = np.array([nodes[p]["weight"] == -1 for p in pts], np.uint8)
        parity = np.array([s.parity for s in code.primal_stabilizers], np.uint8)
        flips = uf_native.decode_batch(ct, parity[None], erased[None], threads=1)[0]
        mine += any(int((bits ^ flips)[pl].sum()) % 2 for pl in ct.planes)
        fails += not correct(code, "UF", weight_options={"method": "uniform"})
    return fails, mine


if __name__ == "__main__":
    out = []
    with Pool(8) as pool:
        for d, b, delta, ps, n in CASES:
            t0 = time.time()
            res = pool.map(work, [(d, b, delta, ps, n, 500 + s) for s in range(8)])
            rec = {"noise": "CVLayer", "decoder": "UF", "distance": d, "boundaries": b, "delta": delta, "p_swap": ps,
                   "weight_opts": {"method": "uniform"}, "mode": "compat", "failures": int(sum(r[0] for r in res)),
                   "failures_same_noise": int(sum(r[1] for r in res)), "trials": 8 * n, "seeds": "500..507",
                   "trials_per_chain": n, "source": "live reference (unionfind package, PyGraph stand-in)",
                   "seconds": round(time.time() - t0, 1)}
            print(rec, file=sys.stderr, flush=True)
            out.append(rec)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "failure_rates_uf_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
