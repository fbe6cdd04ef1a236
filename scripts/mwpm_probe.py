"""Single-core timing of the native matcher on matching graphs of real trials (generated with the
C oracle, so no GPU is needed): the solver on the complete graph ("dense" mode) vs sparse candidate graph + pricing.
    python scripts/mwpm_probe.py [reps]"""
import sys, time; sys.path.insert(0, ".")
import numpy as np
from flamingpy_b200 import mwpm_native as mw
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.matching import tri_index
from flamingpy_b200.noise import sample_batch
from oracle import c_oracle

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for d, bnd, delta, ps, T in [(9, "open", 0.08, 0.5, 8), (13, "open", 0.09, 0.0, 6), (13, "periodic", 0.09, 0.0, 6),
                             (17, "periodic", 0.09, 0.0, 2)]:
    lat = build_lattice(d, "primal", bnd)
    x, st = sample_batch(np.random.default_rng(5), lat.N, delta, ps, T, fresh=True)
    cb = c_oracle.run(lat, delta, ps, {"method": "blueprint", "delta": delta}, x, st).complexes["primal"]
    has = len(cb.bnd_len) > 0
    graphs = [(int(cb.odd_count[t]), cb.pair_len[cb.pair_off[t]:cb.pair_off[t + 1]],
               cb.bnd_len[cb.odd_off[t]:cb.odd_off[t + 1]] if has else None) for t in range(T)]
    res = {}
    for mode in ("dense", "sparse"):
        if mode == "dense" and d > 13:
            continue
        mw.configure(mode, 12)
        mw.stats(reset=True)
        best, tot = [1e9] * T, 0.0
        for r in range(reps if mode == "sparse" else 1):
            for t, (K, pl, bl) in enumerate(graphs):
                t0 = time.perf_counter(); q = mw.solve(K, pl, bl); best[t] = min(best[t], time.perf_counter() - t0)
                if r == 0:
                    tot += sum(bl[u] if v < 0 else pl[tri_index(min(u, v), max(u, v), K)] for u, v in q)
        res[mode] = (1e3 * sum(best) / T, tot)
    mw.configure("auto", 12)
    print(f"d={d} {bnd} K~{int(cb.odd_count.mean())}: " + "; ".join(f"{m} {v[0]:.2f} ms/graph (weight {v[1]:.6f})" for m, v in res.items()),
          mw.stats(), flush=True)
