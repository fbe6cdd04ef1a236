"""One small batch through every kernel variant (for compute-sanitizer runs)."""
import sys; sys.path.insert(0, ".")
import numpy as np
from flamingpy_b200 import _lib
from flamingpy_b200.engine import TrialEngine
from flamingpy_b200.lattice import build_lattice
wo = {"method": "blueprint", "delta": 0.09}
for d, ec, b, T in [(5, "primal", "open", 8), (9, "primal", "open", 4), (13, "both", "periodic", 2), (15, "primal", "periodic", 1),
                    (19, "primal", "periodic", 1), (21, "primal", "open", 1), (25, "primal", "open", 1)]:
    lat = build_lattice(d, ec, b)
    eng = TrialEngine(lat, 0.09, 0.0, wo, mode="fresh")
    batch = eng.run_rng(3, 0, T)
    K = [int(batch.complexes[e].odd_count[0]) for e in lat.ec]
    if K[0] >= 2:
        eng.paths_toggle([0, 0], [0, 0], [0, 1], [1, -1] if lat.complexes[lat.ec[0]].has_boundary else [1, 0])
    print("ok", d, ec, b, "K", K, "searches per CTA", eng.timings()["graph_warps_per_cta"], flush=True)
    eng.close()
lat = build_lattice(7, "primal", "open")
eng = TrialEngine(lat, 1.0, 0.0, {"method": "uniform"}, mode="fresh")
eng.run_iid(1, 0, 16, 0.05)
eng.build_table()
b = eng.run_iid(1, 0, 16, 0.05)
eng.run_bits(b.bits)
print("ok iid + table", flush=True)
