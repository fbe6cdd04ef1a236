import sys, time; sys.path.insert(0, ".")
import numpy as np
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.engine import TrialEngine
d = int(sys.argv[1]) if len(sys.argv) > 1 else 21
for method in ("blueprint", "uniform"):
    lat = build_lattice(d, "primal", "open")
    wo = {"method": method, "delta": 0.09}
    eng = TrialEngine(lat, 0.09, 1, wo)
    t0 = time.perf_counter(); eng.build_table(); tb = time.perf_counter() - t0
    for B in (16, 64):
        for rep in range(2):
            eng.run_rng(1, rep * B, B, fetch=False)
        tm = eng.timings(); sz = eng.sizes()
        by = 8 * sz.total_pairs[0] + 9 * sz.total_odd[0]
        print(f"d={d} {method} table build {tb:.2f}s  B={B} graph={tm['graph_ms']:.2f}ms total={tm['total_ms']:.2f}ms -> {B/tm['total_ms']*1e3:.0f} trials/s; "
              f"K/trial={sz.total_odd[0]/B:.0f}; gather writes {by/1e6:.0f} MB -> {by/tm['graph_ms']/1e6:.0f} GB/s (output bytes only)")
    eng.close()
