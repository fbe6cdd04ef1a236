"""i.i.d. Pauli-noise arm: device-stage and end-to-end throughput.  python scripts/iid_probe.py"""
import sys, time; sys.path.insert(0, ".")
import numpy as np
from flamingpy_b200 import _lib, mwpm_native
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.engine import TrialEngine
from flamingpy_b200.noise import IidNoise
from flamingpy_b200.simulations import ec_monte_carlo
for d, b, p, B in [(9, "open", 0.03, 8192), (13, "open", 0.03, 4096), (13, "periodic", 0.03, 4096), (21, "open", 0.03, 256)]:
    code = SurfaceCode(d, "primal", b)
    lat = code.lattice
    for table in (False, True):
        eng = TrialEngine(lat, 1.0, 0.0, {"method": "uniform"}, mode="fresh")
        if table:
            t0 = time.perf_counter(); eng.build_table(); tb = time.perf_counter() - t0
        for rep in range(3):
            eng.run_iid(3, rep * B, B, p, fetch=False)
        tm, sz = eng.timings(), eng.sizes()
        K = sz.total_odd[0] / B
        out_b = 8 * sz.total_pairs[0] + (9 * sz.total_odd[0] if lat.complexes["primal"].has_boundary else 0)
        print(f"d={d} {b} p={p} {'table gather' if table else 'search'}: {B / tm['total_ms'] * 1e3:.0f} trials/s "
              f"(graph {tm['graph_ms']:.2f} ms of {tm['total_ms']:.2f}; K~{K:.0f}; graph output {out_b / tm['graph_ms'] / 1e6:.0f} GB/s"
              + (f"; table built in {tb:.2f}s)" if table else ")"), flush=True)
        eng.close()
    noise = IidNoise(code, p)
    n = 8 * B if d < 21 else 2 * B
    ec_monte_carlo(2 * B, code, noise, "MWPM", {"weight_opts": {"method": "uniform"}}, seed=1, batch=B)
    mwpm_native.stats(reset=True)
    t0 = time.perf_counter()
    e = ec_monte_carlo(n, code, noise, "MWPM", {"weight_opts": {"method": "uniform"}}, seed=2, batch=B)
    dt = time.perf_counter() - t0
    print(f"   end to end: {n} trials in {dt:.2f}s = {n / dt:.0f} trials/s, failure rate {e / n:.4f}", flush=True)
