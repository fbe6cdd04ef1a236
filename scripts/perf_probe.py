"""Quick device-side throughput probe (developer tool, not the bench)."""
import sys, time, argparse
import numpy as np
sys.path.insert(0, ".")
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.engine import TrialEngine

ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=13)
ap.add_argument("--ec", default="both")
ap.add_argument("--bound", default="periodic")
ap.add_argument("--delta", type=float, default=0.09)
ap.add_argument("--pswap", type=float, default=0.0)
ap.add_argument("--mode", default="compat")
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--step-factor", type=float, nargs="*", default=[2.0])
a = ap.parse_args()
lat = build_lattice(a.d, a.ec, a.bound)
for sf in a.step_factor:
    eng = TrialEngine(lat, a.delta, a.pswap, {"method": "blueprint", "delta": a.delta}, mode=a.mode, delta_step=sf)
    for s in range(a.steps):
        t0 = time.perf_counter()
        eng.run_rng(1, s * a.batch, a.batch, fetch=False)
        wall = time.perf_counter() - t0
        tm = eng.timings(); sz = eng.sizes()
        print(f"d={a.d} {a.ec} {a.bound} ps={a.pswap} mode={a.mode} sf={sf} B={a.batch} wall={wall*1e3:.1f}ms "
              f"gen={tm['gen_ms']:.2f} noise={tm['noise_ms']:.2f} synd={tm['syndrome_ms']:.2f} graph={tm['graph_ms']:.2f} "
              f"total={tm['total_ms']:.2f}ms -> {a.batch/tm['total_ms']*1e3:.0f} trials/s; K/trial={[sz.total_odd[i]/a.batch for i in range(sz.n_complex)]} "
              f"warps={tm['graph_warps_per_cta']} ctas={tm['graph_ctas']} smem={tm['graph_smem_bytes']}", flush=True)
    eng.close()
