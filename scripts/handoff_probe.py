"""Complete EC trials per second through ec_monte_carlo with the full-fetch and the sparse hand-off
(developer tool)."""
import sys, time, argparse
sys.path.insert(0, ".")
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.noise import CVLayer
from flamingpy_b200.simulations import ec_monte_carlo
from flamingpy_b200 import mwpm_native

ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=13)
ap.add_argument("--ec", default="primal")
ap.add_argument("--bound", default="periodic")
ap.add_argument("--delta", type=float, default=0.09)
ap.add_argument("--trials", type=int, default=4096)
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--mode", default="fresh")
ap.add_argument("--workers", type=int, nargs="*", default=[1, 2])
ap.add_argument("--reserve", type=int, nargs="*", default=[None], help="reserve_sms values to try (sparse hand-off)")
a = ap.parse_args()
code = SurfaceCode(a.d, a.ec, a.bound)
noise = CVLayer(code, delta=a.delta, p_swap=0)
args = {"weight_opts": {"method": "blueprint", "delta": a.delta}}
ec_monte_carlo(2048, code, noise, "MWPM", args, seed=5, batch=a.batch, mode=a.mode)  # warm-up
for handoff in ("full", "sparse"):
    for mw in a.workers:
        for rsv in (a.reserve if handoff == "sparse" else [None]):
            mwpm_native.stats(reset=True)
            t0 = time.perf_counter()
            f, st = ec_monte_carlo(a.trials, code, noise, "MWPM", args, seed=5, batch=a.batch, mode=a.mode, handoff=handoff,
                                   return_stats=True, match_workers=mw, reserve_sms=rsv)
            dt = time.perf_counter() - t0
            print(f"d={a.d} {a.ec} {a.bound} {handoff:6s} workers={mw} reserve={rsv}: {a.trials/dt:9.0f} trials/s  "
                  f"failures {f}  {st}", flush=True)
