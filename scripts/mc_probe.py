"""End-to-end Monte-Carlo throughput (GPU matching graphs + native matcher + GPU path toggles +
logical check) through simulations.ec_monte_carlo.  `python scripts/mc_probe.py [dense|sparse|auto]`."""
import sys, time; sys.path.insert(0, ".")
from flamingpy_b200 import mwpm_native
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.noise import CVLayer
from flamingpy_b200.simulations import ec_monte_carlo
mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
mwpm_native.configure(mode)
for d, b, delta, n in [(9, "open", 0.08, 16384), (13, "open", 0.09, 8192), (13, "periodic", 0.09, 8192),
                       (17, "periodic", 0.09, 1024)]:
    if mode == "dense" and d > 13:
        continue
    code = SurfaceCode(d, "primal", b); noise = CVLayer(code, delta=delta, p_swap=0)
    wo = {"weight_opts": {"method": "blueprint", "delta": delta}}
    ec_monte_carlo(1024, code, noise, "MWPM", wo, seed=1, mode="fresh", batch=256)  # warm-up: engines, pinned buffers
    mwpm_native.stats(reset=True)
    t0 = time.perf_counter()
    e, st = ec_monte_carlo(n, code, noise, "MWPM", wo, seed=2, mode="fresh", batch=256, return_stats=True)
    dt = time.perf_counter() - t0
    print(f"[{mode}] d={d} {b} delta={delta}: {n} trials end to end in {dt:.2f}s = {n/dt:.0f} trials/s; errors {e}; "
          f"device {st['device_ms']:.0f} ms; matcher {mwpm_native.stats()}")
