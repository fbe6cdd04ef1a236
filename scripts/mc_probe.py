import sys, time; sys.path.insert(0, ".")
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.noise import CVLayer
from flamingpy_b200.simulations import ec_monte_carlo
for d, b, delta, n in [(9, "open", 0.08, 2048), (13, "open", 0.09, 1024), (13, "periodic", 0.09, 512)]:
    code = SurfaceCode(d, "primal", b); noise = CVLayer(code, delta=delta, p_swap=0)
    wo = {"weight_opts": {"method": "blueprint", "delta": delta}}
    ec_monte_carlo(64, code, noise, "MWPM", wo, seed=1, mode="fresh", batch=64)
    t0 = time.perf_counter()
    e, st = ec_monte_carlo(n, code, noise, "MWPM", wo, seed=2, mode="fresh", batch=256, return_stats=True)
    dt = time.perf_counter() - t0
    print(f"d={d} {b} delta={delta}: {n} trials end to end (GPU graph + native MWPM + GPU paths + check) in {dt:.2f}s = {n/dt:.0f} trials/s; errors {e}; device {st['device_ms']:.0f} ms")
