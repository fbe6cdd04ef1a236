"""Complete error-correction trials per second with the Union-Find decoder (device: noise -> syndrome; host:
csrc/uf.cpp on all cores, or with --backend device the k_uf kernel).
    python scripts/uf_probe.py [distance] [trials] [--backend device]"""
import sys, time; sys.path.insert(0, ".")
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.noise import CVLayer
from flamingpy_b200.simulations import ec_monte_carlo

backend = "native"
if "--backend" in sys.argv:
    i = sys.argv.index("--backend")
    backend = sys.argv[i + 1]
    del sys.argv[i:i + 2]
d = int(sys.argv[1]) if len(sys.argv) > 1 else 13
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
for bnd, ps in (("open", 0.2), ("periodic", 0.0)):
    code = SurfaceCode(d, "primal", bnd)
    layer = CVLayer(code, delta=0.09, p_swap=ps)
    args = {"weight_opts": {"method": "uniform"}}
    ec_monte_carlo(2048, code, layer, "UF", args, seed=1, batch=1024, mode="fresh", backend=backend)  # warm-up
    t0 = time.perf_counter()
    e, st = ec_monte_carlo(trials, code, layer, "UF", args, seed=2, batch=4096, mode="fresh", return_stats=True, backend=backend)
    dt = time.perf_counter() - t0
    print(f"UF ({st['uf_backend']} decoder) d={d} {bnd} p_swap={ps}: {trials} trials in {dt:.2f} s = {trials / dt:.0f} trials/s, failures {e} "
          f"({e / trials:.4f}), device {st['device_ms']:.0f} ms", flush=True)
