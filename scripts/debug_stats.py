import sys; sys.path.insert(0,'.')
import numpy as np
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.engine import TrialEngine
from flamingpy_b200.noise import sample_batch
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.noise import CVLayer
from flamingpy_b200.simulations import ec_monte_carlo
lat = build_lattice(3, "primal", "open")
delta = 0.09
wo = {"method": "blueprint", "delta": delta}
eng = TrialEngine(lat, delta, 0, wo, mode="fresh")
T = 20000
b = eng.run_rng(20261019, 0, T, stage=2, want_draws=True)
x, st = sample_batch(np.random.default_rng(1), lat.N, delta, 0, T, fresh=True)
b2 = eng.run_injected(x, st, stage=2)
for name, bb, xx in (("gpu-rng", b, b.x), ("numpy", b2, x)):
    cb = bb.complexes["primal"]
    print(name, "bit rate", bb.bits.mean(), "hom var", bb.hom.var(), "x var", xx.var(), "x kurt", (xx**4).mean()/xx.var()**2,
          "mean K", cb.odd_count.mean(), "mean w", bb.weights.mean())
# correlations between q and p draws of same node, and neighbours
N = lat.N
q, p = b.x[:, :N], b.x[:, N:]
print("corr(q_i,p_i)", np.mean(q*p)/q.var(), "corr(q_i,q_{i+1})", np.mean(q[:, :-1]*q[:, 1:])/q.var(), "corr across trials", np.mean(q[:-1]*q[1:])/q.var())
print("corr(q_i^2,p_i^2)", np.corrcoef((q**2).ravel(), (p**2).ravel())[0,1])
code = SurfaceCode(3, "primal", "open"); noise = CVLayer(code, delta=delta, p_swap=0)
for mode in ("compat", "fresh"):
    for seed in (1, 2):
        e = ec_monte_carlo(8000, code, noise, "MWPM", {"weight_opts": wo}, seed=seed, mode=mode, batch=4000)
        print(mode, seed, e, e/8000)
