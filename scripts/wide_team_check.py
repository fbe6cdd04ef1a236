"""Parity (vs the C oracle) and timing of the W-warp team kernel on large lattices.
    python scripts/wide_team_check.py d boundaries trials"""
import sys, time; sys.path.insert(0, ".")
import numpy as np
from flamingpy_b200 import _lib
from flamingpy_b200.engine import TrialEngine
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.noise import sample_batch
from oracle import c_oracle
d, bnd, T = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
delta = 0.09
lat = build_lattice(d, "primal", bnd)
wo = {"method": "blueprint", "delta": delta}
x, st = sample_batch(np.random.default_rng(11), lat.N, delta, 0.0, T, fresh=True)
eng = TrialEngine(lat, delta, 0.0, wo, mode="fresh")
for rep in range(3):
    b = eng.run_injected(x, st)
tm = eng.timings()
print(f"d={d} {bnd}: graph {tm['graph_ms']:.2f} ms for {T} trials = {T / tm['graph_ms'] * 1e3:.1f} trials/s; "
      f"searches in flight per CTA {tm['graph_warps_per_cta']} smem {tm['graph_smem_bytes']}", flush=True)
t0 = time.perf_counter()
ref = c_oracle.run(lat, delta, 0.0, wo, x, st)
print(f"oracle {time.perf_counter() - t0:.1f} s", flush=True)
cb, rb = b.complexes["primal"], ref.complexes["primal"]
assert np.array_equal(cb.odd_count, rb.odd_count) and np.array_equal(cb.odd_ids, rb.odd_ids)
rel = np.max(np.abs(cb.pair_len - rb.pair_len) / np.maximum(1e-300, np.abs(rb.pair_len))) if len(rb.pair_len) else 0.0
print("pair_len max rel diff", rel, "pairs", len(rb.pair_len))
assert rel < 1e-9
if len(rb.bnd_len):
    relb = np.max(np.abs(cb.bnd_len - rb.bnd_len) / np.abs(rb.bnd_len))
    print("bnd_len max rel diff", relb, "side mismatches", int((cb.bnd_side != rb.bnd_side).sum()))
    assert relb < 1e-9
print("PARITY_OK")
