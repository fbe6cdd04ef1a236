"""Runs the two streaming kernels at full size for an ncu capture (developer tool):
k_noise on the default d=13 both-complex workload, k_gather on the d=21 all-p table workload."""
import sys; sys.path.insert(0, ".")
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.engine import TrialEngine
lat = build_lattice(13, "both", "periodic")
eng = TrialEngine(lat, 0.09, 0, {"method": "blueprint", "delta": 0.09}, mode="fresh")
for i in range(3):
    eng.run_rng(1, i * 2048, 2048, stage=1, fetch=False)
print("noise", eng.timings())
eng.close()
lat = build_lattice(21, "primal", "open")
eng = TrialEngine(lat, 0.09, 1, {"method": "blueprint", "delta": 0.09})
eng.build_table()
for i in range(3):
    eng.run_rng(1, i * 64, 64, fetch=False)
print("gather", eng.timings(), eng.sizes().total_pairs[0])
