"""Target for ncu: one batch of the W-warp team kernel.  python scripts/teamw_profile_target.py d boundaries trials"""
import sys; sys.path.insert(0, ".")
from flamingpy_b200.engine import TrialEngine
from flamingpy_b200.lattice import build_lattice
d, b, T = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
lat = build_lattice(d, "primal", b)
eng = TrialEngine(lat, 0.09, 0.0, {"method": "blueprint", "delta": 0.09}, mode="fresh")
for rep in range(3):
    eng.run_rng(5, rep * T, T, fetch=False)
print(eng.timings(), eng.sizes().total_odd[0])
