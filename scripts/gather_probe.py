"""Device time of the static-weight gather (developer tool)."""
import sys, os
sys.path.insert(0, ".")
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.engine import TrialEngine
d = int(sys.argv[1]) if len(sys.argv) > 1 else 21
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lat = build_lattice(d, "primal", "open")
eng = TrialEngine(lat, 0.09, 1.0, {"method": "blueprint", "delta": 0.09})
eng.build_table()
for s in range(4):
    eng.run_rng(1, s * B, B, fetch=False)
    tm = eng.timings(); sz = eng.sizes()
    gb = (8 * sz.total_pairs[0] + 13 * sz.total_odd[0]) / 1e9
    print(f"d={d} B={B} groups={os.environ.get('FPB_GATHER_GROUPS','auto')} by_trial={'FPB_GATHER_BY_TRIAL' in os.environ} graph={tm['graph_ms']:.2f} ms  {gb/tm['graph_ms']*1e3:.0f} GB/s written  {B/tm['total_ms']*1e3:.0f} trials/s")
