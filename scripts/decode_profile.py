"""Where the host time of one decode batch goes (no overlap): device stage, fetch into pinned
buffers, matcher, path toggles, logical check.  python scripts/decode_profile.py [d] [boundaries]"""
import sys, time; sys.path.insert(0, ".")
import numpy as np
from flamingpy_b200 import _lib, mwpm_native
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.decoder import logical_failures
d = int(sys.argv[1]) if len(sys.argv) > 1 else 13
bnd = sys.argv[2] if len(sys.argv) > 2 else "open"
delta, B = 0.09, 256
code = SurfaceCode(d, "primal", bnd)
eng = code.engine(delta, 0.0, {"method": "blueprint", "delta": delta}, mode="fresh")
bufs = {}
for it in range(5):
    t = [time.perf_counter()]
    eng.run_rng(7, it * B, B, stage=_lib.STAGE_GRAPH, fetch=False); eng.synchronize() if hasattr(eng, "synchronize") else None
    t.append(time.perf_counter())
    b = eng.fetch_decode(bufs); t.append(time.perf_counter())
    cb = b.complexes["primal"]
    tr, s_, d_ = mwpm_native.solve_batch(cb.odd_count, cb.pair_len, cb.bnd_len if len(cb.bnd_len) else None); t.append(time.perf_counter())
    flips = eng.paths_toggle(tr, np.zeros(len(tr), np.int32), s_, d_); t.append(time.perf_counter())
    f = logical_failures(code, b, flips); t.append(time.perf_counter())
    names = ["device", "fetch", "matcher", "toggle", "check"]
    print(f"batch {it}: " + "  ".join(f"{n} {1e3*(t[i+1]-t[i]):.1f} ms" for i, n in enumerate(names)),
          f" K~{cb.odd_count.mean():.0f} fails {int(f.sum())} D2H {cb.pair_len.nbytes/1e6:.0f} MB", flush=True)
