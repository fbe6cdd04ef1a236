"""Fuzz the sparse hand-off (k_knn / k_pairs_at / k_price + the batch matcher, fpb_mwpm_sparse_*) against the matcher on
the fetched triangle, on random lattices large enough for the sparse path (48 odd cubes and more) and on small ones
(every pair requested): equal total matching weight per trial.  Weights with many ties (p_swap > 0, integer) included.
    FPB_LIB=tests/emu/_build/libfpb_emu.so FPB_BINDING=ctypes python scripts/emu_fuzz_handoff.py [seconds] [seed]"""
import sys, time; sys.path.insert(0, ".")
import numpy as np
from flamingpy_b200 import mwpm_native
from flamingpy_b200.engine import TrialEngine
from flamingpy_b200.lattice import build_lattice
from tests.test_mwpm_native import total_weight

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
t0, n_cfg, n_graphs, n_fallback = time.time(), 0, 0, 0
while time.time() - t0 < budget:
    bnd = str(rng.choice(["open", "toric", "periodic"]))
    dims = tuple(int(v) for v in rng.integers(3, 10, 3))
    if rng.random() < 0.5:
        dims = (dims[0],) * 3
    ec = str(rng.choice(["primal", "dual", "both"])) if bnd == "periodic" else str(rng.choice(["primal", "dual"]))
    delta = float(rng.choice([0.08, 0.1, 0.13, 0.2]))
    ps = float(rng.choice([0.0, 0.0, 0.2, 0.5]))
    ps = int(ps) if ps == 0.0 else ps
    wopts = [{"method": "blueprint", "delta": delta}, {"method": "blueprint", "delta": delta, "integer": True, "multiplier": 10},
             {"method": "uniform"}][int(rng.choice([0, 0, 1, 2]))]
    try:
        lat = build_lattice(dims, ec, bnd)
    except ValueError:
        continue
    eng = TrialEngine(lat, delta, ps, wopts, mode="fresh", device=0)
    n = int(rng.integers(2, 9))
    batch = eng.run_rng(int(rng.integers(0, 2**31)), 0, n)
    tag = (seed, n_cfg, dims, ec, bnd, delta, ps, wopts, n)
    for ci, e in enumerate(lat.ec):
        cb = batch.complexes[e]
        has_b = lat.complexes[e].has_boundary
        bl_all = cb.bnd_len if has_b else None
        st = {}
        tr, src, dst = mwpm_native.solve_batch_sparse(eng, ci, cb.odd_count, bl_all, stats=st, threads=2)
        ref = mwpm_native.solve_batch(cb.odd_count, cb.pair_len, bl_all, 2)
        for t in range(n):
            K = int(cb.odd_count[t])
            if K == 0:
                continue
            pl = cb.pairs(t)
            bl = cb.boundary(t)[0] if has_b else None
            got = list(zip(src[tr == t].tolist(), dst[tr == t].tolist()))
            exp = list(zip(ref[1][ref[0] == t].tolist(), ref[2][ref[0] == t].tolist()))
            wg, we = total_weight(K, got, pl, bl), total_weight(K, exp, pl, bl)
            assert abs(wg - we) <= 1e-9 + 1e-12 * abs(we), ("handoff", tag, e, t, K, wg, we)
            n_graphs += 1
        n_fallback += st.get("fallback_trials", 0)
    eng.close()
    n_cfg += 1
print(f"ok: {n_cfg} configurations, {n_graphs} graphs, {n_fallback} fallbacks in {time.time() - t0:.0f} s (seed {seed})")
