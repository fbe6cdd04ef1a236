"""Fuzz the sparse-candidate + pricing driver against the complete-graph driver on random matching graphs (no GPU
needed): equal total weight on every graph; every tenth graph of up to 70 cubes also against networkx.  Both paths are exercised: the triangle-based solver (``mwpm_native.solve``) and
the sparse hand-off state machine against the NumPy device stand-in of tests/test_mwpm_sparse_handoff.py.
    python scripts/mwpm_fuzz.py [seconds] [seed]"""
import sys, time; sys.path.insert(0, ".")
import numpy as np
from flamingpy_b200 import mwpm_native as mw
from flamingpy_b200.matching import solve_matching
from tests.test_mwpm_native import _random_graph, total_weight
from tests.test_mwpm_sparse_handoff import DeviceStandIn

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
t0, n_graphs, n_batches, n_nx = time.time(), 0, 0, 0
try:
    while time.time() - t0 < budget:
        boundary = bool(rng.integers(0, 2))
        K = int(rng.choice([4, 6, 9, 14, 22, 40, 64, 100, 150, 240, 400], p=[.1, .1, .1, .1, .1, .15, .1, .1, .07, .05, .03]))
        K += int(rng.integers(0, 3))
        if not boundary and K % 2:
            K += 1
        kind = int(rng.integers(0, 4))
        pair, bnd = _random_graph(rng, K, kind)
        b = bnd if boundary else None
        mw.configure("dense")
        ref = mw.solve(K, pair, b)
        knn = int(rng.integers(2, 14))
        mw.configure("sparse", knn=knn)
        got = mw.solve(K, pair, b)
        wr, wg = total_weight(K, ref, pair, bnd), total_weight(K, got, pair, bnd)
        assert abs(wr - wg) <= 1e-9 + 1e-12 * abs(wr), ("solve", seed, n_graphs, K, kind, boundary, knn, wr, wg)
        n_graphs += 1
        if n_graphs % 10 == 0 and K <= 70:  # an independent implementation
            wn = total_weight(K, solve_matching(K, pair, b, backend="networkx"), pair, bnd)
            assert abs(wr - wn) <= 1e-9 + 1e-9 * abs(wn), ("networkx", seed, n_graphs, K, kind, boundary, wr, wn)
            n_nx += 1
        if n_graphs % 5 == 0:  # the hand-off path, a few graphs per batch
            Ks = [int(k) + (int(k) % 2 if not boundary else 0) for k in rng.choice([0, 3, 30, 50, 70, 110, 160], 4)]
            graphs = [_random_graph(rng, k, int(rng.integers(0, 4))) if k else (np.empty(0), np.empty(0)) for k in Ks]
            dev = DeviceStandIn(Ks, [g[0] for g in graphs], [g[1] if boundary else None for g in graphs])
            bl = np.concatenate([g[1] for g in graphs]) if boundary else None
            mw.configure("auto", knn=12)
            tr, src, dst = mw.solve_batch_sparse(dev, 0, np.array(Ks, np.int32), bl, threads=2)
            mw.configure("dense")
            rf = mw.solve_batch(np.array(Ks, np.int32), np.concatenate([g[0] for g in graphs]), bl, 2)
            for t, k in enumerate(Ks):
                if not k:
                    continue
                w1 = total_weight(k, list(zip(src[tr == t].tolist(), dst[tr == t].tolist())), graphs[t][0], graphs[t][1])
                w0 = total_weight(k, list(zip(rf[1][rf[0] == t].tolist(), rf[2][rf[0] == t].tolist())), graphs[t][0], graphs[t][1])
                assert abs(w1 - w0) <= 1e-9 + 1e-12 * abs(w0), ("handoff", seed, n_batches, t, k, w0, w1)
            n_batches += 1
finally:
    mw.configure("auto", knn=12)
print(f"ok: {n_graphs} graphs ({n_nx} also against networkx), {n_batches} hand-off batches in {time.time() - t0:.0f} s (seed {seed})")
