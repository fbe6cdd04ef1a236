"""Native blossom matcher against networkx on random matching graphs (the reference's own
cross-backend check, tests/decoders/test_matching.py:74-131: equal total weight)."""

import itertools as it

import numpy as np
import pytest

from flamingpy_b200 import mwpm_native
from flamingpy_b200.matching import solve_matching, tri_index


def total_weight(K, pairs, pair_len, bnd_len):
    tot = 0.0
    seen = set()
    for u, v in pairs:
        assert u not in seen
        seen.add(u)
        if v >= 0:
            assert v not in seen
            seen.add(v)
            a, b = (u, v) if u < v else (v, u)
            tot += pair_len[tri_index(a, b, K)]
        else:
            tot += bnd_len[u]
    assert seen == set(range(K))  # perfect
    return tot


@pytest.mark.parametrize("K", [1, 2, 3, 4, 7, 10, 17, 30, 41])
@pytest.mark.parametrize("boundary", [False, True])
def test_native_equals_networkx_total_weight(K, boundary):
    if not boundary and K % 2:
        pytest.skip("odd K without boundary has no perfect matching")
    rng = np.random.default_rng(K * 2 + boundary)
    for rep in range(6):
        if rep % 3 == 0:
            pair = rng.integers(0, 10, K * (K - 1) // 2).astype(float)  # many ties (reference uses 0..9)
            bnd = rng.integers(0, 6, K).astype(float)
        elif rep % 3 == 1:
            # metric-like: points in 3-D with L1 distances, boundary = distance to a wall
            pts = rng.uniform(0, 10, (K, 3))
            pair = np.array([np.abs(pts[a] - pts[b]).sum() for a, b in it.combinations(range(K), 2)])
            bnd = np.minimum(pts[:, 0], 10 - pts[:, 0]) + 0.1
        else:
            pair = rng.uniform(0.5, 60, K * (K - 1) // 2)
            bnd = rng.uniform(0.5, 30, K)
        b = bnd if boundary else None
        got = mwpm_native.solve(K, pair, b)
        ref = solve_matching(K, pair, b, backend="networkx")
        wg, wr = total_weight(K, got, pair, bnd), total_weight(K, ref, pair, bnd)
        assert wg == pytest.approx(wr, rel=1e-9, abs=1e-9), (K, rep)
        assert sorted(solve_matching(K, pair, b, backend="native")) == sorted(got)


def test_batch_interface_matches_single_calls():
    rng = np.random.default_rng(0)
    Ks = [0, 4, 1, 9, 6, 0, 13]
    pairs = [rng.uniform(1, 20, k * (k - 1) // 2) for k in Ks]
    bnds = [rng.uniform(1, 10, k) for k in Ks]
    trial, src, dst = mwpm_native.solve_batch(np.array(Ks, np.int32), np.concatenate(pairs), np.concatenate(bnds), 2)
    for t, k in enumerate(Ks):
        sel = trial == t
        got = sorted(zip(src[sel].tolist(), dst[sel].tolist()))
        assert got == sorted(mwpm_native.solve(k, pairs[t], bnds[t]))
    # without boundary lengths
    Ks2 = [2, 6, 0, 8]
    pairs2 = [rng.uniform(1, 20, k * (k - 1) // 2) for k in Ks2]
    trial, src, dst = mwpm_native.solve_batch(np.array(Ks2, np.int32), np.concatenate(pairs2), None, 2)
    assert np.all(dst >= 0) and len(trial) == sum(Ks2) // 2


def test_odd_graph_without_boundary_is_an_error():
    with pytest.raises(ValueError):
        mwpm_native.solve(3, np.ones(3), None)


def _random_graph(rng, K, kind):
    if kind == 0:  # small integers: many ties
        return rng.integers(0, 10, K * (K - 1) // 2).astype(float), rng.integers(0, 6, K).astype(float)
    if kind == 1:  # metric, continuous
        pts = rng.uniform(0, 10, (K, 3))
        d = np.abs(pts[:, None, :] - pts[None, :, :]).sum(-1)
        return d[np.triu_indices(K, 1)], np.minimum(pts[:, 0], 10 - pts[:, 0]) + 0.1
    if kind == 2:  # no structure at all
        return rng.uniform(0.5, 60, K * (K - 1) // 2), rng.uniform(0.5, 30, K)
    pts = rng.integers(0, 7, (K, 3)).astype(float)  # lattice points: metric with massive ties
    d = np.abs(pts[:, None, :] - pts[None, :, :]).sum(-1)
    return d[np.triu_indices(K, 1)], np.minimum(pts[:, 0], 6 - pts[:, 0]) + 1.0


@pytest.mark.parametrize("K", [4, 5, 6, 9, 16, 23, 38, 60, 97, 150])
@pytest.mark.parametrize("boundary", [False, True])
def test_sparse_pricing_solver_finds_the_dense_optimum(K, boundary):
    """The sparse solver (k-nearest candidate graph + pricing against the complete graph) must
    return a perfect matching of exactly the dense solver's total weight."""
    if not boundary and K % 2:
        pytest.skip("odd K without boundary has no perfect matching")
    rng = np.random.default_rng(100 + K)
    try:
        for rep in range(8):
            pair, bnd = _random_graph(rng, K, rep % 4)
            b = bnd if boundary else None
            mwpm_native.configure("dense")
            ref = mwpm_native.solve(K, pair, b)
            mwpm_native.configure("sparse", knn=int(rng.integers(2, 14)))
            got = mwpm_native.solve(K, pair, b)
            wr, wg = total_weight(K, ref, pair, bnd), total_weight(K, got, pair, bnd)
            assert wg == pytest.approx(wr, rel=1e-12, abs=1e-9), (K, rep)
    finally:
        mwpm_native.configure("auto", knn=12)


def test_sparse_solver_on_oracle_matching_graphs():
    """Matching graphs of real trials (C oracle, d=7 periodic and open): the sparse solver's
    matchings weigh what the dense solver's do, and it rarely needs a retry with a denser candidate graph."""
    from flamingpy_b200.lattice import build_lattice
    from flamingpy_b200.noise import sample_batch
    from oracle import c_oracle

    try:
        for bnd_kind in ("periodic", "open"):
            lat = build_lattice(7, "primal", bnd_kind)
            rng = np.random.default_rng(3)
            x, st = sample_batch(rng, lat.N, 0.1, 0.3, 12, fresh=True)
            b = c_oracle.run(lat, 0.1, 0.3, {"method": "blueprint", "delta": 0.1}, x, st)
            cb = b.complexes[lat.ec[0]]
            has = len(cb.bnd_len) > 0
            mwpm_native.stats(reset=True)
            for t in range(12):
                K = int(cb.odd_count[t])
                if K < 4:
                    continue
                pl = cb.pair_len[cb.pair_off[t]:cb.pair_off[t + 1]]
                bl = cb.bnd_len[cb.odd_off[t]:cb.odd_off[t + 1]] if has else None
                mwpm_native.configure("dense")
                ref = mwpm_native.solve(K, pl, bl)
                mwpm_native.configure("sparse", knn=12)
                got = mwpm_native.solve(K, pl, bl)
                assert total_weight(K, got, pl, bl) == pytest.approx(total_weight(K, ref, pl, bl), rel=1e-12, abs=1e-9)
            s = mwpm_native.stats()
            assert s["sparse_solves"] > 0 and s["retries"] <= s["sparse_solves"] // 4
    finally:
        mwpm_native.configure("auto", knn=12)


def test_supplied_candidates_give_the_same_optimum():
    """fpb_mwpm_batch_cand: candidate lists (here: the true nearest partners, then deliberately bad
    ones) only seed the sparse solver; pricing keeps the total weight at the dense optimum."""
    rng = np.random.default_rng(42)
    Ks = [60, 75, 0, 52, 130]
    graphs = [_random_graph(rng, k, 1 + (i % 3)) if k else (np.empty(0), np.empty(0)) for i, k in enumerate(Ks)]
    pair = np.concatenate([g[0] for g in graphs])
    bnd = np.concatenate([g[1] for g in graphs])
    kk = 12

    def lists(bad):
        rows = []
        for k, (pl, bl) in zip(Ks, graphs):
            if not k:
                continue
            d = np.full((k, k), np.inf)
            iu = np.triu_indices(k, 1)
            d[iu] = np.minimum(pl, bl[iu[0]] + bl[iu[1]])
            d = np.minimum(d, d.T)
            order = np.argsort(d if not bad else -np.where(np.isinf(d), -1, d), axis=1, kind="stable")[:, :kk]
            rows.append(order.astype(np.int32))
        return np.concatenate(rows) if rows else np.empty((0, kk), np.int32)

    try:
        mwpm_native.configure("dense")
        ref = mwpm_native.solve_batch(np.array(Ks, np.int32), pair, bnd, 2)
        mwpm_native.configure("sparse", knn=12)
        for bad in (False, True):
            got = mwpm_native.solve_batch(np.array(Ks, np.int32), pair, bnd, 2, cand=lists(bad))
            off_p = np.concatenate([[0], np.cumsum([k * (k - 1) // 2 for k in Ks])])
            off_o = np.concatenate([[0], np.cumsum(Ks)])
            for t, k in enumerate(Ks):
                if not k:
                    continue
                pl, bl = pair[off_p[t]:off_p[t + 1]], bnd[off_o[t]:off_o[t + 1]]
                w = [total_weight(k, list(zip(r[1][r[0] == t].tolist(), r[2][r[0] == t].tolist())), pl, bl)
                     for r in (ref, got)]
                assert w[1] == pytest.approx(w[0], rel=1e-12, abs=1e-9), (t, bad)
    finally:
        mwpm_native.configure("auto", knn=12)


def test_matcher_threads_are_shared_between_local_ranks(monkeypatch):
    import os

    n = len(os.sched_getaffinity(0))
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    assert mwpm_native.default_threads() == n
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert mwpm_native.default_threads() == max(1, n // 8)


def test_invariant_checking_build_survives_fuzz(tmp_path):
    """The sparse solver compiled with -DFPB_MWPM_CHECK verifies, after every dual adjustment and every
    scanned edge, dual feasibility on all edges and the tree of every labelled blossom (it aborts on a
    violation); a short fuzz run against the dense solver goes through it in a subprocess."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = str(tmp_path / "libfpb_mwpm_check.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-DFPB_MWPM_CHECK",
                           "-D_GLIBCXX_ASSERTIONS", "-o", lib, os.path.join(root, "flamingpy_b200", "csrc", "mwpm.cpp")])
    env = dict(os.environ, FPB_MWPM_LIB=lib)
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "mwpm_fuzz.py"), "8", "7"], cwd=root, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok:"), out.stderr[-2000:]
