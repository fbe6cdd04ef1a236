"""Sparse hand-off on the GPU (``fpb_knn_len`` / ``fpb_pairs_at`` / ``fpb_price`` + ``fpb_mwpm_sparse_*``):
the device kernels against NumPy, and complete error-correction runs whose matching graphs never leave
HBM against the full-fetch path - equal total matching weight per trial, equal failure counts."""

import numpy as np
import pytest

from flamingpy_b200 import _lib, mwpm_native
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.engine import TrialEngine
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.matching import tri_index
from flamingpy_b200.noise import CVLayer
from flamingpy_b200.simulations import ec_monte_carlo
from tests.test_mwpm_native import total_weight

pytestmark = pytest.mark.gpu


def _reduced(K, pl, bl):
    d = np.full((K, K), np.inf)
    iu = np.triu_indices(K, 1)
    d[iu] = pl if bl is None else np.minimum(pl, bl[iu[0]] + bl[iu[1]])
    return np.minimum(d, d.T)


@pytest.mark.parametrize("d,ec,b", [(7, "primal", "open"), (6, "both", "periodic")])
def test_device_handoff_kernels_against_numpy(d, ec, b):
    lat = build_lattice(d, ec, b)
    wo = {"method": "blueprint", "delta": 0.1}
    eng = TrialEngine(lat, 0.1, 0.0, wo, mode="fresh")
    batch = eng.run_rng(3, 0, 6)
    rng = np.random.default_rng(0)
    for ci, e in enumerate(lat.ec):
        cb = batch.complexes[e]
        has_b = lat.complexes[e].has_boundary
        ids, lens, wmax = eng.knn_len(ci, 8)
        y = rng.integers(-2**36, -2**30, int(cb.odd_off[-1])).astype(np.int64)
        scale = np.full(batch.n_trials, 2.0**32)
        wshift = (np.arange(batch.n_trials) % 3).astype(np.int32)
        active = (np.arange(batch.n_trials) % 2 == 0).astype(np.uint8)
        top = np.where(rng.random(len(y)) < 0.5, rng.integers(0, 4, len(y)), -1).astype(np.int32)
        ztop = rng.integers(0, 2**31, len(y)).astype(np.int64)
        vt, va, vb, vl = eng.price(ci, y, scale, wshift, active, cap=16, top=top, ztop=ztop)  # small cap: retry path
        got = sorted(zip(vt.tolist(), va.tolist(), vb.tolist(), vl.tolist()))
        exp = []
        qt, qa, qb = [], [], []
        for t in range(batch.n_trials):
            K = int(cb.odd_count[t])
            pl = cb.pairs(t)
            bl = cb.boundary(t)[0] if has_b else None
            red = _reduced(K, pl, bl)
            sl = slice(int(cb.odd_off[t]), int(cb.odd_off[t + 1]))
            order = np.argsort(red, axis=1, kind="stable")[:, :8]
            assert np.array_equal(ids[sl][:, : order.shape[1]], order)
            assert np.array_equal(lens[sl][:, : order.shape[1]], np.take_along_axis(red, order, axis=1))
            m = max(red[np.isfinite(red)].max() if K > 1 else 0.0, 2 * bl.max() if has_b and K else 0.0)
            assert wmax[t] == m
            if active[t]:
                yy = y[sl]
                for a_ in range(K):
                    for b_ in range(a_ + 1, K):
                        fixed = int(red[a_, b_] * scale[t] + 0.5)
                        zc = int(ztop[sl][a_]) if top[sl][a_] >= 0 and top[sl][a_] == top[sl][b_] else 0
                        if int(yy[a_]) + int(yy[b_]) + zc + ((4 * fixed) << int(wshift[t])) < 0:
                            exp.append((t, a_, b_, float(red[a_, b_])))
            for _ in range(5):
                if K >= 2:
                    a_, b_ = rng.choice(K, 2, replace=False)
                    qt.append(t), qa.append(a_), qb.append(b_)
        assert got == sorted(exp) and len(exp) > 0
        out = eng.pairs_at(ci, qt, qa, qb)
        assert np.array_equal(eng.pairs_of_trial(ci, 1, int(cb.odd_count[1])), cb.pairs(1))
        for i in range(len(qt)):
            K = int(cb.odd_count[qt[i]])
            x, z = min(qa[i], qb[i]), max(qa[i], qb[i])
            assert out[i] == cb.pairs(qt[i])[tri_index(x, z, K)]
    with pytest.raises(_lib.FpbError):
        eng.pairs_at(0, [0], [0], [10**6])
    eng.close()


@pytest.mark.parametrize("d,ec,b,delta", [(9, "primal", "open", 0.09), (9, "both", "periodic", 0.1),
                                          (5, "dual", "open", 0.12), (13, "primal", "periodic", 0.09)])
def test_sparse_handoff_gives_the_full_fetch_optimum(d, ec, b, delta):
    """Per trial: the matching found without fetching the triangle weighs exactly what the matching on
    the complete graph weighs."""
    lat = build_lattice(d, ec, b)
    wo = {"method": "blueprint", "delta": delta}
    eng = TrialEngine(lat, delta, 0.0, wo, mode="fresh")
    n = 24 if d < 13 else 6
    batch = eng.run_rng(17, 0, n)
    for ci, e in enumerate(lat.ec):
        cb = batch.complexes[e]
        has_b = lat.complexes[e].has_boundary
        bl_all = cb.bnd_len if has_b else None
        st = {}
        tr, src, dst = mwpm_native.solve_batch_sparse(eng, ci, cb.odd_count, bl_all, stats=st)
        ref = mwpm_native.solve_batch(cb.odd_count, cb.pair_len, bl_all)
        for t in range(n):
            K = int(cb.odd_count[t])
            if K == 0:
                continue
            pl = cb.pairs(t)
            bl = cb.boundary(t)[0] if has_b else None
            got = list(zip(src[tr == t].tolist(), dst[tr == t].tolist()))
            exp = list(zip(ref[1][ref[0] == t].tolist(), ref[2][ref[0] == t].tolist()))
            assert total_weight(K, got, pl, bl) == pytest.approx(total_weight(K, exp, pl, bl), rel=1e-12, abs=1e-9)
        assert st["fallback_trials"] <= n // 4
        print(d, ec, b, e, "K mean", cb.odd_count.mean(), st)
    eng.close()


@pytest.mark.parametrize("d,ec,b,delta,p_swap", [(7, "primal", "open", 0.1, 0.2), (6, "both", "periodic", 0.105, 0.0)])
def test_ec_monte_carlo_same_failures_with_either_handoff(d, ec, b, delta, p_swap):
    code = SurfaceCode(d, ec, b)
    noise = CVLayer(code, delta=delta, p_swap=p_swap)
    args = {"weight_opts": {"method": "blueprint", "delta": delta}}
    f_full, s_full = ec_monte_carlo(600, code, noise, "MWPM", args, seed=9, batch=200, mode="fresh", handoff="full",
                                    return_stats=True)
    f_sparse, s_sparse = ec_monte_carlo(600, code, noise, "MWPM", args, seed=9, batch=200, mode="fresh",
                                        handoff="sparse", return_stats=True)
    assert s_sparse["handoff"] == "sparse" and s_full["handoff"] == "full"
    assert f_full == f_sparse and f_full > 0
