"""The sparse hand-off of the host matcher (``mwpm_native.solve_batch_sparse``, ``fpb_mwpm_sparse_*``)
on CPU: a NumPy stand-in answers the three device calls (``knn_len`` / ``pairs_at`` / ``price``) from
full triangles, so the resumable solver state machine can be checked against the dense optimum
without a GPU.  The same driver runs against the real kernels in tests/test_gpu_handoff.py."""

import numpy as np
import pytest

from flamingpy_b200 import mwpm_native
from tests.test_mwpm_native import _random_graph, total_weight


class DeviceStandIn:
    """What ``TrialEngine`` offers ``solve_batch_sparse``, computed from host triangles exactly as
    the kernels define it (k_knn, k_pairs_at, k_price in csrc/fpb_kernels.cuh)."""

    def __init__(self, Ks, pairs, bnds):
        self.Ks, self.pairs, self.bnds = Ks, pairs, bnds
        self.calls = {"knn_len": 0, "pairs_at": 0, "price": 0, "priced_pairs": 0}

    def _reduced(self, t):
        K, pl, bl = self.Ks[t], self.pairs[t], self.bnds[t]
        d = np.full((K, K), np.inf)
        iu = np.triu_indices(K, 1)
        d[iu] = pl if bl is None else np.minimum(pl, bl[iu[0]] + bl[iu[1]])
        return np.minimum(d, d.T)

    def knn_len(self, ci, k):
        self.calls["knn_len"] += 1
        ids, lens, wmax = [], [], []
        for t, K in enumerate(self.Ks):
            if K == 0:
                wmax.append(0.0)
                continue
            d = self._reduced(t)
            order = np.argsort(d, axis=1, kind="stable")[:, :k]
            ln = np.take_along_axis(d, order, axis=1)
            if order.shape[1] < k:
                order = np.pad(order, ((0, 0), (0, k - order.shape[1])), constant_values=-1)
                ln = np.pad(ln, ((0, 0), (0, k - ln.shape[1])), constant_values=np.inf)
            order = np.where(np.isinf(ln), -1, order)
            ids.append(order.astype(np.int32))
            lens.append(ln)
            finite = d[np.isfinite(d)]
            m = finite.max() if finite.size else 0.0
            if self.bnds[t] is not None:
                m = max(m, 2 * self.bnds[t].max())
            wmax.append(m)
        return (np.ascontiguousarray(np.concatenate(ids)), np.ascontiguousarray(np.concatenate(lens)),
                np.array(wmax, np.float64))

    def pairs_at(self, ci, trial, a, b):
        self.calls["pairs_at"] += 1
        out = np.empty(len(trial))
        for i, (t, x, y) in enumerate(zip(trial, a, b)):
            K = self.Ks[t]
            x, y = min(x, y), max(x, y)
            out[i] = self.pairs[t][x * K - x * (x + 1) // 2 + (y - x - 1)]
        return out

    def pairs_of_trial(self, ci, t, K):
        return np.array(self.pairs[t], np.float64)

    def price(self, ci, y, scale, wshift, active, cap=0, top=None, ztop=None):
        self.calls["price"] += 1
        vt, va, vb, vl = [], [], [], []
        off = 0
        for t, K in enumerate(self.Ks):
            if active[t] and K > 1:
                d = self._reduced(t)
                yy = y[off:off + K]
                iu = np.triu_indices(K, 1)
                fixed = (d[iu] * scale[t] + 0.5).astype(np.int64)  # to_fixed of csrc/mwpm.cpp
                rc = yy[iu[0]] + yy[iu[1]] + ((4 * fixed) << int(wshift[t]))
                if top is not None:
                    tt, zz = top[off:off + K], ztop[off:off + K]
                    rc = rc + np.where((tt[iu[0]] >= 0) & (tt[iu[0]] == tt[iu[1]]), zz[iu[0]], 0)
                hit = np.nonzero(rc < 0)[0]
                self.calls["priced_pairs"] += len(hit)
                vt += [t] * len(hit)
                va += iu[0][hit].tolist()
                vb += iu[1][hit].tolist()
                vl += d[iu][hit].tolist()
            off += K
        return (np.array(vt, np.int32), np.array(va, np.int32), np.array(vb, np.int32), np.array(vl, np.float64))


def _check(Ks, kinds, boundary, seed):
    rng = np.random.default_rng(seed)
    graphs = [_random_graph(rng, k, kind) if k else (np.empty(0), np.empty(0)) for k, kind in zip(Ks, kinds)]
    pairs = [g[0] for g in graphs]
    bnds = [g[1] if boundary else None for g in graphs]
    dev = DeviceStandIn(Ks, pairs, bnds)
    bl = np.concatenate([g[1] for g in graphs]) if boundary else None
    st = {}
    tr, src, dst = mwpm_native.solve_batch_sparse(dev, 0, np.array(Ks, np.int32), bl, threads=2, stats=st)
    ref = mwpm_native.solve_batch(np.array(Ks, np.int32), np.concatenate(pairs), bl, 2)
    for t, K in enumerate(Ks):
        if not K:
            assert not (tr == t).any()
            continue
        got = list(zip(src[tr == t].tolist(), dst[tr == t].tolist()))
        exp = list(zip(ref[1][ref[0] == t].tolist(), ref[2][ref[0] == t].tolist()))
        b = graphs[t][1]
        assert total_weight(K, got, pairs[t], b) == pytest.approx(total_weight(K, exp, pairs[t], b), rel=1e-12, abs=1e-9), t
    return dev, st


@pytest.mark.parametrize("boundary", [False, True])
def test_sparse_handoff_finds_the_dense_optimum(boundary):
    Ks = [60, 0, 76, 52, 131, 4, 48, 90] if boundary else [60, 0, 76, 52, 130, 4, 48, 90]
    for seed, kinds in ((1, [1] * 8), (2, [2] * 8), (3, [3, 1, 0, 2, 3, 1, 0, 3]), (4, [0] * 8)):
        dev, st = _check(Ks, kinds, boundary, seed)
        assert dev.calls["knn_len"] == 1 and dev.calls["price"] >= 1
        assert st["fallback_trials"] <= 4


def test_sparse_handoff_small_and_odd_graphs():
    """Graphs below the sparse threshold ask for every pair and are solved densely; odd K with a
    boundary uses the dummy vertex; trivial trials stay empty."""
    dev, st = _check([1, 2, 3, 0, 7, 47, 49, 51], [1] * 8, True, 11)
    assert st["requested_pairs"] >= 47 * 46 // 2
    _check([2, 4, 0, 10, 46], [2] * 5, False, 12)


@pytest.mark.parametrize("boundary", [False, True])
def test_pricing_flags_a_small_part_of_the_triangle_on_metric_graphs(boundary):
    """On metric graphs (what the lattice produces) the candidate graph is almost optimal: over all
    rounds the pricing pass returns a few percent of the triangle, and a few rounds suffice."""
    Ks = [200, 260, 180]
    dev, st = _check(Ks, [1, 1, 1], boundary, 21)
    total_pairs = sum(k * (k - 1) // 2 for k in Ks)
    assert dev.calls["priced_pairs"] < (0.10 if boundary else 0.02) * total_pairs
    assert st["pricing_rounds"] <= 8 and st["fallback_trials"] <= 1
