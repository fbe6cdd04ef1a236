"""Parity of the CUDA path (through the C ABI) against the oracle and the
committed reference vectors.  Bars (BASELINE.json north_star): bit-exact
homodyne values, bits, parities and odd sets; weights and path lengths within
1e-9 relative."""

import os

import numpy as np
import pytest

from flamingpy_b200.lattice import build_lattice
from oracle import restate
from tests import golden_util as gu

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _engine(lat, delta, p_swap, wopts, **kw):
    from flamingpy_b200.engine import TrialEngine

    return TrialEngine(lat, delta, p_swap, wopts, **kw)


def _compare_with_oracle(lat, batch, states, xs, delta, wopts, trials=None):
    T = len(xs)
    for t in (range(T) if trials is None else trials):
        res = restate.trial_from_draws(lat, states[t], xs[t], delta, wopts, want_graph=True)
        off = 0
        for e in lat.ec:
            ct = lat.complexes[e]
            m = res.per_complex[e]
            sl = slice(off, off + ct.Q)
            assert np.array_equal(batch.hom[t, sl], m["hom"]), (t, e)
            assert np.array_equal(batch.bits[t, sl], m["bits"]), (t, e)
            np.testing.assert_allclose(batch.weights[t, sl], m["weights"], rtol=RTOL, atol=0)
            cb = batch.complexes[e]
            assert np.array_equal(cb.parity[t], m["parity"])
            g = m["graph"]
            assert np.array_equal(cb.odd(t), g.odd)
            np.testing.assert_allclose(cb.pairs(t), g.pair_len, rtol=RTOL, atol=0)
            if ct.has_boundary:
                bl, bs = cb.boundary(t)
                np.testing.assert_allclose(bl, g.bnd_len, rtol=RTOL, atol=0)
                assert np.array_equal(bs, g.bnd_side)
            off += ct.Q


@pytest.mark.parametrize("path", gu.golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_golden_vectors(path):
    g = gu.load(path)
    lat = build_lattice(g["distance"], g["ec"], g["boundaries"])
    eng = _engine(lat, g["delta"], g["p_swap"], g["weight_opts"])
    batch = eng.run_injected(g["x"], g["state"])
    off = 0
    for e in lat.ec:
        ct = lat.complexes[e]
        sl = slice(off, off + ct.Q)
        assert np.array_equal(batch.hom[:, sl], g[f"{e}_hom"])
        assert np.array_equal(batch.bits[:, sl], g[f"{e}_bits"])
        np.testing.assert_allclose(batch.weights[:, sl], g[f"{e}_weights"], rtol=RTOL, atol=0)
        cb = batch.complexes[e]
        assert np.array_equal(cb.parity, g[f"{e}_parity"])
        for t in range(g["trials"]):
            assert np.array_equal(cb.odd(t), gu.ragged(g, f"{e}_odd", t))
            np.testing.assert_allclose(cb.pairs(t), gu.ragged(g, f"{e}_pair_len", t), rtol=RTOL, atol=0)
            if ct.has_boundary:
                bl, bs = cb.boundary(t)
                np.testing.assert_allclose(bl, gu.ragged(g, f"{e}_bnd_len", t), rtol=RTOL, atol=0)
                assert np.array_equal(bs, gu.ragged(g, f"{e}_bnd_side", t))
        off += ct.Q
    eng.close()


SEEDED = [
    (5, "primal", "open", 0.09, 0.25, {"method": "blueprint", "delta": 0.09}, 6),
    (6, "dual", "open", 0.1, 0.5, {"method": "blueprint", "delta": 0.1}, 4),
    (5, "both", "periodic", 0.09, 0, {"method": "blueprint", "delta": 0.09}, 4),
    (6, "primal", "toric", 0.08, 0.1, {"method": "blueprint", "delta": 0.08}, 3),
    ((3, 5, 4), "dual", "periodic", 0.11, 0.3, {"method": "blueprint", "delta": 0.11}, 4),
    (9, "primal", "open", 0.08, 0.5, {"method": "blueprint", "delta": 0.08}, 2),
    (7, "primal", "open", 0.09, 1, {"method": "uniform"}, 2),
    (5, "primal", "open", 0.06, 0.2, {"method": "blueprint", "delta": 0.06, "integer": True, "multiplier": 100}, 3),
]


@pytest.mark.parametrize("d,ec,b,delta,p_swap,wopts,trials", SEEDED)
def test_seeded_injected_vs_oracle(d, ec, b, delta, p_swap, wopts, trials):
    lat = build_lattice(d, ec, b)
    chain = restate.LabelChain(lat.N, p_swap)
    rng = np.random.default_rng(2024)
    states, xs = [], []
    for _ in range(trials):
        st = chain.next(rng)
        states.append(st)
        xs.append(restate.draw_quadratures(rng, restate.sigma_from_state(st, delta)))
    eng = _engine(lat, delta, p_swap, wopts)
    batch = eng.run_injected(np.stack(xs), np.stack(states))
    _compare_with_oracle(lat, batch, states, xs, delta, wopts)
    eng.close()


@pytest.mark.parametrize("step", [1.0, 3.0, 8.0])
def test_bucket_width_does_not_change_lengths(step):
    lat = build_lattice(7, "primal", "open")
    delta, p_swap = 0.1, 0.3
    wopts = {"method": "blueprint", "delta": delta}
    ref = _engine(lat, delta, p_swap, wopts, delta_step=1.0, mode="fresh")
    a = ref.run_rng(5, 0, 16)
    eng = _engine(lat, delta, p_swap, wopts, delta_step=step, mode="fresh")
    b = eng.run_rng(5, 0, 16)
    ca, cb = a.complexes["primal"], b.complexes["primal"]
    assert np.array_equal(ca.odd_ids, cb.odd_ids)
    np.testing.assert_allclose(cb.pair_len, ca.pair_len, rtol=1e-12, atol=0)
    np.testing.assert_allclose(cb.bnd_len, ca.bnd_len, rtol=1e-12, atol=0)
    assert np.array_equal(ca.bnd_side, cb.bnd_side)


@pytest.mark.parametrize("d,ec,b,p_swap", [(5, "primal", "open", 0.25), (4, "both", "periodic", 0.0),
                                          (9, "primal", "open", 0.5)])
def test_device_rng_pipeline_vs_oracle(d, ec, b, p_swap):
    """Device-generated draws dumped and replayed through the oracle."""
    lat = build_lattice(d, ec, b)
    delta = 0.09
    wopts = {"method": "blueprint", "delta": delta}
    eng = _engine(lat, delta, p_swap, wopts, mode="compat")
    T = 6 if d < 9 else 2
    batch = eng.run_rng(99, 0, T, want_draws=True)
    _compare_with_oracle(lat, batch, batch.state, batch.x, delta, wopts)


def test_device_rng_is_counter_based():
    """Results do not depend on how the trial range is batched."""
    lat = build_lattice(5, "primal", "open")
    eng = _engine(lat, 0.09, 0.25, {"method": "blueprint", "delta": 0.09})
    whole = eng.run_rng(7, 0, 12, want_draws=True)
    part = eng.run_rng(7, 8, 4, want_draws=True)
    assert np.array_equal(whole.x[8:], part.x)
    assert np.array_equal(whole.state[8:], part.state)
    assert np.array_equal(whole.hom[8:], part.hom)
    cw, cp = whole.complexes["primal"], part.complexes["primal"]
    assert np.array_equal(cw.odd_ids[cw.odd_off[8]:], cp.odd_ids)
    assert np.array_equal(cw.pair_len[cw.pair_off[8]:], cp.pair_len)


def test_device_rng_compat_recurrence_and_statistics():
    """Labels: p fraction ~ p_swap; compat mode follows GKP_t = V \\ (p_t U GKP_{t-1});
    draws: unit-variance normals scaled by the float32-rounded sigmas."""
    lat = build_lattice(7, "primal", "open")
    delta, ps = 0.09, 0.3
    T = 24
    fresh = _engine(lat, delta, ps, None, mode="fresh").run_rng(3, 0, T, stage=1, want_draws=True)
    compat = _engine(lat, delta, ps, None, mode="compat").run_rng(3, 0, T, stage=1, want_draws=True)
    p_mask = fresh.state == 2
    assert np.array_equal(p_mask, compat.state == 2)
    assert abs(p_mask.mean() - ps) < 4 * np.sqrt(ps * (1 - ps) / p_mask.size)
    prev = np.zeros(lat.N, bool)
    for t in range(T):
        gkp = ~p_mask[t] & ~prev
        expect = np.where(p_mask[t], 2, np.where(gkp, 1, 0))
        assert np.array_equal(compat.state[t], expect), t
        prev = gkp
    # chains restart every chain_len trials
    ch = _engine(lat, delta, ps, None, mode="compat", chain_len=5).run_rng(3, 0, 10, stage=1, want_draws=True)
    assert np.array_equal(ch.state[5] != 0, np.ones(lat.N, bool))
    # draws
    N = lat.N
    sig = restate.sigma_from_state(fresh.state[0], delta).astype(np.float64)
    z = np.concatenate([fresh.x[t] / restate.sigma_from_state(fresh.state[t], delta) for t in range(T)])
    assert abs(z.mean()) < 5 / np.sqrt(z.size)
    assert abs(z.var() - 1) < 5 * np.sqrt(2 / z.size)
    assert abs(((z ** 4).mean()) - 3) < 0.15
    assert np.all(fresh.x[0][sig == 0] == 0)
    # p_swap = 0 in compat mode alternates noisy / silent trials
    alt = _engine(lat, delta, 0, None, mode="compat").run_rng(3, 0, 4, stage=1, want_draws=True)
    assert [int(s.any()) for s in alt.state] == [1, 0, 1, 0]
    assert not alt.x[1].any()


def _toggle_case(p_swap):
    lat = build_lattice(5, "primal", "open")
    ct = lat.complexes["primal"]
    delta = 0.09
    wopts = {"method": "blueprint", "delta": delta}
    eng = _engine(lat, delta, p_swap, wopts, mode="fresh")
    T = 8
    batch = eng.run_rng(11, 0, T, want_draws=True)
    return lat, ct, delta, wopts, eng, T, batch


def test_paths_toggle_matches_oracle_decode():
    """Recovery through the device path walk gives the oracle's corrected bits (p_swap = 0:
    continuous weights, so every shortest path is unique)."""
    lat, ct, delta, wopts, eng, T, batch = _toggle_case(0)
    trial, cx, src, dst = [], [], [], []
    expect = np.zeros((T, ct.Q), np.uint8)
    for t in range(T):
        res = restate.trial_from_draws(lat, batch.state[t], batch.x[t], delta, wopts, want_graph=True, want_paths=True)
        md = res.per_complex["primal"]["graph"]
        K = len(md.odd)
        for u, v in restate.mwpm_pairs(md, True):
            u, v = (u, v) if u < v else (v, u)
            if u >= K:
                continue
            trial.append(t); cx.append(0); src.append(u); dst.append(v if v < K else -1)
            path = md.pair_paths[(u, v)] if v < K else md.bnd_paths[u]
            for q in path:
                expect[t, q] ^= 1
    flips = eng.paths_toggle(trial, cx, src, dst)
    assert np.array_equal(flips, expect)


def test_paths_toggle_is_a_shortest_path_with_ties():
    """With p-squeezed neighbours many weights are the constants -ln{1/4, 1/3, 2/5}
    (decoder.py:86) and shortest paths tie; any of them is a valid correction.  Check every
    pair and connector query: the toggled qubits flip exactly the end cubes and their weights
    add up to the matching-graph length."""
    lat, ct, delta, wopts, eng, T, batch = _toggle_case(0.25)
    cb = batch.complexes["primal"]
    w = batch.weights
    for t in range(3):
        K = int(cb.odd_count[t])
        odd = cb.odd(t)
        pl = cb.pairs(t)
        bl, _ = cb.boundary(t)
        qs = [(u, v) for u in range(K) for v in range(u + 1, K)] + [(u, -1) for u in range(K)]
        for u, v in qs:
            fl = eng.paths_toggle([t], [0], [u], [v])[t]
            length = pl[u * K - u * (u + 1) // 2 + (v - u - 1)] if v >= 0 else bl[u]
            np.testing.assert_allclose(w[t][fl.astype(bool)].sum(), length, rtol=1e-9)
            par = restate.stabilizer_parity(ct, fl)
            ends = sorted([int(odd[u])] + ([int(odd[v])] if v >= 0 else []))
            assert np.nonzero(par)[0].tolist() == ends


def test_d13_properties():
    """Full-size headline lattice: size-independent properties."""
    lat = build_lattice(13, "primal", "periodic")
    ct = lat.complexes["primal"]
    delta = 0.09
    wopts = {"method": "blueprint", "delta": delta}
    eng = _engine(lat, delta, 0, wopts, mode="fresh")
    T = 4
    batch = eng.run_rng(1, 0, T, want_draws=True)
    cb = batch.complexes["primal"]
    assert np.all(cb.odd_count % 2 == 0)  # periodic lattice: even number of odd cubes
    for t in range(T):
        # oracle for everything but the graph (cheap), bit-exact
        res = restate.trial_from_draws(lat, batch.state[t], batch.x[t], delta, wopts, want_graph=False)
        m = res.per_complex["primal"]
        assert np.array_equal(batch.hom[t], m["hom"])
        assert np.array_equal(batch.bits[t], m["bits"])
        assert np.array_equal(cb.parity[t], m["parity"])
        np.testing.assert_allclose(batch.weights[t], m["weights"], rtol=RTOL, atol=0)
        # a few rows of the pair matrix against the oracle's Dijkstra
        odd = cb.odd(t)
        K = len(odd)
        pl = cb.pairs(t)
        assert len(pl) == K * (K - 1) // 2
        for a in (0, K // 2, K - 2):
            dist, _, _ = restate._dijkstra(ct, m["weights"], [(int(odd[a]), 0.0, -1)], False)
            row = a * K - a * (a + 1) // 2
            np.testing.assert_allclose(pl[row:row + K - 1 - a], dist[odd[a + 1:]], rtol=RTOL, atol=0)
        # triangle inequality on a sample
        def d_(i, j):
            i, j = (i, j) if i < j else (j, i)
            return pl[i * K - i * (i + 1) // 2 + (j - i - 1)]
        rng = np.random.default_rng(t)
        for _ in range(200):
            i, j, k = rng.choice(K, 3, replace=False)
            assert d_(i, j) <= d_(i, k) + d_(k, j) + 1e-9
        assert pl.min() >= m["weights"].min() - 1e-12
