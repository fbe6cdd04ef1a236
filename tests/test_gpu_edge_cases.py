"""Edge cases of the C ABI path: empty syndromes, single odd cubes, tiny batches, ragged
batches mixing silent and noisy trials, argument errors."""

import numpy as np
import pytest

from flamingpy_b200 import _lib
from flamingpy_b200.lattice import build_lattice
from oracle import c_oracle, restate

pytestmark = pytest.mark.gpu


def _engine(lat, delta, p_swap, wopts=None, **kw):
    from flamingpy_b200.engine import TrialEngine

    return TrialEngine(lat, delta, p_swap, wopts or {"method": "blueprint", "delta": delta}, **kw)


def _same_as_oracle(lat, batch, ref):
    assert np.array_equal(batch.hom, ref.hom)
    assert np.array_equal(batch.bits, ref.bits)
    np.testing.assert_allclose(batch.weights, ref.weights, rtol=1e-9, atol=0)
    for e in lat.ec:
        a, b = batch.complexes[e], ref.complexes[e]
        assert np.array_equal(a.parity, b.parity)
        assert np.array_equal(a.odd_count, b.odd_count)
        assert np.array_equal(a.odd_ids, b.odd_ids)
        np.testing.assert_allclose(a.pair_len, b.pair_len, rtol=1e-9, atol=0)
        np.testing.assert_allclose(a.bnd_len, b.bnd_len, rtol=1e-9, atol=0)
        assert np.array_equal(a.bnd_side, b.bnd_side)


def test_all_silent_batch_has_empty_graph():
    lat = build_lattice(5, "primal", "open")
    eng = _engine(lat, 0.09, 0)
    N = lat.N
    batch = eng.run_injected(np.zeros((3, 2 * N)), np.zeros((3, N), np.uint8))
    cb = batch.complexes["primal"]
    assert not batch.bits.any() and not cb.parity.any()
    assert cb.odd_count.tolist() == [0, 0, 0]
    assert cb.odd_ids.size == 0 and cb.pair_len.size == 0 and cb.bnd_len.size == 0
    assert eng.timings()["graph_ctas"] == 0
    assert eng.paths_toggle([], [], [], []).sum() == 0


def test_single_trial_and_single_odd_cube():
    """One flipped boundary qubit: K = 1, no pairs, one connector edge."""
    lat = build_lattice(4, "primal", "open")
    ct = lat.complexes["primal"]
    N = lat.N
    x = np.zeros((1, 2 * N))
    q = int(ct.low_points[0])
    x[0, N + q] = np.sqrt(np.pi)  # p homodyne one bin over -> bit 1 on a LOW boundary qubit
    st = np.ones((1, N), np.uint8)
    wo = {"method": "blueprint", "delta": 0.09}
    eng = _engine(lat, 0.09, 0, wo)
    batch = eng.run_injected(x, st)
    cb = batch.complexes["primal"]
    assert cb.odd_count.tolist() == [1] and cb.pair_len.size == 0
    ref = c_oracle.run(lat, 0.09, 0, wo, x, st)
    _same_as_oracle(lat, batch, ref)
    assert cb.bnd_side.tolist() == [0]
    flips = eng.paths_toggle([0], [0], [0], [-1])
    local = int(np.nonzero(ct.qubits == q)[0][0])
    assert np.nonzero(flips[0])[0].tolist() == [local]


def test_ragged_batch_mixing_silent_and_noisy_trials_full_size():
    """Compat chain at d=13 (silent / noisy alternate) against the C oracle, all outputs."""
    lat = build_lattice(13, "primal", "open")
    delta = 0.09
    wo = {"method": "blueprint", "delta": delta}
    chain = restate.LabelChain(lat.N, 0)
    rng = np.random.default_rng(77)
    st = np.stack([chain.next(rng) for _ in range(4)])
    x = np.stack([restate.draw_quadratures(rng, restate.sigma_from_state(s, delta)) for s in st])
    eng = _engine(lat, delta, 0, wo)
    batch = eng.run_injected(x, st)
    ref = c_oracle.run(lat, delta, 0, wo, x, st)
    _same_as_oracle(lat, batch, ref)
    assert batch.complexes["primal"].odd_count[1] == 0 and batch.complexes["primal"].odd_count[0] > 100


def test_both_complexes_full_size_against_c_oracle():
    """BASELINE configs[2] lattice (d=13 periodic, ec=both), two noisy trials, every pair length."""
    lat = build_lattice(13, "both", "periodic")
    delta = 0.09
    wo = {"method": "blueprint", "delta": delta}
    eng = _engine(lat, delta, 0, wo, mode="fresh")
    batch = eng.run_rng(5, 0, 2, want_draws=True)
    ref = c_oracle.run(lat, delta, 0, wo, batch.x, batch.state)
    _same_as_oracle(lat, batch, ref)
    assert batch.complexes["primal"].pair_len.size > 200000


def test_headline_workload_compat_chain_full_size_against_c_oracle():
    """The bench's default workload as it runs there: d=13 periodic, ec=both, compat chain (noisy / silent
    trials alternate), device RNG - every output of four consecutive trials against the C oracle."""
    lat = build_lattice(13, "both", "periodic")
    delta = 0.09
    wo = {"method": "blueprint", "delta": delta}
    eng = _engine(lat, delta, 0, wo, mode="compat")
    batch = eng.run_rng(20261019, 2048, 4, want_draws=True)
    ref = c_oracle.run(lat, delta, 0, wo, batch.x, batch.state)
    _same_as_oracle(lat, batch, ref)
    k = batch.complexes["primal"].odd_count
    assert k[0] > 400 and k[1] == 0 and k[2] > 400 and k[3] == 0


def test_config4_lattice_d21_all_p():
    """d=21 open, p_swap=1 (configs[3]): too large for the shared-memory weight copy, so the
    graph stage runs the variant that keeps weights and face tables in global memory."""
    lat = build_lattice(21, "primal", "open")
    delta = 0.09
    wo = {"method": "blueprint", "delta": delta}
    eng = _engine(lat, delta, 1, wo)
    batch = eng.run_rng(1, 0, 1, want_draws=True)
    assert (batch.state == 2).all()
    ref = c_oracle.run(lat, delta, 1, wo, batch.x, batch.state)
    _same_as_oracle(lat, batch, ref)
    # all-p lattice: weights are the constants -ln{1/4, 1/3, 2/5} (decoder.py:86)
    assert set(np.round(np.unique(batch.weights), 12)) <= set(np.round(-np.log([0.25, 1 / 3, 0.4]), 12))
    K = int(batch.complexes["primal"].odd_count[0])
    assert K > 3000
    # a correction path on the big lattice: flips the two end cubes
    ct = lat.complexes["primal"]
    fl = eng.paths_toggle([0, 0], [0, 0], [0, 5], [3, -1])[0]
    par = restate.stabilizer_parity(ct, fl)
    odd = batch.complexes["primal"].odd(0)
    assert set(np.nonzero(par)[0].tolist()) == {int(odd[0]), int(odd[3]), int(odd[5])}


@pytest.mark.parametrize("d,ec,b,team_w", [(17, "both", "periodic", None), (19, "primal", "periodic", None),
                                           (19, "primal", "open", "0"), (21, "primal", "periodic", None),
                                           (21, "primal", "periodic", "16"), (25, "primal", "open", None)])
def test_sweep_lattices_d17_to_d25(d, ec, b, team_w, monkeypatch):
    """Threshold-sweep lattices of configs[4] and the sizes between them.  One to five searches
    fit an SM there, so the graph stage runs the W-warp team kernel (4 warps per search at
    d=17/19, 8 at d=21, 16 at d=25; FPB_TEAM_W overrides: "0" = the one-warp global-weights
    variant); one noisy trial each against the C oracle."""
    if team_w is not None:
        monkeypatch.setenv("FPB_TEAM_W", team_w)
    lat = build_lattice(d, ec, b)
    delta = 0.1
    wo = {"method": "blueprint", "delta": delta}
    eng = _engine(lat, delta, 0, wo, mode="fresh")
    batch = eng.run_rng(9, 0, 1, want_draws=True)
    ref = c_oracle.run(lat, delta, 0, wo, batch.x, batch.state)
    _same_as_oracle(lat, batch, ref)
    print(d, ec, b, "timings", eng.timings(), "K", [int(batch.complexes[e].odd_count[0]) for e in lat.ec])


def test_argument_errors_are_loud():
    lat = build_lattice(3, "primal", "open")
    eng = _engine(lat, 0.09, 0.25)
    with pytest.raises(_lib.FpbError):
        eng.run_rng(1, 0, 0)
    with pytest.raises(_lib.FpbError):
        eng.run_rng(1, -5, 4)
    with pytest.raises(ValueError):
        eng.run_injected(np.zeros((2, 7)), np.zeros((2, lat.N), np.uint8))
    eng.run_rng(1, 0, 4, stage=_lib.STAGE_NOISE)
    with pytest.raises(_lib.FpbError):
        eng.fetch(_lib.STAGE_GRAPH)  # graph outputs of a noise-only run
    with pytest.raises(_lib.FpbError):
        eng.paths_toggle([0], [0], [0], [1])
    b = eng.run_rng(1, 0, 4)
    K = int(b.complexes["primal"].odd_count[0])
    with pytest.raises(_lib.FpbError):
        eng.paths_toggle([0], [0], [K], [0])  # endpoint out of range
    from flamingpy_b200.engine import TrialEngine

    with pytest.raises(_lib.FpbError):
        TrialEngine(lat, -1.0, 0.25)
    with pytest.raises(_lib.FpbError):
        TrialEngine(lat, 0.09, 0.25, device=99)


def test_uniform_and_integer_weights_with_zero_weight_edges():
    """Integer weights with a tiny multiplier round to 0 on many edges: zero-weight edges must
    not break the bucketed search (fallback bucket width, label correcting)."""
    lat = build_lattice(5, "primal", "open")
    delta = 0.09
    wo = {"method": "blueprint", "delta": delta, "integer": True, "multiplier": 0.4}
    eng = _engine(lat, delta, 0.3, wo, mode="fresh")
    batch = eng.run_rng(3, 0, 6, want_draws=True)
    assert (batch.weights == 0).any()
    ref = c_oracle.run(lat, delta, 0.3, wo, batch.x, batch.state)
    assert np.array_equal(batch.weights, ref.weights)
    cb, rb = batch.complexes["primal"], ref.complexes["primal"]
    assert np.array_equal(cb.pair_len, rb.pair_len)
    assert np.array_equal(cb.bnd_len, rb.bnd_len)


@pytest.mark.parametrize("d,ec,b,wo", [
    (7, "primal", "open", {"method": "blueprint", "delta": 0.09}),
    (6, "dual", "open", {"method": "uniform"}),
    (5, "both", "periodic", {"method": "blueprint", "delta": 0.09, "integer": True, "multiplier": 100}),
])
def test_static_weight_table_is_bit_identical_to_search(d, ec, b, wo):
    """p_swap = 1 (or uniform weights): weights are trial-independent, the matching graph gathered
    from the precomputed all-pairs table equals the per-trial search bit for bit."""
    lat = build_lattice(d, ec, b)
    search = _engine(lat, 0.09, 1, wo)
    table = _engine(lat, 0.09, 1, wo)
    table.build_table()
    a = search.run_rng(4, 0, 24)
    t = table.run_rng(4, 0, 24)
    for e in lat.ec:
        ca, ct_ = a.complexes[e], t.complexes[e]
        assert np.array_equal(ca.odd_ids, ct_.odd_ids)
        assert np.array_equal(ca.pair_len, ct_.pair_len)
        assert np.array_equal(ca.bnd_len, ct_.bnd_len)
        assert np.array_equal(ca.bnd_side, ct_.bnd_side)
    assert a.complexes[lat.ec[0]].pair_len.size > 0
    # and against the oracle
    full = table.run_rng(4, 0, 3, want_draws=True)
    ref = c_oracle.run(lat, 0.09, 1, wo, full.x, full.state)
    _same_as_oracle(lat, full, ref)


def test_static_table_guards():
    lat = build_lattice(4, "primal", "open")
    with pytest.raises(_lib.FpbError, match="trial-dependent"):
        _engine(lat, 0.09, 0.5, {"method": "blueprint", "delta": 0.09}).build_table()
    eng = _engine(lat, 0.09, 1, {"method": "blueprint", "delta": 0.09})
    eng.build_table()
    N = lat.N
    x = np.random.default_rng(0).normal(0, 0.3, (2, 2 * N))
    eng.run_injected(x, np.full((2, N), 2, np.uint8))  # all p: fine
    with pytest.raises(_lib.FpbError, match="not all p"):
        eng.run_injected(x, np.ones((2, N), np.uint8))  # GKP labels: weights differ from the table


@pytest.mark.parametrize("delta", [1e-4, 1e-3, 0.3, 2.0, 10.0])
def test_weight_limits_at_extreme_delta(delta):
    """reference tests/cv/test_gkp.py:80-102: Z_err_cond -> 0 at tiny delta (weight -> -ln(float_min)
    where the series underflows) and -> 0.5 (weight ln 2) at huge delta; same values as the oracle."""
    lat = build_lattice(4, "primal", "open")
    wo = {"method": "blueprint", "delta": delta}
    eng = _engine(lat, delta, 0, wo, mode="fresh")
    batch = eng.run_rng(21, 0, 8, stage=_lib.STAGE_SYNDROME, want_draws=True)
    ref = c_oracle.run(lat, delta, 0, wo, batch.x, batch.state, stage=_lib.STAGE_SYNDROME)
    assert np.array_equal(batch.hom, ref.hom) and np.array_equal(batch.bits, ref.bits)
    w, wr = batch.weights, ref.weights
    normal = wr < 700  # numerator is a normal double there (see DESIGN.md on denormals)
    np.testing.assert_allclose(w[normal], wr[normal], rtol=1e-9, atol=0)
    # denormal numerators carry a few bits only, and at the very edge of underflow one exp() may
    # return the smallest denormal (weight ~744) where the other returns 0 (weight -ln(float_min) =
    # 708.4): not a 1e-9 quantity; both sides must agree that the weight is "huge"
    edge = (~normal) & ((w > 740) | (wr > 740))
    np.testing.assert_allclose(w[~normal & ~edge], wr[~normal & ~edge], rtol=0, atol=0.75)
    assert np.all(w[edge] > 700) and np.all(wr[edge] > 700)
    assert np.all(w >= np.log(2) - 1e-12)
    if delta >= 2.0:
        assert np.all(w < 2.0)  # error probability close to 1/2 everywhere
    if delta <= 1e-3:
        assert np.median(w) > 50


def test_engine_reuse_across_batch_sizes_and_handles():
    """Buffers grow and shrink logically across runs; two handles on one device are independent."""
    lat = build_lattice(5, "primal", "open")
    wo = {"method": "blueprint", "delta": 0.09}
    a = _engine(lat, 0.09, 0.25, wo, mode="fresh")
    b = _engine(lat, 0.09, 0.25, wo, mode="fresh")
    big = a.run_rng(3, 0, 300)
    small = a.run_rng(3, 100, 7)
    other = b.run_rng(3, 100, 7)
    again = a.run_rng(3, 0, 300)
    for x, y in ((small, other), (big, again)):
        assert np.array_equal(x.hom, y.hom)
        assert np.array_equal(x.complexes["primal"].pair_len, y.complexes["primal"].pair_len)
    cb, cs = big.complexes["primal"], small.complexes["primal"]
    assert np.array_equal(cb.pair_len[cb.pair_off[100]:cb.pair_off[107]], cs.pair_len)
    a.close()
    a.close()  # idempotent
    assert b.run_rng(3, 0, 2).n_trials == 2


def test_two_host_threads_two_engines():
    """The e2e arm of bench.py drives two engines from two threads; results are unaffected."""
    import threading

    lat = build_lattice(7, "primal", "open")
    wo = {"method": "blueprint", "delta": 0.09}
    engs = [_engine(lat, 0.09, 0, wo, mode="fresh") for _ in range(2)]
    ref = engs[0].run_rng(5, 0, 64).complexes["primal"].pair_len.copy()
    out = [None, None]

    def work(j):
        for _ in range(5):
            out[j] = engs[j].run_rng(5, 0, 64).complexes["primal"].pair_len

    ths = [threading.Thread(target=work, args=(j,)) for j in range(2)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert np.array_equal(out[0], ref) and np.array_equal(out[1], ref)


def test_device_tensors_are_zero_copy_views():
    """torch views over the handle's device buffers (no host round trip) match the fetched arrays."""
    import torch

    lat = build_lattice(5, "primal", "open")
    eng = _engine(lat, 0.09, 0.25, {"method": "blueprint", "delta": 0.09}, mode="fresh")
    batch = eng.run_rng(2, 0, 16)
    dv = eng.device_tensors()
    assert dv["hom"].is_cuda and dv["hom"].dtype == torch.float64
    assert np.array_equal(dv["hom"].cpu().numpy(), batch.hom)
    assert np.array_equal(dv["weights"].cpu().numpy(), batch.weights)
    cb = batch.complexes["primal"]
    assert np.array_equal(dv["primal"]["odd_ids"].cpu().numpy(), cb.odd_ids)
    assert np.array_equal(dv["primal"]["pair_len"].cpu().numpy(), cb.pair_len)
    assert np.array_equal(dv["primal"]["bnd_len"].cpu().numpy(), cb.bnd_len)
    # a device-side reduction without leaving the GPU
    assert float(dv["primal"]["pair_len"].min()) == float(cb.pair_len.min())


@pytest.mark.parametrize("d,ec,b,delta,p_swap", [(5, "primal", "open", 0.1, 0.25), (6, "both", "periodic", 0.09, 0.0),
                                                 (11, "primal", "open", 0.09, 0.0), (4, "dual", "toric", 0.12, 0.5)])
def test_rustworkx_compat_int_lengths(d, ec, b, delta, p_swap):
    """FPB_LENGTH_RX_TRUNC: the sum of int(weight) along the float-shortest path
    (codes/graphs/stabilizer_graph.py:486-500), against the restatement; sides chosen on the int lengths."""
    lat = build_lattice(d, ec, b)
    wo = {"method": "blueprint", "delta": delta}
    eng = _engine(lat, delta, p_swap, wo, mode="fresh", lengths="rustworkx")
    flt = _engine(lat, delta, p_swap, wo, mode="fresh")
    n = 3 if d > 6 else 6
    batch = eng.run_rng(4, 0, n, want_draws=True)
    ref = flt.run_rng(4, 0, n)
    for t in range(n):
        res = restate.trial_from_draws(lat, batch.state[t], batch.x[t], delta, wo, want_graph=False)
        for e in lat.ec:
            ct = lat.complexes[e]
            m = res.per_complex[e]
            g = restate.matching_graph_rx(ct, m["weights"], m["parity"])
            cb = batch.complexes[e]
            assert np.array_equal(cb.odd(t), g.odd)
            assert np.array_equal(cb.pairs(t), g.pair_len)  # integers: exact
            assert np.all(cb.pairs(t) <= ref.complexes[e].pairs(t) + 1e-9)  # truncation only lowers a path's weight
            if ct.has_boundary:
                bl, bs = cb.boundary(t)
                assert np.array_equal(bl, g.bnd_len) and np.array_equal(bs, g.bnd_side)
    # integer weights: both notions of length coincide
    wi = {"method": "blueprint", "delta": delta, "integer": True, "multiplier": 10}
    a = _engine(lat, delta, p_swap, wi, mode="fresh", lengths="rustworkx").run_rng(4, 0, 2)
    c = _engine(lat, delta, p_swap, wi, mode="fresh").run_rng(4, 0, 2)
    for e in lat.ec:
        assert np.array_equal(a.complexes[e].pair_len, c.complexes[e].pair_len)
