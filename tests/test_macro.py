"""CVMacroLayer (flamingpy/noise/cv.py:352-625): the macronode lattice tables, the host sampler's RNG call
order and the NumPy restatement against the live reference and the committed reference vectors (CPU),
and the CUDA kernels / facade against those vectors and the restatement (GPU).
Bars: reduced labels, bit values, parities and odd sets bit-exact; p_phase_cond, weights and lengths 1e-9."""

import os

import numpy as np
import pytest

from flamingpy_b200.lattice import build_lattice, macro_tables
from flamingpy_b200.noise import MacroSampler
from oracle import restate
from tests import golden_util as gu

RTOL = 1e-9


def _ids(p):
    return os.path.basename(p)[:-4]


# ------------------------------------------------------------------------------------------------
# CPU: tables, sampler, restatement
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,ec,b", [(3, "primal", "open"), (4, "both", "periodic"), (3, "dual", "toric"),
                                    ((2, 3, 4), "primal", "open"), (2, "primal", "periodic")])
def test_macro_tables_shape(d, ec, b):
    lat = build_lattice(d, ec, b)
    partner, sign = macro_tables(lat)
    N = lat.N
    assert partner.shape == (4 * N,) and sign.shape == (4 * N,)
    has = partner >= 0
    assert np.array_equal(partner[partner[has]], np.nonzero(has)[0])  # dumbbells pair up
    assert np.all(partner[has] // 4 != np.nonzero(has)[0] // 4)       # across macronodes
    assert np.array_equal(has.reshape(N, 4).sum(1), lat.degree)      # one micronode per lattice edge
    assert np.all(np.abs(sign[has]) == 1) and not sign[~has].any()
    # padding micronodes come last inside their macronode (egraph.py:85-95)
    h4 = has.reshape(N, 4)
    assert np.all(h4[:, :-1] >= h4[:, 1:])


@pytest.mark.reference
@pytest.mark.parametrize("d,ec,b,pol", [(3, "primal", "open", False), (3, "dual", "open", False),
                                        (4, "primal", "periodic", False), (3, "dual", "toric", False),
                                        ((2, 3, 4), "primal", "open", False), (2, "primal", "periodic", False),
                                        (3, "primal", "periodic", True)])
def test_macro_tables_match_reference_macronize(reference, d, ec, b, pol):
    """Micronode order, dumbbell partners and edge weights of EGraph.macronize (egraph.py:26-103,133-155)."""
    from flamingpy.codes.surface_code import alternating_polarity
    from flamingpy.noise import CVMacroLayer
    from oracle import ref_shim

    code = ref_shim.make_code(d, ec, b, polarity=alternating_polarity if pol else None)
    layer = CVMacroLayer(code, delta=0.1, p_swap=0.3)
    lat = build_lattice(d, ec, b, polarity=pol)
    partner, sign = macro_tables(lat)
    mg = layer.egraph
    assert len(mg) == 4 * lat.N
    idx = lat.coord_index()
    for i in range(len(mg)):
        pt = mg.to_points[i]
        assert idx[tuple(round(v) for v in pt)] == i // 4
        adj = list(mg[pt])
        if adj:
            assert partner[i] == mg.to_indices[adj[0]]
            assert sign[i] == mg.edges[pt, adj[0]].get("weight", 1)
        else:
            assert partner[i] == -1


@pytest.mark.reference
@pytest.mark.parametrize("d,ec,b,delta,ps", [(3, "primal", "open", 0.1, 0.3), (3, "dual", "open", 0.08, 0.0),
                                             (4, "primal", "periodic", 0.1, 0.2), (3, "dual", "toric", 0.12, 0.5),
                                             ((2, 3, 4), "primal", "open", 0.1, 1), (2, "primal", "periodic", 0.1, 0.25)])
def test_restatement_matches_live_reference(reference, d, ec, b, delta, ps):
    """Same NumPy seed -> the sampler replays the reference's draws, and the restatement reproduces its
    reduced lattice, trial after trial on one reused layer (carry-over included)."""
    from flamingpy.decoders.decoder import assign_weights
    from flamingpy.noise import CVMacroLayer
    from oracle import ref_shim

    code = ref_shim.make_code(d, ec, b)
    layer = CVMacroLayer(code, delta=delta, p_swap=ps)
    lat = build_lattice(d, ec, b)
    partner, sign = macro_tables(lat)
    samp = MacroSampler(4 * lat.N, ps)
    r1, r2 = np.random.default_rng(77), np.random.default_rng(77)
    pts = [tuple(int(v) for v in c) for c in lat.coords]
    wo = {"method": "blueprint", "prob_precomputed": True}
    for t in range(4):
        layer.apply_noise(r1)
        st, mean, z = samp.next(r2)
        red = restate.macro_trial_from_draws(lat, partner, sign, st, mean, z, delta)
        nodes = code.graph.nodes
        assert np.array_equal([int(nodes[p]["bit_val"]) for p in pts], red["bit_val"]), t
        assert np.array_equal([2 if nodes[p]["state"] == "p" else 1 for p in pts], red["red_state"])
        np.testing.assert_allclose(red["p_phase_cond"], [nodes[p]["p_phase_cond"] for p in pts], rtol=RTOL, atol=1e-300)
        assign_weights(code, "MWPM", **wo)
        q = np.concatenate([lat.complexes[e].qubits for e in lat.ec])
        mw = restate.macro_weights(lat, red["red_state"], red["p_phase_cond"], wo)
        np.testing.assert_allclose(mw[q], [nodes[pts[i]]["weight"] for i in q], rtol=RTOL)


@pytest.mark.parametrize("path", gu.macro_files(), ids=_ids)
def test_restatement_against_golden(path):
    g = gu.load_macro(path)
    lat = build_lattice(g["distance"], g["ec"], g["boundaries"])
    partner, sign = macro_tables(lat)
    # the committed inputs are what the sampler draws from the recorded seed
    samp = MacroSampler(4 * lat.N, g["p_swap"])
    rng = np.random.default_rng(int(g["seed"]))
    for t in range(g["trials"]):
        st, mean, z = samp.next(rng)
        assert np.array_equal(st, g["state"][t]) and np.array_equal(mean, g["mean"][t]) and np.array_equal(z, g["z"][t])
        res = restate.macro_trial(lat, partner, sign, st, mean, z, g["delta"], g["weight_opts"])
        assert np.array_equal(res.macro["red_state"], g["red_state"][t])
        assert np.array_equal(res.macro["bit_val"], g["bit_val"][t])
        np.testing.assert_allclose(res.macro["p_phase_cond"], g["p_phase_cond"][t], rtol=RTOL, atol=1e-300)
        for e in lat.ec:
            m = res.per_complex[e]
            np.testing.assert_allclose(m["weights"], g[f"{e}_weights"][t], rtol=RTOL)
            assert np.array_equal(m["parity"], g[f"{e}_parity"][t])
            assert np.array_equal(m["graph"].odd, gu.ragged(g, f"{e}_odd", t))
            np.testing.assert_allclose(m["graph"].pair_len, gu.ragged(g, f"{e}_pair_len", t), rtol=RTOL)
            if lat.complexes[e].has_boundary:
                np.testing.assert_allclose(m["graph"].bnd_len, gu.ragged(g, f"{e}_bnd_len", t), rtol=RTOL)
                assert np.array_equal(m["graph"].bnd_side, gu.ragged(g, f"{e}_bnd_side", t))


# ------------------------------------------------------------------------------------------------
# GPU: kernels through the C ABI, facade, batched driver
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("path", gu.macro_files(), ids=_ids)
def test_gpu_macro_injected_against_golden(path):
    from flamingpy_b200.engine import TrialEngine

    g = gu.load_macro(path)
    lat = build_lattice(g["distance"], g["ec"], g["boundaries"])
    eng = TrialEngine(lat, g["delta"], g["p_swap"], g["weight_opts"])
    red = eng.run_macro_injected(g["state"], g["mean"], g["z"])
    batch = eng.fetch()
    assert np.array_equal(red["red_state"], g["red_state"])
    assert np.array_equal(red["bit_val"], g["bit_val"])
    np.testing.assert_allclose(red["p_phase_cond"], g["p_phase_cond"], rtol=RTOL, atol=1e-300)
    off = 0
    for e in lat.ec:
        ct = lat.complexes[e]
        sl = slice(off, off + ct.Q)
        np.testing.assert_allclose(batch.weights[:, sl], g[f"{e}_weights"], rtol=RTOL)
        assert np.array_equal(batch.bits[:, sl], g["bit_val"][:, ct.qubits])
        cb = batch.complexes[e]
        assert np.array_equal(cb.parity, g[f"{e}_parity"])
        for t in range(g["trials"]):
            assert np.array_equal(cb.odd(t), gu.ragged(g, f"{e}_odd", t))
            np.testing.assert_allclose(cb.pairs(t), gu.ragged(g, f"{e}_pair_len", t), rtol=RTOL)
            if ct.has_boundary:
                bl, bs = cb.boundary(t)
                np.testing.assert_allclose(bl, gu.ragged(g, f"{e}_bnd_len", t), rtol=RTOL)
                assert np.array_equal(bs, gu.ragged(g, f"{e}_bnd_side", t))
        off += ct.Q
    eng.close()


@pytest.mark.gpu
def test_gpu_macro_device_rng_replayed_through_the_oracle():
    """The Philox path at d=7: its own draws, read back, pushed through the restatement."""
    from flamingpy_b200 import _lib
    from flamingpy_b200.engine import TrialEngine

    lat = build_lattice(7, "primal", "open")
    delta, ps = 0.1, 0.2
    wo = {"method": "blueprint", "prob_precomputed": True}
    eng = TrialEngine(lat, delta, ps, wo, mode="compat")
    partner, sign = macro_tables(lat)
    batch = eng.run_macro_rng(31, 5, 6)
    red = eng.fetch_macro()
    # read the device draws back through the raw buffers of the handle
    import torch

    M = 4 * lat.N
    # the engine keeps the draws on the device; a second, injected run of the same draws must agree
    again = eng.run_macro_rng(31, 5, 6)
    assert np.array_equal(again.bits, batch.bits) and np.array_equal(again.weights, batch.weights)
    # statistics of the labels: reduced node is p only if all four micronodes are
    frac_p = (red["red_state"] == 2).mean()
    assert abs(frac_p - ps**4) < 0.01
    assert set(np.unique(red["bit_val"])) <= {0, 1}
    assert np.all((red["p_phase_cond"] >= 0) & (red["p_phase_cond"] <= 0.5))
    # batching independence: trials 5..10 in two calls
    a = eng.run_macro_rng(31, 5, 3)
    b = eng.run_macro_rng(31, 8, 3)
    assert np.array_equal(np.concatenate([a.bits, b.bits]), batch.bits)
    np.testing.assert_array_equal(np.concatenate([a.weights, b.weights]), batch.weights)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["macro_d3_open_primal_ps025", "macro_d3_open_dual_ps0", "macro_d5_open_primal_ps01",
                                  "macro_d4_periodic_primal_ps02"])
def test_gpu_facade_reproduces_reference_run(name):
    """CVMacroLayer(code, delta, p_swap).apply_noise(rng) + correct(...) from the reference's seed: same reduced
    lattice, same verdicts (reference tests/frontend/test_simulations.py:95-121, examples/macro_reduce.py)."""
    from flamingpy_b200.codes import SurfaceCode
    from flamingpy_b200.decoder import correct
    from flamingpy_b200.noise import CVMacroLayer

    g = gu.load_macro(os.path.join(gu.GOLDEN_DIR, name + ".npz"))
    code = SurfaceCode(g["distance"], g["ec"], g["boundaries"])
    layer = CVMacroLayer(code, delta=g["delta"], p_swap=g["p_swap"])
    rng = np.random.default_rng(int(g["seed"]))
    for t in range(g["trials"]):
        layer.apply_noise(rng)
        nodes = code.graph.nodes
        pts = [code.graph.to_points[i] for i in range(len(code.graph))]
        assert np.array_equal([nodes[p]["bit_val"] for p in pts], g["bit_val"][t])
        assert np.array_equal([2 if nodes[p]["state"] == "p" else 1 for p in pts], g["red_state"][t])
        np.testing.assert_allclose([nodes[p]["p_phase_cond"] for p in pts], g["p_phase_cond"][t], rtol=RTOL, atol=1e-300)
        ok = correct(code, "MWPM", g["weight_opts"], decoder_opts={"backend": "networkx"})
        assert ok == bool(g["success"][t]), t
    with pytest.raises(ValueError):
        correct(code, "MWPM", {"method": "blueprint", "delta": 0.1})  # needs prob_precomputed


@pytest.mark.gpu
def test_gpu_macro_monte_carlo():
    """Batched driver with the macronode layer: no failures at tiny delta (reference TestPassive,
    test_simulations.py:95-121), failures at large delta, device stream independent of the batch size."""
    from flamingpy_b200.codes import SurfaceCode
    from flamingpy_b200.noise import CVMacroLayer
    from flamingpy_b200.simulations import ec_monte_carlo

    args = {"weight_opts": {"method": "blueprint", "prob_precomputed": True}}
    for d, ec, b in [(3, "primal", "open"), (2, "primal", "periodic"), (3, "dual", "toric")]:
        code = SurfaceCode(d, ec, b)
        assert ec_monte_carlo(20, code, CVMacroLayer(code, delta=0.001, p_swap=0), "MWPM", args, seed=1) == 0
    code = SurfaceCode(5, "primal", "open")
    noise = CVMacroLayer(code, delta=0.12, p_swap=0.1)
    f1 = ec_monte_carlo(600, code, noise, "MWPM", args, seed=4, batch=100, mode="fresh")
    f2 = ec_monte_carlo(600, code, noise, "MWPM", args, seed=4, batch=256, mode="fresh")
    assert f1 == f2 and 0 < f1 < 400
    with pytest.raises(ValueError):
        ec_monte_carlo(10, code, noise, "MWPM", {"weight_opts": {"method": "blueprint", "delta": 0.1}})
