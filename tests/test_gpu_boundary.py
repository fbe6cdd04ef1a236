"""The drop-in boundary on the GPU: the pybind11 module ``flamingpy_b200.cpp.fpbpy`` driven directly
against the committed reference vectors, both bindings giving identical results, and the reference's
``MatchingGraph`` plug-in contract (``decoders/mwpm/matching.py:36-173``, ``mwpm/algos.py:23-96``)."""

import os

import numpy as np
import pytest

from flamingpy_b200 import _lib
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.decoder import assign_weights, build_match_graph, correct, mwpm_decoder
from flamingpy_b200.engine import TrialEngine, config_dict
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.matching import GpuMatchingGraph
from flamingpy_b200.noise import CVLayer
from tests import golden_util as gu

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _handle(g, lat, **kw):
    fpbpy = _lib.pybind_module()
    assert fpbpy is not None, "flamingpy_b200/cpp/fpbpy*.so missing: run __graft_entry__.build()"
    wo = {"method": "uniform", "integer": False, "multiplier": 1}
    wo.update(g["weight_opts"])
    cfg, q_off, Qtot = config_dict(lat, g["delta"], g["p_swap"], wo, **kw)
    return fpbpy.Handle(cfg), Qtot


@pytest.mark.parametrize("name", ["d3_open_primal_ps025", "d3_periodic_both_ps0", "d5_open_primal_ps05",
                                  "d2_periodic_both_ps0", "d223_toric_dual_ps025", "d3_open_primal_int100"])
def test_fpbpy_run_injected_against_golden(name):
    """run_injected -> sizes -> fetch_into / pair_lengths through the pybind11 module alone."""
    g = gu.load(os.path.join(gu.GOLDEN_DIR, name + ".npz"))
    lat = build_lattice(g["distance"], g["ec"], g["boundaries"])
    h, Qtot = _handle(g, lat)
    T = g["trials"]
    h.run_injected(g["x"], g["state"], _lib.STAGE_GRAPH)
    s = h.sizes()
    assert s["n_trials"] == T and s["n_complex"] == len(lat.ec) and s["total_qubits"] == Qtot
    bufs = {"hom": np.empty((T, Qtot)), "weights": np.empty((T, Qtot)),
            "bits": np.empty((T, (Qtot + 31) // 32), np.uint32)}
    for i, e in enumerate(lat.ec):
        ct = lat.complexes[e]
        assert s["total_odd"][i] == len(g[f"{e}_odd"]) and s["total_pairs"][i] == len(g[f"{e}_pair_len"])
        bufs[f"odd_count{i}"] = np.empty(T, np.int32)
        bufs[f"odd_ids{i}"] = np.empty(s["total_odd"][i], np.int32)
        if ct.has_boundary:
            bufs[f"bnd_len{i}"] = np.empty(s["total_odd"][i])
            bufs[f"bnd_side{i}"] = np.empty(s["total_odd"][i], np.uint8)
    h.fetch_into(bufs)
    off = 0
    for i, e in enumerate(lat.ec):
        ct = lat.complexes[e]
        sl = slice(off, off + ct.Q)
        assert np.array_equal(bufs["hom"][:, sl], g[f"{e}_hom"])  # bit-exact
        np.testing.assert_allclose(bufs["weights"][:, sl], g[f"{e}_weights"], rtol=RTOL, atol=0)
        bits = np.unpackbits(bufs["bits"].view(np.uint8), axis=1, bitorder="little")[:, sl]
        assert np.array_equal(bits, g[f"{e}_bits"])
        assert np.array_equal(bufs[f"odd_ids{i}"], g[f"{e}_odd"])
        assert np.array_equal(bufs[f"odd_count{i}"], np.diff(g[f"{e}_odd_off"]))
        pl = h.pair_lengths(i)  # freshly owned array, lemonpy style
        np.testing.assert_allclose(pl, g[f"{e}_pair_len"], rtol=RTOL, atol=0)
        if ct.has_boundary:
            np.testing.assert_allclose(bufs[f"bnd_len{i}"], g[f"{e}_bnd_len"], rtol=RTOL, atol=0)
            assert np.array_equal(bufs[f"bnd_side{i}"], g[f"{e}_bnd_side"])
        off += ct.Q
    t = h.timings()
    assert t["launches"] >= 4 and t["total_ms"] > 0
    h.close()
    with pytest.raises(RuntimeError):
        h.sizes()


def test_fpbpy_run_rng_equals_ctypes_binding(monkeypatch):
    """The two bindings drive the same library: identical device-RNG batches, and errors surface
    as ``_lib.FpbError`` from both."""
    lat = build_lattice(5, "primal", "open")
    wo = {"method": "blueprint", "delta": 0.09}
    eng_py = TrialEngine(lat, 0.09, 0.25, wo, mode="fresh")
    assert eng_py.binding == "pybind11"
    monkeypatch.setenv("FPB_BINDING", "ctypes")
    eng_ct = TrialEngine(lat, 0.09, 0.25, wo, mode="fresh")
    assert eng_ct.binding == "ctypes"
    monkeypatch.delenv("FPB_BINDING")
    a = eng_py.run_rng(7, 100, 24, want_draws=True)
    b = eng_ct.run_rng(7, 100, 24, want_draws=True)
    for k in ("hom", "bits", "weights", "state", "x"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    ca, cb = a.complexes["primal"], b.complexes["primal"]
    for k in ("parity", "odd_ids", "pair_len", "bnd_len", "bnd_side"):
        assert np.array_equal(getattr(ca, k), getattr(cb, k)), k
    sa, sb = eng_py.sizes(), eng_ct.sizes()
    assert (sa.n_trials, sa.total_odd, sa.total_pairs) == (sb.n_trials, sb.total_odd, sb.total_pairs)
    for eng in (eng_py, eng_ct):
        with pytest.raises(_lib.FpbError):
            eng.run_rng(7, -1, 4)
        with pytest.raises(_lib.FpbError):
            eng.paths_toggle([99], [0], [0], [1])
        eng.close()


def _reference_style_flips(code, ec, mg):
    """The loop of ``mwpm_decoder`` (``mwpm/algos.py:88-95``), verbatim on our objects."""
    import itertools as it

    stab_graph = getattr(code, ec + "_stab_graph")
    matching = mg.min_weight_perfect_matching()
    qubits_to_flip = set()
    for match in matching:
        if match not in it.product(mg.virtual_points, mg.virtual_points):
            path = mg.edge_path(match)
            pairs = [(path[i], path[i + 1]) for i in range(len(path) - 1)]
            for pair in pairs:
                common_vertex = stab_graph.edge_data(*pair)["common_vertex"]
                qubits_to_flip ^= {common_vertex}
    return matching, qubits_to_flip


@pytest.mark.parametrize("name", ["d5_open_primal_ps05", "d3_open_dual_ps025", "d5_periodic_dual_ps0",
                                  "d3_periodic_both_ps0", "d2_periodic_both_ps0"])
def test_matching_graph_plugin_contract(name):
    """``build_match_graph(code, ec, <MatchingGraph class>)`` calls ``matching_backend(ec, code)``
    (``mwpm/algos.py:29-47``): GpuMatchingGraph builds itself from the code's ``weight`` / ``bit_val``
    attributes, its lengths equal the reference's, and ``edge_path`` gives stabilizer-graph node
    paths that the reference's decoding loop consumes unchanged."""
    g = gu.load(os.path.join(gu.GOLDEN_DIR, name + ".npz"))
    code = SurfaceCode(g["distance"], g["ec"], g["boundaries"])
    layer = CVLayer(code, delta=g["delta"], p_swap=g["p_swap"])
    rng = np.random.default_rng(int(g["seed"]))
    wo = g["weight_opts"]
    for t in range(min(3, g["trials"])):
        layer.apply_noise(rng)
        assign_weights(code, "MWPM", **wo)
        bits_before = code.graph.bit_val.copy()
        all_flips = set()
        for e in code.ec:
            mg = build_match_graph(code, e, GpuMatchingGraph)  # the class itself, as the reference allows
            assert isinstance(mg, GpuMatchingGraph) and mg.ec == e
            sg = getattr(code, e + "_stab_graph")
            odd = list(sg.odd_parity_stabilizers())
            assert [s.index for s in odd] == gu.ragged(g, f"{e}_odd", t).tolist()
            K = len(odd)
            ref_pairs = gu.ragged(g, f"{e}_pair_len", t)
            pos = 0
            for a in range(K):
                for b in range(a + 1, K):
                    assert mg.edge_weight((odd[a], odd[b])) == pytest.approx(ref_pairs[pos], rel=RTOL)
                    pos += 1
            if sg.has_bound_points() and K:
                bl = gu.ragged(g, f"{e}_bnd_len", t)
                assert len(mg.virtual_points) == K
                bset = set(sg.bound_points())
                for i, vp in enumerate(mg.virtual_points):
                    assert vp[1] == i and vp[0] in bset  # (boundary point, i): matching.py:159
                    assert mg.edge_weight((odd[i], vp)) == pytest.approx(bl[i], rel=RTOL)
                    path = mg.edge_path((odd[i], vp))
                    assert path[0] == vp[0] and path[-1] is odd[i]
                    w = sum(sg.edge_data(a, b)["weight"] for a, b in zip(path[:-1], path[1:]))
                    assert w == pytest.approx(bl[i], rel=1e-9)
            # every real-pair path is a walk of adjacent cubes whose weights sum to the edge weight
            for a in range(min(K, 4)):
                for b in range(a + 1, min(K, 6)):
                    path = mg.edge_path((odd[a], odd[b]))
                    assert path[0] is odd[a] and path[-1] is odd[b]
                    w = sum(sg.edge_data(u, v)["weight"] for u, v in zip(path[:-1], path[1:]))
                    assert w == pytest.approx(mg.edge_weight((odd[a], odd[b])), rel=1e-9)
            matching, flips = _reference_style_flips(code, e, mg)
            nodes = [n for m in matching for n in m]
            assert len(set(nodes)) == len(nodes) == K + len(mg.virtual_points)
            assert flips == mwpm_decoder(code, e, "native")
            all_flips |= flips
        # the reference-style decode flips exactly what correct() flips
        ok = correct(code, "MWPM", wo)
        flipped = {code.graph.to_points[i] for i in np.nonzero(code.graph.bit_val != bits_before)[0]}
        if not wo.get("integer"):
            assert flipped == all_flips
        assert ok == bool(g["success"][t]) or wo.get("integer")


def test_run_weighted_equals_injected():
    """``fpb_run_weighted`` (weights + bit values in) gives the graph of the run that made them."""
    lat = build_lattice(6, "both", "periodic")
    wo = {"method": "blueprint", "delta": 0.1}
    eng = TrialEngine(lat, 0.1, 0.0, wo, mode="fresh")
    a = eng.run_rng(11, 0, 6)
    b = eng.run_weighted(a.weights, a.bits)
    assert np.array_equal(b.weights, a.weights) and np.array_equal(b.bits, a.bits)
    assert not b.hom.any()
    for e in lat.ec:
        for k in ("parity", "odd_ids", "pair_len"):
            assert np.array_equal(getattr(a.complexes[e], k), getattr(b.complexes[e], k)), (e, k)
    with pytest.raises(_lib.FpbError):
        eng.run_weighted(a.weights, a.bits, stage=_lib.STAGE_NOISE)
    eng.close()
