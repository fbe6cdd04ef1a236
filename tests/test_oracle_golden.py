"""The CPU restatement against the committed reference vectors (no reference
tree needed, runs on the GPU box too)."""

import os

import numpy as np
import pytest

from flamingpy_b200.lattice import build_lattice
from oracle import restate
from tests import golden_util as gu


@pytest.mark.parametrize("path", gu.golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_restatement_reproduces_golden(path):
    g = gu.load(path)
    lat = build_lattice(g["distance"], g["ec"], g["boundaries"])
    assert len(gu.golden_files()) >= 10
    for t in range(g["trials"]):
        res = restate.trial_from_draws(lat, g["state"][t], g["x"][t], g["delta"], g["weight_opts"],
                                       want_graph=True, want_paths=True)
        for e in lat.ec:
            m = res.per_complex[e]
            assert np.array_equal(m["hom"], g[f"{e}_hom"][t])
            assert np.array_equal(m["bits"], g[f"{e}_bits"][t])
            assert np.array_equal(m["parity"], g[f"{e}_parity"][t])
            np.testing.assert_allclose(m["weights"], g[f"{e}_weights"][t], rtol=1e-12, atol=0)
            gr = m["graph"]
            assert np.array_equal(gr.odd, gu.ragged(g, f"{e}_odd", t))
            np.testing.assert_allclose(gr.pair_len, gu.ragged(g, f"{e}_pair_len", t), rtol=1e-12, atol=0)
            np.testing.assert_allclose(gr.bnd_len, gu.ragged(g, f"{e}_bnd_len", t), rtol=1e-12, atol=0)
            assert np.array_equal(gr.bnd_side, gu.ragged(g, f"{e}_bnd_side", t))
        if g["method"] == "blueprint" and not g["integer"] and g["p_swap"] != 1:
            ok = True
            for e in lat.ec:
                m = res.per_complex[e]
                good, _ = restate.decode_complex(lat.complexes[e], m["weights"], m["bits"], m["graph"])
                ok = ok and good
            assert ok == bool(g["success"][t])


def test_label_chain_carry_over():
    """SURVEY.md finding 0.3: with p_swap=0 a reused layer alternates noisy /
    silent trials; a fresh layer is always noisy."""
    rng = np.random.default_rng(0)
    chain = restate.LabelChain(10, 0)
    states = [chain.next(rng) for _ in range(4)]
    assert [int(s.sum()) for s in states] == [10, 0, 10, 0]
    fresh = restate.LabelChain(10, 0, fresh=True)
    assert all(int(fresh.next(rng).sum()) == 10 for _ in range(3))
    full = restate.LabelChain(10, 1)
    assert all((full.next(rng) == 2).all() for _ in range(3))
