"""The C restatement (oracle/fpb_oracle.c) against the committed reference
vectors and against the NumPy restatement on seeded noise."""

import os

import numpy as np
import pytest

from flamingpy_b200.lattice import build_lattice
from oracle import c_oracle, restate
from tests import golden_util as gu


@pytest.mark.parametrize("path", gu.golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_c_oracle_reproduces_golden(path):
    g = gu.load(path)
    lat = build_lattice(g["distance"], g["ec"], g["boundaries"])
    batch = c_oracle.run(lat, g["delta"], g["p_swap"], g["weight_opts"], g["x"], g["state"])
    off = 0
    for e in lat.ec:
        ct = lat.complexes[e]
        sl = slice(off, off + ct.Q)
        assert np.array_equal(batch.hom[:, sl], g[f"{e}_hom"])
        assert np.array_equal(batch.bits[:, sl], g[f"{e}_bits"])
        np.testing.assert_allclose(batch.weights[:, sl], g[f"{e}_weights"], rtol=1e-12, atol=0)
        cb = batch.complexes[e]
        assert np.array_equal(cb.parity, g[f"{e}_parity"])
        for t in range(g["trials"]):
            assert np.array_equal(cb.odd(t), gu.ragged(g, f"{e}_odd", t))
            np.testing.assert_allclose(cb.pairs(t), gu.ragged(g, f"{e}_pair_len", t), rtol=1e-12, atol=0)
            if ct.has_boundary:
                bl, bs = cb.boundary(t)
                np.testing.assert_allclose(bl, gu.ragged(g, f"{e}_bnd_len", t), rtol=1e-12, atol=0)
                assert np.array_equal(bs, gu.ragged(g, f"{e}_bnd_side", t))
        off += ct.Q


def test_c_oracle_matches_numpy_restatement_d7():
    lat = build_lattice(7, "both", "periodic")
    delta, ps = 0.1, 0.2
    wopts = {"method": "blueprint", "delta": delta}
    chain = restate.LabelChain(lat.N, ps)
    rng = np.random.default_rng(5)
    states, xs = [], []
    for _ in range(3):
        st = chain.next(rng)
        states.append(st)
        xs.append(restate.draw_quadratures(rng, restate.sigma_from_state(st, delta)))
    batch = c_oracle.run(lat, delta, ps, wopts, np.stack(xs), np.stack(states), n_threads=2)
    for t in range(3):
        res = restate.trial_from_draws(lat, states[t], xs[t], delta, wopts)
        off = 0
        for e in lat.ec:
            ct = lat.complexes[e]
            m = res.per_complex[e]
            sl = slice(off, off + ct.Q)
            assert np.array_equal(batch.hom[t, sl], m["hom"])
            assert np.array_equal(batch.bits[t, sl], m["bits"])
            np.testing.assert_allclose(batch.weights[t, sl], m["weights"], rtol=1e-12, atol=0)
            cb = batch.complexes[e]
            assert np.array_equal(cb.odd(t), m["graph"].odd)
            np.testing.assert_allclose(cb.pairs(t), m["graph"].pair_len, rtol=1e-12, atol=0)
            off += ct.Q
