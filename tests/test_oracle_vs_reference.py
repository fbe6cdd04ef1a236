"""Pin the CPU restatement (oracle/restate.py) to the live reference.

The reference stores no golden vectors for this path, so
the oracle is pinned by running the reference itself on seeded noise: same
NumPy seed -> same labels and draws -> bit-exact homodyne values, bits,
parities and odd sets; weights and path lengths to 1e-12 relative.
"""

import numpy as np
import pytest

from flamingpy_b200.lattice import build_lattice
from oracle import restate

pytestmark = pytest.mark.reference

CASES = [
    # d, ec, boundaries, delta, p_swap, weight opts, trials
    (3, "primal", "open", 0.09, 0.25, {"method": "blueprint", "delta": 0.09}, 6),
    (3, "dual", "open", 0.09, 0.25, {"method": "blueprint", "delta": 0.09}, 4),
    (4, "dual", "periodic", 0.09, 0, {"method": "blueprint", "delta": 0.09}, 4),
    (3, "primal", "toric", 0.1, 0.5, {"method": "blueprint", "delta": 0.1}, 4),
    (5, "primal", "open", 0.09, 0.25, {"method": "blueprint", "delta": 0.09}, 3),
    (4, "primal", "open", 0.09, 1, {"method": "uniform"}, 3),
    (4, "primal", "open", 0.09, 1, {"method": "blueprint", "delta": 0.09}, 2),
    (3, "both", "periodic", 0.09, 0, {"method": "blueprint", "delta": 0.09}, 4),
    (3, "primal", "open", 0.09, 0.25,
     {"method": "blueprint", "delta": 0.09, "integer": True, "multiplier": 100}, 3),
    ((2, 3, 4), "primal", "open", 0.12, 0.1, {"method": "blueprint", "delta": 0.12}, 3),
    # size-2 periodic axes: two cubes share two qubits, one edge kept (stabilizer_graph.py:295-298)
    (2, "primal", "periodic", 0.12, 0, {"method": "blueprint", "delta": 0.12}, 4),
    (2, "dual", "periodic", 0.12, 0.25, {"method": "blueprint", "delta": 0.12}, 4),
    ((2, 2, 3), "dual", "toric", 0.12, 0.25, {"method": "blueprint", "delta": 0.12}, 4),
    ((3, 2, 4), "primal", "periodic", 0.1, 0, {"method": "blueprint", "delta": 0.1}, 3),
]


@pytest.mark.parametrize("d,ec,b,delta,p_swap,wopts,trials", CASES)
def test_chain_matches_reference(reference, d, ec, b, delta, p_swap, wopts, trials):
    from oracle import capture, ref_shim

    code = ref_shim.make_code(d, ec, b)
    layer = capture.make_layer(code, delta, p_swap)
    lat = build_lattice(d, ec, b)
    chain = restate.TrialChain(lat, delta, p_swap, wopts)
    seed = 1234
    rng_ref = np.random.default_rng(seed)
    rng_mine = np.random.default_rng(seed)
    for t in range(trials):
        ref = capture.capture_trial(code, layer, lat, rng_ref, wopts, run_correct=True)
        mine = chain.next(rng_mine, want_graph=True, want_paths=True)
        assert np.array_equal(ref["state"], mine.state), t
        assert np.array_equal(ref["label_p"], mine.state == 2)
        assert np.array_equal(ref["sigma"], restate.sigma_from_state(mine.state, delta))
        assert np.array_equal(ref["x"], mine.x)
        for e in lat.ec:
            r, m = ref["complex"][e], mine.per_complex[e]
            assert np.array_equal(r["hom"], m["hom"])  # bit-exact
            assert np.array_equal(r["bits"], m["bits"])
            assert np.array_equal(r["parity"], m["parity"])
            np.testing.assert_allclose(m["weights"], r["weights"], rtol=1e-12, atol=0)
            g = m["graph"]
            assert np.array_equal(r["odd"], g.odd)
            np.testing.assert_allclose(g.pair_len, r["pair_len"], rtol=1e-12, atol=0)
            np.testing.assert_allclose(g.bnd_len, r["bnd_len"], rtol=1e-12, atol=0)
            assert np.array_equal(r["bnd_side"], g.bnd_side)
            K = len(g.odd)
            nb = K if (lat.complexes[e].has_boundary and K) else 0
            assert r["n_edges"] == K * (K - 1) // 2 + nb + nb * (nb - 1) // 2
        if wopts["method"] == "blueprint" and not wopts.get("integer") and p_swap != 1:
            # float weights: the optimum matching is unique up to zero-weight virtual pairs
            assert chain.correct(mine) == ref["success"], t


def test_binner_matches_numpy_divmod(reference):
    from flamingpy.cv.gkp import GKP_binner, integer_fractional

    rng = np.random.default_rng(7)
    a = restate.ALPHA
    x = np.concatenate([
        rng.normal(0, 3, 200000),
        rng.normal(0, 300, 50000),
        (np.arange(-40, 41) + 0.5) * a,
        np.nextafter((np.arange(-40, 41) + 0.5) * a, np.inf),
        np.nextafter((np.arange(-40, 41) + 0.5) * a, -np.inf),
        np.arange(-40, 41) * a,
        [0.0, -0.0, a / 2, -a / 2],
    ])
    n_ref, f_ref = integer_fractional(x, a)
    n, f = restate.integer_fractional(x, a)
    assert np.array_equal(n_ref, n)
    assert np.array_equal(f_ref, f)
    bits_ref = GKP_binner(x)
    bits, _ = restate.gkp_binner(x)
    assert np.array_equal(bits_ref, bits)


def test_z_err_cond_matches_reference(reference):
    from flamingpy.cv.gkp import Z_err_cond

    rng = np.random.default_rng(11)
    hom = rng.normal(0, 2, 500)
    for var in (0.09 * 3, 0.09 * 5, 0.5, 1e-3, 5.0):
        _, frac = restate.gkp_binner(hom)
        mine = restate.z_err_cond(np.full(len(hom), var), frac)
        ref = np.array([Z_err_cond(var, h) for h in hom])
        np.testing.assert_allclose(mine, ref, rtol=1e-12, atol=0)
