"""Union-Find + peeling decoder (``csrc/uf.cpp``, ``include/fpb_uf.h``) on CPU against the live
reference's vectors (``tests/golden/uf_*.npz``, ``oracle/gen_golden_uf.py``): the Support table when
the clusters stop growing is reproduced exactly; the recovery clears the syndrome inside the grown
edges; the logical verdict equals the reference's wherever it does not hang on the choice of the
spanning tree (the reference's own choice comes from rustworkx and from Python set order)."""

import ctypes
import json
import os
import re

import numpy as np
import pytest

from flamingpy_b200 import uf_native
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.decoder import uf_erasures
from flamingpy_b200.lattice import build_lattice
from tests import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _syndrome_of(ct, flips):
    _, ea, eb = uf_native.uf_graph(ct)
    syn = np.zeros((flips.shape[0], ct.S), np.uint8)
    for q in np.nonzero(ea >= 0)[0]:
        syn[:, ea[q]] ^= flips[:, q]
        if eb[q] < ct.S:
            syn[:, eb[q]] ^= flips[:, q]
    return syn


def _failures(ct, bits, flips):
    fail = np.zeros(bits.shape[0], bool)
    for plane in ct.planes:
        fail |= ((bits ^ flips)[:, plane].sum(axis=1) % 2).astype(bool)
    return fail


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fpb_uf.h")).read()
    declared = set(re.findall(r"\b(fpb_uf[a-z_]*)\s*\(", header))
    assert declared == {"fpb_uf_batch", "fpb_uf_erasures", "fpb_uf_check_correction", "fpb_uf_max_threads"}
    lib = ctypes.CDLL(uf_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


@pytest.mark.parametrize("path", gu.uf_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_against_reference_vectors(path):
    g = gu.load_uf(path)
    lat = build_lattice(g["distance"], g["ec"], g["boundaries"])
    T = g["trials"]
    fail = np.zeros(T, bool)
    for e in lat.ec:
        ct = lat.complexes[e]
        parity, w = g[f"{e}_parity"], g[f"{e}_weights"]
        flips, grown = uf_native.decode_batch(ct, parity, (w == -1).astype(np.uint8), want_grown=True)
        assert np.array_equal(grown, g[f"{e}_grown"])  # the reference's Support table, edge for edge
        assert np.array_equal(_syndrome_of(ct, flips), parity)  # "no unsatisfied stabilizers after recovery"
        assert np.all(flips <= grown)
        assert np.array_equal(_syndrome_of(ct, g[f"{e}_ref_flips"]), parity)  # the reference's recovery, same property
        fail |= _failures(ct, g[f"{e}_bits"], flips)
    agree = int((~fail == g["success"]).sum())
    if all(b == "open" for b in lat.boundaries):
        assert agree == T
    else:  # closed directions: a cluster that wraps around holds a non-trivial cycle; its verdict is a coin flip
        assert agree >= T // 2


def test_erasures_are_the_reference_weights():
    """``uf_erasures`` (two or more p-squeezed neighbours) against ``assign_weights(code, "UF")`` of the reference."""
    for path in gu.uf_files():
        g = gu.load_uf(path)
        code = SurfaceCode(g["distance"], g["ec"], g["boundaries"])
        er = uf_erasures(code, g["state"])
        want = np.concatenate([g[f"{e}_weights"] == -1 for e in code.ec], axis=1)
        if er is None:
            assert not want.any()
        else:
            assert np.array_equal(er.astype(bool), want)


def test_graph_of_the_decoder():
    """One edge per syndrome qubit, boundary points are leaves; the doubly shared faces of a size-2
    periodic axis keep one edge (stabilizer_graph.py:295-298)."""
    for d, b, ec in [(3, "open", "primal"), (5, "open", "dual"), (4, "periodic", "primal"), (2, "periodic", "primal")]:
        ct = build_lattice(d, ec, b).complexes[ec]
        n_nodes, ea, eb = uf_native.uf_graph(ct)
        present = ea >= 0
        assert np.all(ea[present] < ct.S) and np.all(eb[present] > ea[present])
        leaves = eb[present][eb[present] >= ct.S]
        assert len(set(leaves.tolist())) == len(leaves) == n_nodes - ct.S
        if d == 2:
            assert int((~present).sum()) == ct.Q // 2  # every pair of cubes shares two qubits
        else:
            assert present.all()
        assert n_nodes - ct.S == (len(ct.low_points) + len(ct.high_points) if ct.has_boundary else 0)


@pytest.mark.parametrize("d,b", [(9, "open"), (8, "periodic"), (6, "toric")])
def test_recovery_clears_random_syndromes(d, b):
    """Size-independent property at a larger size: whatever the syndrome and the erasures, the
    recovery lies inside the grown edges and reproduces the syndrome exactly."""
    ct = build_lattice(d, "primal", b).complexes["primal"]
    rng = np.random.default_rng(d)
    T = 40
    _, ea, eb = uf_native.uf_graph(ct)
    err = (rng.random((T, ct.Q)) < 0.06).astype(np.uint8)
    err[:, ea < 0] = 0
    parity = _syndrome_of(ct, err)  # a syndrome some error pattern produces (even on closed complexes)
    erased = (rng.random((T, ct.Q)) < 0.05).astype(np.uint8)
    flips, grown = uf_native.decode_batch(ct, parity, erased, threads=2, want_grown=True)
    assert np.array_equal(_syndrome_of(ct, flips), parity)
    assert np.all(flips <= grown) and np.all(grown[:, ea >= 0] >= erased[:, ea >= 0])
    again = uf_native.decode_batch(ct, parity, erased, threads=1)
    assert np.array_equal(flips, again)  # deterministic, independent of the thread count
    assert not uf_native.decode_batch(ct, np.zeros((3, ct.S), np.uint8)).any()  # no syndrome, no recovery


def test_odd_syndrome_without_boundary_is_an_error():
    ct = build_lattice(3, "primal", "periodic").complexes["primal"]
    parity = np.zeros((2, ct.S), np.uint8)
    parity[1, 4] = 1
    with pytest.raises(ValueError, match="trial 1"):
        uf_native.decode_batch(ct, parity)


def test_same_noise_failure_counts_recorded_with_the_reference():
    """oracle/gen_failure_rates_uf_ref.py decodes every reference trial with this decoder, too: on open
    boundaries the counts agree to a few in thousands of trials."""
    recs = json.load(open(os.path.join(ROOT, "tests", "golden", "failure_rates_uf_reference.json")))
    assert any(r["boundaries"] == "open" for r in recs)
    for r in recs:
        if r["boundaries"] == "open":
            assert abs(r["failures"] - r["failures_same_noise"]) <= max(3, r["trials"] // 1000)


# ----------------------------------------------------------------------------------------------
# The facade (assign_weights / uf_decoder / correct / ec_monte_carlo with decoder="UF") on CPU: an
# engine stand-in answers the device calls of that path from the C oracle, so that the host logic is
# covered without a GPU.  tests/test_gpu_uf.py runs the same replay against the real engine.
# ----------------------------------------------------------------------------------------------
class OracleEngine:
    """What the Union-Find path asks of ``TrialEngine``, answered by ``oracle/c_oracle`` (CPU)."""

    def __init__(self, lat, delta, p_swap, mode="compat"):
        self.lat, self.delta, self.p_swap, self.mode = lat, delta, p_swap, mode
        self._batch = self._state = None

    def run_injected(self, x, state, stage=2):
        from oracle import c_oracle

        self._state = np.ascontiguousarray(state, np.uint8)
        self._batch = c_oracle.run(self.lat, self.delta, self.p_swap, {"method": "uniform"}, x, state, stage=stage)
        return self._batch

    def run_rng(self, seed, first, n, stage=2, fetch=True):
        from flamingpy_b200.noise import sample_batch

        x, st = sample_batch(np.random.default_rng([seed, first]), self.lat.N, self.delta, self.p_swap, n, fresh=True)
        return self.run_injected(x, st, stage)

    def fetch_syndrome(self, want_state=False, packed_bits=False):  # unpacked bytes either way: both are accepted
        return (self._batch.bits, [self._batch.complexes[e].parity for e in self.lat.ec],
                self._state if want_state else None)

    def timings(self):
        return {"total_ms": 0.0}


def replay_uf_golden(g, code, layer):
    """Shared with tests/test_gpu_uf.py: same NumPy seed as the reference run behind the golden file ->
    same draws, bits and UF weights; the recovery leaves no unsatisfied stabilizer; on open boundaries
    the verdict of correct(code, "UF") is the reference's, trial after trial."""
    from flamingpy_b200.decoder import assign_weights, correct, recovery, uf_decoder

    rng = np.random.default_rng(int(g["seed"]))
    open_code = all(b == "open" for b in code.lattice.boundaries)
    agree = 0
    for t in range(g["trials"]):
        layer.apply_noise(rng)
        assert np.array_equal(code._trial[1], g["state"][t])
        assign_weights(code, "UF", method="uniform")
        for e in code.ec:
            ct = code.lattice.complexes[e]
            assert np.array_equal(np.array(layer.bit_values(ct.qubits)), g[f"{e}_bits"][t])
            w = np.array([code.graph.nodes[p]["weight"] for p in getattr(code, e + "_syndrome_coords")])
            assert np.array_equal(w, g[f"{e}_weights"][t])
        if t % 4 == 0:  # the step-by-step form: uf_decoder -> recovery -> no odd stabilizers
            for e in code.ec:
                recovery(uf_decoder(code, e), code, e)
                assert not list(getattr(code, e + "_stab_graph").odd_parity_stabilizers())
        ok = correct(code, "UF", {"method": "uniform"})
        agree += ok == bool(g["success"][t])
        if open_code:
            assert ok == bool(g["success"][t]), t
    assert agree >= g["trials"] // 2


@pytest.mark.parametrize("path", gu.uf_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_facade_replays_reference_run_on_the_oracle_engine(path):
    from flamingpy_b200.noise import CVLayer

    g = gu.load_uf(path)
    code = SurfaceCode(g["distance"], g["ec"], g["boundaries"])
    stub = OracleEngine(code.lattice, g["delta"], g["p_swap"])
    code.engine = lambda *a, **k: stub
    replay_uf_golden(g, code, CVLayer(code, delta=g["delta"], p_swap=g["p_swap"]))


def test_monte_carlo_driver_with_uf_on_the_oracle_engine():
    from flamingpy_b200.noise import CVLayer
    from flamingpy_b200.simulations import ec_monte_carlo

    code = SurfaceCode(3, "primal", "open")
    stub = OracleEngine(code.lattice, 0.09, 0.25, mode="fresh")
    code.engine = lambda *a, **k: stub
    layer = CVLayer(code, delta=0.09, p_swap=0.25)
    args = {"weight_opts": {"method": "uniform"}}
    errors, st = ec_monte_carlo(3000, code, layer, "UF", args, seed=11, batch=1000, mode="fresh", return_stats=True)
    assert st["trials_done"] == 3000
    # fresh-layer rate of the reference's UF decoder at these settings is about 0.27 (compat: 4347 / 16000)
    assert 0.2 < errors / 3000 < 0.36
    with pytest.raises(Exception, match="Incompatible decoder"):
        ec_monte_carlo(10, code, layer, "UF", {"weight_opts": {"method": "blueprint", "delta": 0.09}})
    with pytest.raises(KeyError):
        ec_monte_carlo(10, code, layer, "BP", args)


def test_restatement_is_pinned_to_the_reference_and_agrees_with_the_library_beyond_it():
    """oracle/restate_uf.py (plain Python, the reference's data structures) reproduces the reference's Support
    tables on the golden runs; then it is the oracle for csrc/uf.cpp on seeded syndromes and erasures at sizes the
    golden files do not cover: identical Support tables, and recoveries with identical syndrome."""
    from oracle import restate_uf

    for path in gu.uf_files():
        g = gu.load_uf(path)
        lat = build_lattice(g["distance"], g["ec"], g["boundaries"])
        for e in lat.ec:
            ct = lat.complexes[e]
            n_nodes, ea, eb = uf_native.uf_graph(ct)
            for t in range(min(g["trials"], 6)):
                flips, grown = restate_uf.decode(n_nodes, ct.S, ea, eb, g[f"{e}_parity"][t], g[f"{e}_weights"][t] == -1)
                assert np.array_equal(grown, g[f"{e}_grown"][t])
                assert np.array_equal(_syndrome_of(ct, flips[None])[0], g[f"{e}_parity"][t])
    rng = np.random.default_rng(77)
    for d, b, p_err, p_er in [(7, "open", 0.05, 0.08), (6, "periodic", 0.04, 0.0), (5, "toric", 0.08, 0.15), (9, "open", 0.03, 0.02)]:
        ct = build_lattice(d, "primal", b).complexes["primal"]
        n_nodes, ea, eb = uf_native.uf_graph(ct)
        T = 10
        err = (rng.random((T, ct.Q)) < p_err).astype(np.uint8)
        err[:, ea < 0] = 0
        parity = _syndrome_of(ct, err)
        erased = (rng.random((T, ct.Q)) < p_er).astype(np.uint8)
        flips, grown = uf_native.decode_batch(ct, parity, erased, want_grown=True)
        for t in range(T):
            f_o, g_o = restate_uf.decode(n_nodes, ct.S, ea, eb, parity[t], erased[t])
            assert np.array_equal(grown[t], g_o), (d, b, t)
            assert np.array_equal(_syndrome_of(ct, f_o[None])[0], parity[t])
            assert np.array_equal(_syndrome_of(ct, flips[t][None])[0], parity[t])


@pytest.mark.parametrize("d,ec,bnd", [(5, "primal", "open"), (4, "dual", "toric"), (3, "both", "periodic"), ((2, 3, 4), "dual", "open")])
def test_check_correction_in_the_library_equals_the_numpy_restatement(d, ec, bnd):
    """``decoder.logical_failures`` (``fpb_uf_check_correction``: plane parities of bit value XOR recovery,
    decoders/decoder.py:197-214) on unpacked bytes and on the device's packed words against plain NumPy."""
    from flamingpy_b200.decoder import logical_failures
    from flamingpy_b200.engine import TrialBatch, unpack_bits

    code = SurfaceCode(d, ec, bnd)
    Q = sum(code.lattice.complexes[e].Q for e in code.ec)
    rng = np.random.default_rng(Q)
    T = 257
    words = rng.integers(0, 2**32, (T, (Q + 31) // 32), dtype=np.uint32)
    flips = rng.integers(0, 2, (T, Q), dtype=np.uint8)
    bits = unpack_bits(words, Q)
    corrected = bits ^ flips
    want = np.zeros(T, bool)
    off = 0
    for e in code.ec:
        ct = code.lattice.complexes[e]
        for plane in ct.planes:
            want |= (corrected[:, off + np.asarray(plane)].sum(axis=1) % 2).astype(bool)
        off += ct.Q
    assert 0 < want.sum() < T
    assert np.array_equal(logical_failures(code, TrialBatch(T, 2, bits=bits), flips), want)
    assert np.array_equal(logical_failures(code, TrialBatch(T, 2, bits=words), flips), want)
    assert np.array_equal(logical_failures(code, TrialBatch(T, 2, bits=bits), flips[:, ::1].copy(order="F")), want)
    # the recovery as the device delivers it (fpb_paths_toggle / fpb_uf_decode): packed words, LSB first
    fw = np.packbits(np.pad(flips, ((0, 0), (0, words.shape[1] * 32 - Q))), axis=1, bitorder="little").view(np.uint32)
    assert np.array_equal(logical_failures(code, TrialBatch(T, 2, bits=words), fw), want)
    assert np.array_equal(logical_failures(code, TrialBatch(T, 2, bits=bits), fw), want)
