"""Fuzz the whole trial pipeline against the NumPy oracle on random small lattices (sizes 2..7 per axis, every
boundary / complex combination the code accepts, p_swap in {0, ..., 1}, blueprint / integer / uniform weights) and the
correction-path and Union-Find stages with it; every fourth configuration also through the macronode kernels, every
other fourth through the bit-level arm with and without the all-pairs table.  Meant for the CPU SIMT emulation of the kernels (no GPU needed):
    python tests/emu/build_emu.py
    FPB_LIB=tests/emu/_build/libfpb_emu.so FPB_BINDING=ctypes python scripts/emu_fuzz_pipeline.py [seconds] [seed] [lo] [hi]
(lo, hi: cubes per axis, default 2..7; 9..13 reaches the two-warp team kernels, S > 1024 cubes)
(on a GPU box it runs against the nvcc build the same way)."""
import sys, time; sys.path.insert(0, ".")
import numpy as np
from flamingpy_b200 import _lib, uf_native
from flamingpy_b200.engine import TrialEngine
from flamingpy_b200.lattice import build_lattice
from flamingpy_b200.lattice import macro_tables
from flamingpy_b200.noise import MacroSampler
from oracle import restate
from tests.test_gpu_parity import _compare_with_oracle


def bits_case(rng, lat, tag):
    """Bit-level trials (IidNoise arm): fpb_run_bits with the search and with the all-pairs table gather (also the
    row-driven one if FPB_GATHER_ROWS is set) against restate.trial_from_bits, unit weights, bit-exact."""
    T = int(rng.integers(1, 5))
    p = float(rng.choice([0.01, 0.05, 0.15, 0.5]))
    bit_val = (rng.random((T, lat.N)) < p).astype(np.uint8)
    qn = np.concatenate([lat.complexes[e].qubits for e in lat.ec])
    for table in (False, True):
        eng = TrialEngine(lat, 1.0, 0.0, {"method": "uniform"}, mode="fresh", device=0)
        try:
            if table:
                eng.build_table()
            batch = eng.run_bits(bit_val[:, qn])
            for t in range(T):
                res = restate.trial_from_bits(lat, bit_val[t])
                for e in lat.ec:
                    m, cb = res.per_complex[e], batch.complexes[e]
                    assert np.array_equal(cb.parity[t], m["parity"]) and np.array_equal(cb.odd(t), m["graph"].odd), ("bits", tag, table)
                    assert np.array_equal(cb.pairs(t), m["graph"].pair_len), ("bits pairs", tag, table, e, t)
                    if lat.complexes[e].has_boundary and cb.odd_count[t]:
                        bl, bs = cb.boundary(t)
                        assert np.array_equal(bl, m["graph"].bnd_len) and np.array_equal(bs, m["graph"].bnd_side), ("bits bnd", tag, table)
        finally:
            eng.close()
    return 2 * T


def macro_case(rng, lat, delta, ps, tag):
    """CVMacroLayer kernels (measure / reduce / precomputed-probability weights) against restate.macro_trial."""
    wo = [{"method": "blueprint", "prob_precomputed": True}, {"method": "uniform"},
          {"method": "blueprint", "prob_precomputed": True, "integer": True, "multiplier": 10}][int(rng.integers(0, 3))]
    partner, sign = macro_tables(lat)
    samp = MacroSampler(4 * lat.N, ps)
    T = int(rng.integers(1, 4))
    draws = [samp.next(rng) for _ in range(T)]
    eng = TrialEngine(lat, delta, ps, wo, mode="compat", device=0)
    try:
        red = eng.run_macro_injected(np.stack([d[0] for d in draws]), np.stack([d[1] for d in draws]),
                                     np.stack([d[2] for d in draws]))
        batch = eng.fetch()
        for t, (st, mean, z) in enumerate(draws):
            res = restate.macro_trial(lat, partner, sign, st, mean, z, delta, wo)
            assert np.array_equal(red["red_state"][t], res.macro["red_state"]), ("macro state", tag, wo)
            assert np.array_equal(red["bit_val"][t], res.macro["bit_val"]), ("macro bits", tag, wo)
            np.testing.assert_allclose(red["p_phase_cond"][t], res.macro["p_phase_cond"], rtol=1e-9, atol=1e-300)
            off = 0
            for e in lat.ec:
                ct, m, cb = lat.complexes[e], res.per_complex[e], batch.complexes[e]
                np.testing.assert_allclose(batch.weights[t, off:off + ct.Q], m["weights"], rtol=1e-9)
                assert np.array_equal(cb.parity[t], m["parity"]) and np.array_equal(cb.odd(t), m["graph"].odd)
                np.testing.assert_allclose(cb.pairs(t), m["graph"].pair_len, rtol=1e-9)
                if ct.has_boundary:
                    bl, bs = cb.boundary(t)
                    np.testing.assert_allclose(bl, m["graph"].bnd_len, rtol=1e-9)
                    assert np.array_equal(bs, m["graph"].bnd_side)
                off += ct.Q
    finally:
        eng.close()
    return T


budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lo_dim = int(sys.argv[3]) if len(sys.argv) > 3 else 2
hi_dim = int(sys.argv[4]) if len(sys.argv) > 4 else 7
rng = np.random.default_rng(seed)
t0, n_cfg, n_trials = time.time(), 0, 0
while time.time() - t0 < budget:
    bnd = str(rng.choice(["open", "toric", "periodic"]))
    dims = tuple(int(v) for v in rng.integers(lo_dim, hi_dim + 1, 3))
    if rng.random() < 0.4:
        dims = (dims[0],) * 3
    ec = str(rng.choice(["primal", "dual", "both"])) if bnd == "periodic" else str(rng.choice(["primal", "dual"]))
    delta = float(rng.choice([0.05, 0.08, 0.1, 0.13, 0.2]))
    ps = float(rng.choice([0.0, 0.05, 0.3, 0.6, 1.0]))
    ps = int(ps) if ps in (0.0, 1.0) else ps
    kind = int(rng.integers(0, 4))
    wopts = [{"method": "blueprint", "delta": delta}, {"method": "blueprint", "delta": float(rng.choice([0.05, 0.15]))},
             {"method": "blueprint", "delta": delta, "integer": True, "multiplier": int(rng.choice([1, 10, 100]))},
             {"method": "uniform"}][kind]
    try:
        lat = build_lattice(dims, ec, bnd)
    except ValueError:
        continue
    T = int(rng.integers(1, 6)) if hi_dim <= 8 else 1
    chain = restate.LabelChain(lat.N, ps)
    states, xs = [], []
    for _ in range(T):
        st = chain.next(rng)
        states.append(st)
        xs.append(restate.draw_quadratures(rng, restate.sigma_from_state(st, delta)))
    tag = (seed, n_cfg, dims, ec, bnd, delta, ps, wopts, T)
    eng = TrialEngine(lat, delta, ps, wopts, mode="compat", device=0)
    try:
        batch = eng.run_injected(np.stack(xs), np.stack(states))
        _compare_with_oracle(lat, batch, states, xs, delta, wopts)
        # every matching-graph length is the weight of the path the toggle kernel walks (any tie is fine)
        w = batch.weights
        off = 0
        for i, e in enumerate(lat.ec):
            ct = lat.complexes[e]
            cb = batch.complexes[e]
            for t in range(T):
                K = int(cb.odd_count[t])
                if K == 0:
                    continue
                pl = cb.pairs(t)
                qs = [(u, v) for u in range(K) for v in range(u + 1, K)][:12]
                if ct.has_boundary:
                    qs += [(u, -1) for u in range(min(K, 6))]
                    bl, _ = cb.boundary(t)
                for u, v in qs:
                    fl = eng.paths_toggle([t], [i], [u], [v])[t][off:off + ct.Q]
                    want = pl[u * K - u * (u + 1) // 2 + (v - u - 1)] if v >= 0 else bl[u]
                    got = w[t][off:off + ct.Q][fl.astype(bool)].sum()
                    assert abs(got - want) <= 1e-9 * max(1.0, abs(want)), ("toggle", tag, e, t, u, v, got, want)
            off += ct.Q
        # Union-Find on the same syndromes: device decoder against the host decoder
        flips, grown = eng.uf_decode(use_erasures=bool(ps), want_grown=True)
        _, parity, st = eng.fetch_syndrome(want_state=bool(ps))
        off = 0
        for i, e in enumerate(lat.ec):
            ct = lat.complexes[e]
            er = None
            if ps:
                er = uf_native.erasures(lat, np.ascontiguousarray(ct.qubits, np.int32), st)
            hf, hg = uf_native.decode_batch(ct, parity[i], er, want_grown=True)
            assert np.array_equal(grown[:, off:off + ct.Q], hg), ("uf grown", tag, e)
            off += ct.Q
    except ValueError as exc:  # odd syndrome on a complex without boundary cannot come from real errors
        raise AssertionError(("unexpected", tag, exc))
    finally:
        eng.close()
    n_cfg += 1
    n_trials += T
    if n_cfg % 4 == 0 and lat.N <= 600:
        n_trials += macro_case(rng, lat, delta, ps, tag)
    if n_cfg % 4 == 2 and lat.N <= 900:
        n_trials += bits_case(rng, lat, tag)
print(f"ok: {n_cfg} configurations, {n_trials} trials in {time.time() - t0:.0f} s (seed {seed})")
