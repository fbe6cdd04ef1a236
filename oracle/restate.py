"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's Monte-Carlo EC trial.

This is the parity oracle for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it; the product package never does (it fails loudly
when the CUDA library is missing instead of falling back to this).

Pinned against the live reference (XanaduAI/flamingpy v0.10.1b1.dev3, imported
from ``/root/reference`` through ``oracle/ref_shim.py``) by
``tests/test_oracle_vs_reference.py`` on shared seeded noise, and against the
committed vectors in ``tests/golden/`` (generated from the reference by
``oracle/gen_golden.py``).  Each function cites the reference lines it follows
(paths relative to the reference root).

All arrays are per trial; a "complex" is one error complex (primal / dual) with
the static tables of ``flamingpy_b200.lattice.ComplexTables``.
"""

from __future__ import annotations

import heapq
import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

ALPHA = float(np.sqrt(np.pi))  # bin width, cv/gkp.py:53
FLOAT_MIN = 2.2250738585072014e-308  # sys.float_info.min, decoders/decoder.py:27
P_COUNT_ERR = {2: 1 / 4, 3: 1 / 3, 4: 2 / 5}  # decoders/decoder.py:86


# --------------------------------------------------------------------------------------
# a2: state labels with the carry-over of a reused CVLayer
# --------------------------------------------------------------------------------------
class LabelChain:
    """``CVLayer.populate_states`` for one reused layer (= one reference rank).

    Follows ``noise/cv.py:121-140,306-321``: if ``p_swap`` is truthy draw
    ``num_p = rng.binomial(N, p_swap)`` then ``rng.choice(range(N), num_p,
    replace=False)`` (``p_swap == 1`` takes all nodes without touching the
    RNG); the GKP set is everything not in ANY entry of ``self.states`` - and the
    previous trial's ``states["GKP"]`` is still one of those entries, so
    ``GKP_t = V \\ (p_t U GKP_{t-1})`` (SURVEY.md finding 0.3).  Nodes in neither
    set keep standard deviation 0 in ``_covs_sampler``.
    """

    def __init__(self, N: int, p_swap, fresh: bool = False):
        self.N = N
        self.p_swap = p_swap
        self.fresh = fresh
        self.prev_gkp = np.zeros(N, bool)

    def next(self, rng) -> np.ndarray:
        """Return uint8[N] state codes: 0 silent, 1 noisy GKP, 2 p-squeezed."""
        N = self.N
        p_mask = np.zeros(N, bool)
        if self.p_swap:
            if self.p_swap == 1:
                p_mask[:] = True
            else:
                num_p = rng.binomial(N, self.p_swap)
                inds = rng.choice(range(N), size=int(np.floor(num_p)), replace=False)
                p_mask[inds] = True
        prev = np.zeros(N, bool) if self.fresh else self.prev_gkp
        gkp = ~p_mask & ~prev
        self.prev_gkp = gkp
        state = np.zeros(N, np.uint8)
        state[gkp] = 1
        state[p_mask] = 2
        return state


def sigma_from_state(state: np.ndarray, delta: float) -> np.ndarray:
    """float32[2N] standard deviations, ``noise/cv.py:277-290`` ("initial"
    sampling order): q block then p block."""
    N = state.shape[0]
    covs = np.zeros(2 * N, dtype=np.float32)
    noise_q = {2: 1 / (2 * delta) ** 0.5, 1: (delta / 2) ** 0.5}
    noise_p = {2: (delta / 2) ** 0.5, 1: (delta / 2) ** 0.5}
    for code in (2, 1):
        idx = np.nonzero(state == code)[0]
        if len(idx):
            covs[idx] = noise_q[code]
            covs[idx + N] = noise_p[code]
    return covs


def draw_quadratures(rng, sigma: np.ndarray) -> np.ndarray:
    """``outcomes = rng.normal(means, covs)`` with float32 zero means,
    ``noise/cv.py:205,323-326``."""
    means = np.zeros(sigma.shape[0], dtype=np.float32)
    return rng.normal(means, sigma)


# --------------------------------------------------------------------------------------
# a4: CZ entangling  p'_i = sum_j A_ij q_j + p_i
# --------------------------------------------------------------------------------------
def entangle(lat, x: np.ndarray) -> np.ndarray:
    """p-homodyne value of every node after ``SCZ_apply`` (``cv/ops.py:65,84``):
    the sparse product with ``[[I, 0], [A, I]]`` accumulates, for row N+i,
    ``0 + A_ij1 x_j1 + A_ij2 x_j2 + ... + x_{N+i}`` with neighbours in ascending
    index order.  Returns float64[N]; the caller keeps the syndrome entries
    (``noise/cv.py:213``)."""
    N = lat.N
    q, p = x[:N], x[N:]
    out = np.zeros(N, np.float64)
    deg = np.diff(lat.adj_indptr)
    for k in range(int(deg.max())):
        has = deg > k
        pos = lat.adj_indptr[:-1][has] + k
        out[has] = out[has] + lat.adj_vals[pos].astype(np.float64) * q[lat.adj_indices[pos]]
    return out + p


# --------------------------------------------------------------------------------------
# a5: GKP binning (np.divmod semantics spelled out)
# --------------------------------------------------------------------------------------
def integer_fractional(x, alpha: float = ALPHA):
    """``cv/gkp.py:19-34``.  ``np.divmod`` is NumPy's ``npy_divmod``:
    ``mod = fmod(x, a)``; ``div = (x - mod) / a``; if ``mod != 0`` and its sign
    differs from ``a``'s, ``mod += a; div -= 1``; then ``floordiv = floor(div)``,
    bumped by one when ``div - floordiv > 0.5``.  Afterwards a remainder above
    ``alpha / 2`` moves to the next bin."""
    x = np.asarray(x, np.float64)
    mod = np.fmod(x, alpha)
    div = (x - mod) / alpha
    fix = (mod != 0) & (mod < 0)
    mod = np.where(fix, mod + alpha, mod)
    div = np.where(fix, div - 1.0, div)
    mod = np.where(mod == 0, np.copysign(0.0, alpha), mod)
    floordiv = np.floor(div)
    floordiv = np.where(div - floordiv > 0.5, floordiv + 1.0, floordiv)
    floordiv = np.where(div == 0, np.copysign(0.0, x / alpha), floordiv)
    large = (mod > alpha / 2).astype(int)
    f = mod - alpha * large
    n = floordiv.astype(int) + large
    return n, f


def gkp_binner(x):
    """``GKP_binner`` (``cv/gkp.py:37-62``): bit = n mod 2 (Python modulo, 0/1);
    also returns the fractional part."""
    n, f = integer_fractional(x, ALPHA)
    return (n % 2).astype(np.uint8), f


# --------------------------------------------------------------------------------------
# a7: conditional phase-error probability
# --------------------------------------------------------------------------------------
def z_err_cond(var, frac):
    """Scalar branch of ``Z_err_cond`` (``cv/gkp.py:91-133``) with the default
    ``var_num=10, use_hom_val=False``, vectorised over qubits: numerator sums
    ``exp(-(f - (2n+1) sqrt(pi))^2 / var)`` and denominator
    ``exp(-(f - n sqrt(pi))^2 / var)`` over ``n = -10 .. 9`` (``np.arange(-n_max,
    n_max)`` - the asymmetric range is the reference's); 0 where the denominator
    underflows to 0."""
    var = np.asarray(var, np.float64)
    frac = np.asarray(frac, np.float64)
    n = np.arange(-10, 10)
    sq = np.sqrt(np.pi)
    num = np.zeros_like(frac)
    den = np.zeros_like(frac)
    for k in n:
        num = num + np.exp(-((frac - (2 * k + 1) * sq) ** 2) / var)
        den = den + np.exp(-((frac - k * sq) ** 2) / var)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(den != 0, num / den, 0.0)
    return out


# --------------------------------------------------------------------------------------
# a6: weights
# --------------------------------------------------------------------------------------
def p_neighbour_count(lat, state: np.ndarray) -> np.ndarray:
    """Number of neighbours labelled "p" for every node (``decoder.py:64-65``)."""
    is_p = (state == 2).astype(np.int32)
    csum = np.concatenate([[0], np.cumsum(is_p[lat.adj_indices])])
    return (csum[lat.adj_indptr[1:]] - csum[lat.adj_indptr[:-1]]).astype(np.int32)


def assign_weights(
    lat,
    nodes: np.ndarray,
    hom: np.ndarray,
    state: np.ndarray,
    delta: float,
    method: str = "blueprint",
    integer: bool = False,
    multiplier=1,
):
    """``assign_weights`` for the MWPM decoder (``decoders/decoder.py:31-116``)
    for the syndrome qubits ``nodes`` with homodyne values ``hom``."""
    nq = len(nodes)
    if method == "uniform":
        return np.ones(nq, np.float64)
    if method != "blueprint":
        raise ValueError(method)
    deg = lat.degree[nodes]
    pc = p_neighbour_count(lat, state)[nodes]
    _, frac = gkp_binner(hom)
    err = z_err_cond((deg + 1) * delta, frac)
    err = np.minimum(err, 0.5)
    err = np.where(err == 0, FLOAT_MIN, err)
    for c, e in P_COUNT_ERR.items():
        err = np.where(pc == c, e, err)
    if integer:
        # Python round() = round-half-even on the float product
        return np.array([float(round(-multiplier * math.log(e))) for e in err], np.float64)
    return -np.log(err)


# --------------------------------------------------------------------------------------
# a9: stabilizer parity
# --------------------------------------------------------------------------------------
def stabilizer_parity(ct, bits_local: np.ndarray) -> np.ndarray:
    """``Stabilizer.parity`` (``codes/stabilizer.py:43-55``) for every cube of a
    complex; ``bits_local`` is indexed by the complex's local qubit id."""
    b = np.where(ct.cube_qubit >= 0, bits_local[np.maximum(ct.cube_qubit, 0)], 0)
    return (b.sum(axis=1) % 2).astype(np.uint8)


# --------------------------------------------------------------------------------------
# a10 / a11: matching graph
# --------------------------------------------------------------------------------------
def _dijkstra(ct, w_local: np.ndarray, sources, use_connectors: bool):
    """Label-setting shortest paths on the stabilizer graph
    (``codes/graphs/stabilizer_graph.py:392-410`` -> networkx
    ``single_source_dijkstra``).  ``sources`` is a list of (cube, start_distance,
    predecessor_qubit).  Boundary points are leaves, so with
    ``use_connectors=False`` (``shortest_paths_without_high_low``) the search
    never leaves the cube grid."""
    S = ct.S
    dist = np.full(S, np.inf)
    pred = np.full(S, -1, np.int64)  # previous cube (-1 for a seed)
    pred_q = np.full(S, -1, np.int64)  # qubit crossed to get here
    done = np.zeros(S, bool)
    heap = []
    for c, d0, q0 in sources:
        if d0 < dist[c]:
            dist[c] = d0
            pred_q[c] = q0
            heapq.heappush(heap, (d0, c))
    while heap:
        d, u = heapq.heappop(heap)
        if done[u]:
            continue
        done[u] = True
        for k in range(6):
            v = ct.nbr[u, k]
            if v < 0 or v >= S:
                continue
            nd = d + w_local[ct.cube_qubit[u, k]]
            if nd < dist[v]:
                dist[v] = nd
                pred[v] = u
                pred_q[v] = ct.cube_qubit[u, k]
                heapq.heappush(heap, (nd, v))
    return dist, pred, pred_q


@dataclass
class MatchingData:
    odd: np.ndarray  # int32[K] odd cubes in stabilizer order
    pair_len: np.ndarray  # float64[K(K-1)/2] packed upper triangle, row-major (a < b)
    bnd_len: np.ndarray  # float64[K] (empty without boundary points)
    bnd_side: np.ndarray  # uint8[K] 0 low, 1 high
    pair_paths: Optional[Dict] = None  # (a, b) -> list of local qubit ids along the path
    bnd_paths: Optional[List] = None


def matching_graph(ct, w_local: np.ndarray, parity: np.ndarray, want_paths: bool = False) -> MatchingData:
    """``MatchingGraph.with_edges_from`` (``decoders/mwpm/matching.py:97-173``):
    complete graph on the odd cubes with shortest-path lengths computed without
    the LOW/HIGH connectors; with boundary points, one virtual partner per odd
    cube at the nearer connector (ties -> low, ``np.argmin``)."""
    odd = np.nonzero(parity)[0].astype(np.int32)
    K = len(odd)
    pair_len = np.empty(K * (K - 1) // 2, np.float64)
    pair_paths = {} if want_paths else None
    pos = 0
    for a in range(K - 1):
        dist, pred, pred_q = _dijkstra(ct, w_local, [(int(odd[a]), 0.0, -1)], False)
        pair_len[pos : pos + K - 1 - a] = dist[odd[a + 1 :]]
        if want_paths:
            for b in range(a + 1, K):
                pair_paths[(a, b)] = _walk(pred, pred_q, int(odd[b]))
        pos += K - 1 - a
    bnd_len = np.empty(0, np.float64)
    bnd_side = np.empty(0, np.uint8)
    bnd_paths = None
    if ct.has_boundary and K > 0:
        seeds_lo = [(int(c), float(w_local[ct.slot_qubit[s]]), int(ct.slot_qubit[s]))
                    for c, s in zip(ct.low_seed_cube, ct.low_seed_slot)]
        seeds_hi = [(int(c), float(w_local[ct.slot_qubit[s]]), int(ct.slot_qubit[s]))
                    for c, s in zip(ct.high_seed_cube, ct.high_seed_slot)]
        dlo, plo, qlo = _dijkstra(ct, w_local, seeds_lo, True)
        dhi, phi, qhi = _dijkstra(ct, w_local, seeds_hi, True)
        lo, hi = dlo[odd], dhi[odd]
        bnd_side = (hi < lo).astype(np.uint8)  # argmin((low, high)): tie -> low
        bnd_len = np.where(bnd_side == 1, hi, lo)
        if want_paths:
            bnd_paths = [
                _walk(phi, qhi, int(c)) if s else _walk(plo, qlo, int(c)) for c, s in zip(odd, bnd_side)
            ]
    return MatchingData(odd, pair_len, bnd_len, bnd_side, pair_paths, bnd_paths)


def _int_lengths(dist, pred, pred_q, w_local):
    """``RxStabilizerGraph._path_weight`` (``codes/graphs/stabilizer_graph.py:486-500``): rustworkx runs
    Dijkstra on the float weights, then the reference re-adds ``int(weight)`` of every edge along the
    returned path.  Evaluated for every cube from the float shortest-path tree, in settle order."""
    il = np.zeros(len(dist), np.int64)
    for v in np.argsort(dist, kind="stable"):
        if np.isfinite(dist[v]) and pred_q[v] >= 0:
            il[v] = (il[pred[v]] if pred[v] >= 0 else 0) + int(w_local[pred_q[v]])
    return il


def matching_graph_rx(ct, w_local: np.ndarray, parity: np.ndarray) -> MatchingData:
    """``matching_graph`` with the path weights of the reference's default rustworkx backend: the sum
    of ``int(weight)`` along the float-shortest path (see ``_int_lengths``); the nearer connector is
    chosen on these int lengths (``mwpm/matching.py:149-150``, ties -> low).  Restated from the code:
    rustworkx is not installable here, so this is unpinned against a live run (with float weights the
    shortest path is unique almost surely, so its Dijkstra's tie-breaking does not enter)."""
    odd = np.nonzero(parity)[0].astype(np.int32)
    K = len(odd)
    pair_len = np.empty(K * (K - 1) // 2, np.float64)
    pos = 0
    for a in range(K - 1):
        dist, pred, pred_q = _dijkstra(ct, w_local, [(int(odd[a]), 0.0, -1)], False)
        il = _int_lengths(dist, pred, pred_q, w_local)
        pair_len[pos : pos + K - 1 - a] = il[odd[a + 1 :]]
        pos += K - 1 - a
    bnd_len = np.empty(0, np.float64)
    bnd_side = np.empty(0, np.uint8)
    if ct.has_boundary and K > 0:
        out = []
        for cubes, slots in ((ct.low_seed_cube, ct.low_seed_slot), (ct.high_seed_cube, ct.high_seed_slot)):
            seeds = [(int(c), float(w_local[ct.slot_qubit[s]]), int(ct.slot_qubit[s])) for c, s in zip(cubes, slots)]
            dist, pred, pred_q = _dijkstra(ct, w_local, seeds, True)
            out.append(_int_lengths(dist, pred, pred_q, w_local)[odd])
        lo, hi = out
        bnd_side = (hi < lo).astype(np.uint8)
        bnd_len = np.where(bnd_side == 1, hi, lo).astype(np.float64)
    return MatchingData(odd, pair_len, bnd_len, bnd_side)


def _walk(pred, pred_q, target: int):
    """Qubits crossed on the tree path ending at ``target`` (any order)."""
    out = []
    c = target
    while c >= 0 and pred_q[c] >= 0:
        out.append(int(pred_q[c]))
        c = int(pred[c])
    return out


# --------------------------------------------------------------------------------------
# a12 / a13: decode, recover, check
# --------------------------------------------------------------------------------------
def mwpm_pairs(md: MatchingData, has_boundary: bool):
    """Min-weight perfect matching of the matching graph with networkx
    (``decoders/mwpm/matching.py:209-210``).  Nodes 0..K-1 are odd cubes,
    K..2K-1 their virtual partners.  Returns a list of (u, v)."""
    import networkx as nx

    K = len(md.odd)
    if K == 0:
        return []
    G = nx.Graph()
    pos = 0
    for a in range(K - 1):
        for b in range(a + 1, K):
            G.add_edge(a, b, inverse_weight=-md.pair_len[pos])
            pos += 1
    if has_boundary:
        for a in range(K):
            G.add_edge(a, K + a, inverse_weight=-md.bnd_len[a])
        for a in range(K - 1):
            for b in range(a + 1, K):
                G.add_edge(K + a, K + b, inverse_weight=-0.0)
    return list(nx.max_weight_matching(G, maxcardinality=True, weight="inverse_weight"))


def decode_complex(ct, w_local, bits_local, md: Optional[MatchingData] = None):
    """One complex of ``correct`` (``decoders/decoder.py:282-285``):
    ``mwpm_decoder`` (``mwpm/algos.py:50-96``) + ``recovery``
    (``decoder.py:140-141``) + the plane check of ``check_correction``
    (``decoder.py:186-214``).  Returns (success, corrected bits)."""
    parity = stabilizer_parity(ct, bits_local)
    if md is None or md.pair_paths is None:
        md = matching_graph(ct, w_local, parity, want_paths=True)
    K = len(md.odd)
    bits = bits_local.copy()
    for u, v in mwpm_pairs(md, ct.has_boundary):
        u, v = (u, v) if u < v else (v, u)
        if u >= K:
            continue  # virtual - virtual
        path = md.pair_paths[(u, v)] if v < K else md.bnd_paths[u]
        for q in path:
            bits[q] ^= 1
    ok = True
    for plane in ct.planes:
        ok = ok and (int(bits[plane].sum()) % 2 == 0)
    return ok, bits


# --------------------------------------------------------------------------------------
# whole trial
# --------------------------------------------------------------------------------------
@dataclass
class TrialResult:
    state: np.ndarray  # uint8[N]
    x: np.ndarray  # float64[2N] raw draws
    hom_all: np.ndarray  # float64[N] entangled p-homodyne of every node
    per_complex: Dict[str, dict]


def trial_from_draws(lat, state, x, delta, weight_opts=None, want_graph=True, want_paths=False) -> TrialResult:
    """noise -> bits -> weights -> parity -> matching graph for injected draws."""
    wo = {"method": "uniform", "integer": False, "multiplier": 1}
    wo.update(weight_opts or {})
    hom_all = entangle(lat, x)
    out = {}
    for ec in lat.ec:
        ct = lat.complexes[ec]
        hom = hom_all[ct.qubits]
        bits, frac = gkp_binner(hom)
        w = assign_weights(lat, ct.qubits, hom, state, wo.get("delta", delta), wo["method"],
                           wo["integer"], wo["multiplier"])
        par = stabilizer_parity(ct, bits)
        rec = {"hom": hom, "bits": bits, "frac": frac, "weights": w, "parity": par}
        if want_graph:
            rec["graph"] = matching_graph(ct, w, par, want_paths=want_paths)
        out[ec] = rec
    return TrialResult(state, x, hom_all, out)


def iid_bits(rng, n_nodes: int, error_probability: float) -> np.ndarray:
    """``IidNoise.apply_noise`` (``flamingpy/noise/iid.py:37-50``): one ``rng.random()`` per node,
    ``bit_val = int(draw < p)``.  The reference walks ``egraph.nodes`` in the insertion order of
    its ``RHG_graph`` (a Python-set artefact, ``surface_code.py:186-226``); this restatement draws
    in node-index order - same distribution, a different assignment of the stream to nodes - so
    parity with the live reference is pinned on injected BITS (``trial_from_bits``)."""
    return (rng.random(n_nodes) < error_probability).astype(np.uint8)


def trial_from_bits(lat, bit_val, want_graph=True, want_paths=False) -> TrialResult:
    """bits (one per lattice node, node-index order) -> uniform weights (``decoder.py:115-116``)
    -> parity -> matching graph: the decoding front end for ``IidNoise`` realisations."""
    bit_val = np.asarray(bit_val, np.uint8)
    out = {}
    for ec in lat.ec:
        ct = lat.complexes[ec]
        bits = bit_val[ct.qubits]
        w = np.ones(ct.Q, np.float64)
        par = stabilizer_parity(ct, bits)
        rec = {"hom": np.zeros(ct.Q), "bits": bits, "frac": np.zeros(ct.Q), "weights": w, "parity": par}
        if want_graph:
            rec["graph"] = matching_graph(ct, w, par, want_paths=want_paths)
        out[ec] = rec
    return TrialResult(None, None, None, out)


def correct_trial(lat, res: TrialResult) -> bool:
    """``correct`` on a restated trial: decode every complex, all logical checks must pass."""
    ok = True
    for ec in lat.ec:
        rec = res.per_complex[ec]
        good, _ = decode_complex(lat.complexes[ec], rec["weights"], rec["bits"], rec.get("graph"))
        ok = ok and good
    return ok


class TrialChain:
    """A reused ``CVLayer`` + the body of ``ec_mc_trial`` (``simulations.py:46-59``):
    same NumPy RNG call order as the reference, so with the same seed it sees
    the same labels and draws."""

    def __init__(self, lat, delta, p_swap, weight_opts=None, fresh=False):
        self.lat = lat
        self.delta = delta
        self.weight_opts = weight_opts
        self.labels = LabelChain(lat.N, p_swap, fresh=fresh)

    def next(self, rng, want_graph=True, want_paths=False) -> TrialResult:
        state = self.labels.next(rng)
        sig = sigma_from_state(state, self.delta)
        x = draw_quadratures(rng, sig)
        return trial_from_draws(self.lat, state, x, self.delta, self.weight_opts, want_graph, want_paths)

    def correct(self, res: TrialResult) -> bool:
        ok = True
        for ec in self.lat.ec:
            ct = self.lat.complexes[ec]
            rec = res.per_complex[ec]
            good, _ = decode_complex(ct, rec["weights"], rec["bits"], rec.get("graph"))
            ok = ok and good
        return ok


# --------------------------------------------------------------------------------------
# f4: CVMacroLayer (noise/cv.py:352-625) - macronode lattice reduced to the canonical lattice
# --------------------------------------------------------------------------------------
# q- and p-block of the four-splitter (cv/ops.py:104-127: beamsplitters on modes (1,0), (3,2), (2,0),
# (3,1), cast to float32); every entry is exactly +-1/2
SPLITTER = np.array([[1, 1, 1, 1], [-1, 1, -1, 1], [-1, -1, 1, 1], [1, -1, -1, 1]], np.float32) * np.float32(0.5)


def z_err_cond_hom(var, val):
    """``Z_err_cond(var, val, use_hom_val=True)`` (``cv/gkp.py:116-133``): the sums run over the whole
    homodyne value, the numerator over the odd (even) multiples of sqrt(pi) if the value bins to 0 (1)."""
    val = np.asarray(val, np.float64)
    bit, _ = gkp_binner(val)
    factor = 1 - bit.astype(np.int64)
    sq = np.sqrt(np.pi)
    num = np.zeros_like(val)
    den = np.zeros_like(val)
    for k in range(-10, 10):
        num = num + np.exp(-((val - (2 * k + factor) * sq) ** 2) / var)
        den = den + np.exp(-((val - k * sq) ** 2) / var)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(den != 0, num / den, 0.0)


def macro_trial_from_draws(lat, partner, sign, state, mean32, z, delta):
    """One ``CVMacroLayer.apply_noise`` (``noise/cv.py:375-403``) from its random draws.

    ``state`` uint8[4N]: micronode labels (2 = p-squeezed, anything else reads "GKP");
    ``mean32`` float32[4N]: the initial q means of the two-step sampling (``cv.py:323-349``);
    ``z`` float64[4N]: standard normals, the N star p-measurements then the 3N planet q-measurements
    in the reference's order (``cv.py:474-489``).  Returns per canonical node: reduced state label
    (2 = p), bit value, conditional phase error probability and the effective homodyne value."""
    N = lat.N
    M = 4 * N
    state = np.asarray(state)
    is_p = state == 2
    # CZ gates on the float32 means: p_i += sign * q_partner (ops.py:65,84; every micronode has <= 1 partner)
    q = np.asarray(mean32, np.float32)
    p = np.where(partner >= 0, sign.astype(np.float32) * q[np.maximum(partner, 0)], np.float32(0)).astype(np.float32)
    # stars first: GKP micronodes in index order, then the p-squeezed ones (cv.py:529-546)
    lab = is_p.reshape(N, 4)
    perm = np.argsort(lab, axis=1, kind="stable")  # False (GKP) before True (p), stable
    idx = (4 * np.arange(N)[:, None] + perm)  # micronode index at body position k = 0..3
    body = np.empty(M, np.int64)
    body[idx.ravel()] = np.tile(np.arange(1, 5), N)
    # four-splitter in float32, summed pairwise like NumPy's float32 matrix-vector product
    xq = q[idx]
    xp = p[idx]

    def split(x):
        t = SPLITTER[None, :, :] * x[:, None, :]  # [N, 4 out, 4 in]; products are exact
        return ((t[:, :, 0] + t[:, :, 1]) + (t[:, :, 2] + t[:, :, 3])).astype(np.float32)

    nq, np_ = split(xq), split(xp)
    sd = np.float32((delta / 2) ** 0.5)
    z = np.asarray(z, np.float64)
    star_p = np_[:, 0].astype(np.float64) + np.float64(sd) * z[:N]
    planet_q = nq[:, 1:].astype(np.float64) + np.float64(sd) * z[N:].reshape(N, 3)
    hq = np.zeros((N, 5))  # [macronode][body index] q outcome of the planets (cv.py:447-465)
    hq[:, 2:] = planet_q
    red_p = is_p[idx[:, 0]]  # reduced label = label of the star (cv.py:417-420)
    bit_val = np.zeros(N, np.uint8)
    p_err_out = np.zeros(N)
    outcome_out = np.zeros(N)
    for j in range(N):  # cv.py:566-625
        outcome = 2 * star_p[j]
        num_p = 0
        z_gkp = []
        for k in range(4):
            u = idx[j, k]
            nb = partner[u]
            if nb < 0:
                continue
            nbody = body[nb]
            m = hq[nb // 4]
            zval = 0.0 if nbody == 1 else (m[2] - m[4] if nbody == 2 else (m[3] - m[4] if nbody == 3 else m[2] + m[3]))
            if is_p[nb]:
                num_p += 1
                outcome -= zval
            else:
                z_gkp.append(zval)
        p_err = 0.0
        for zval in z_gkp:
            p_err += float(z_err_cond_hom(2 * delta, np.float64(zval)))
        p_err += float(z_err_cond_hom(2 * (2 + num_p) * delta, np.float64(outcome)))
        p_err = max(min(p_err, 0.5), 0)
        bits = int(gkp_binner(np.array([outcome]))[0][0])
        for zval in z_gkp:
            bits += int(gkp_binner(np.array([zval]))[0][0])
        bit_val[j] = bits % 2
        p_err_out[j] = p_err
        outcome_out[j] = outcome
    return {"red_state": np.where(red_p, 2, 1).astype(np.uint8), "bit_val": bit_val, "p_phase_cond": p_err_out,
            "outcome": outcome_out}


def macro_weights(lat, red_state, p_phase_cond, weight_opts):
    """``assign_weights(..., prob_precomputed=True)`` (``decoders/decoder.py:58-93``) for all nodes."""
    wo = {"method": "uniform", "integer": False, "multiplier": 1}
    wo.update(weight_opts or {})
    N = lat.N
    if wo["method"] == "uniform":
        return np.ones(N)
    pc = p_neighbour_count(lat, red_state)
    err = np.minimum(np.asarray(p_phase_cond, np.float64), 0.5)
    err = np.where(err == 0, FLOAT_MIN, err)
    table = np.array([0, 0, 1 / 4, 1 / 3, 2 / 5])
    err = np.where(pc <= 1, err, table[np.minimum(pc, 4)])
    if wo.get("integer"):
        return np.array([float(round(-wo["multiplier"] * math.log(e))) for e in err])
    return -np.log(err)


def macro_trial(lat, partner, sign, state, mean32, z, delta, weight_opts, want_graph=True) -> "TrialResult":
    """A complete macronode trial in the shape of ``trial_from_draws``: reduction, precomputed-
    probability weights, parities and matching graph per complex."""
    red = macro_trial_from_draws(lat, partner, sign, state, mean32, z, delta)
    w_nodes = macro_weights(lat, red["red_state"], red["p_phase_cond"], weight_opts)
    per = {}
    for ec in lat.ec:
        ct = lat.complexes[ec]
        w = w_nodes[ct.qubits]
        bits = red["bit_val"][ct.qubits]
        parity = stabilizer_parity(ct, bits)
        rec = {"hom": red["outcome"][ct.qubits], "bits": bits, "weights": w, "parity": parity}
        if want_graph:
            rec["graph"] = matching_graph(ct, w, parity)
        per[ec] = rec
    res = TrialResult(state=red["red_state"], x=None, hom_all=red["outcome"], per_complex=per)
    res.macro = red
    return res
