"""Matching graphs built from the GPU outputs (``flamingpy/decoders/mwpm/matching.py``).

``GpuMatchingGraph`` implements the reference's ``MatchingGraph`` interface
(``add_edge`` / ``edge_weight`` / ``edge_path`` / ``min_weight_perfect_matching`` /
``total_weight_of`` / ``virtual_points`` / ``ec``, ``matching.py:36-123``) on top of the
dense lengths the CUDA pipeline returns: K odd cubes, the packed upper triangle
of their pairwise shortest-path lengths and, with boundary points, one virtual
partner per cube at its nearer connector (virtual-virtual edges have weight 0).
"""

from __future__ import annotations

import itertools as it
from typing import List, Optional, Tuple

import numpy as np


def tri_index(a: int, b: int, K: int) -> int:
    """Position of pair (a < b) in the packed row-major upper triangle."""
    return a * K - a * (a + 1) // 2 + (b - a - 1)


def dense_inverse_weights(K: int, pair_len: np.ndarray, bnd_len: Optional[np.ndarray]) -> np.ndarray:
    """The ``-length`` adjacency the reference feeds lemon
    (``decoders/mwpm/lemon.py:32``): n = K (no boundary) or 2K nodes, real cubes first."""
    n = 2 * K if bnd_len is not None else K
    A = np.zeros((n, n), np.float64)
    iu = np.triu_indices(K, 1)
    A[:K, :K][iu] = -pair_len
    A[:K, :K] += A[:K, :K].T
    if bnd_len is not None:
        big = -(2.0**64) + 1  # non-edge marker of lemon.py:32
        A[:K, K:] = big
        A[K:, :K] = big
        idx = np.arange(K)
        A[idx, K + idx] = -bnd_len
        A[K + idx, idx] = -bnd_len
        A[K:, K:] = -0.0
    return A


def solve_matching(K: int, pair_len: np.ndarray, bnd_len: Optional[np.ndarray], backend="native"):
    """Minimum-weight perfect matching of one matching graph on the host.
    Returns [(u, v)] over odd-cube indices, v == -1 for "matched to its connector"."""
    if K == 0:
        return []
    if backend in ("native", "rustworkx", "retworkx", "lemon"):  # the compiled backends of the reference
        from . import mwpm_native

        mate = mwpm_native.min_weight_perfect_matching(K, pair_len, bnd_len)
    elif backend == "networkx":
        import networkx as nx

        G = nx.Graph()
        pos = 0
        for a in range(K - 1):
            for b in range(a + 1, K):
                G.add_edge(a, b, inverse_weight=-pair_len[pos])
                pos += 1
        if bnd_len is not None:
            for a in range(K):
                G.add_edge(a, K + a, inverse_weight=-bnd_len[a])
            for a, b in it.combinations(range(K, 2 * K), 2):
                G.add_edge(a, b, inverse_weight=-0.0)
        elif K == 1:
            raise ValueError("a single odd cube cannot be matched without boundary points")
        mate = {}
        for u, v in nx.max_weight_matching(G, maxcardinality=True, weight="inverse_weight"):
            mate[u], mate[v] = v, u
    else:
        raise ValueError(f"unknown matching backend {backend!r}")
    out = []
    for u in range(K):
        v = mate[u]
        if v >= K:
            out.append((u, -1))
        elif u < v:
            out.append((u, v))
    return out


class GpuMatchingGraph:
    """The reference's ``MatchingGraph`` plug-in contract (``decoders/mwpm/matching.py:36-173``) over
    GPU-built lengths.

    ``GpuMatchingGraph(ec, code)`` builds itself from the code's current realisation exactly like
    the reference's backends do (``MatchingGraph.__init__`` -> ``with_edges_from``): it reads the
    ``weight`` and ``bit_val`` node attributes that the noise model and ``assign_weights`` left on
    ``code.graph`` and runs parity -> odd cubes -> shortest paths on the device
    (``fpb_run_weighted``).  So the class can be handed to ``build_match_graph(code, ec, <class>)``
    (``mwpm/algos.py:29-47``), and ``mwpm_decoder``'s loop over ``edge_path`` /
    ``stab_graph.edge_data(...)["common_vertex"]`` (``algos.py:88-95``) runs unchanged on it.

    Nodes are the odd ``Stabilizer`` objects and, with boundary points, one virtual node
    ``(boundary point, i)`` per odd cube (``matching.py:159-168``); virtual-virtual edges weigh 0.
    """

    backend = "native"  # host matcher used by min_weight_perfect_matching ("native" | "networkx")

    def __init__(self, ec, code=None, backend=None):
        self.ec = ec
        self.code = code
        self.virtual_points: List = []
        self.K = 0
        self._pair = np.empty(0)
        self._bnd = None
        self._side = None
        self._extra = {}
        self._stabs = []
        self._index = {}
        self._eng = None
        self._rerun = None
        if backend is not None:
            self.backend = backend if isinstance(backend, str) else "native"
        if code is not None:
            self.with_edges_from(code, ec)

    # -- construction --------------------------------------------------------------------
    def with_edges_from(self, code, ec):
        """``MatchingGraph.with_edges_from`` (``matching.py:97-123``) for the code's current
        ``weight`` / ``bit_val`` attributes."""
        g = code.graph
        if g.weight is None or g.bit_val is None:
            raise ValueError("no weights / bit values on the code: apply a noise model and assign_weights first")
        qn = np.concatenate([code.lattice.complexes[e].qubits for e in code.ec])
        w = np.array(g.weight[qn], np.float64)
        b = np.array(g.bit_val[qn])
        ct = code.lattice.complexes[ec]
        sl = slice(code.q_offset(ec), code.q_offset(ec) + ct.Q)
        if np.isnan(w[sl]).any() or (b[sl] < 0).any():
            raise ValueError(f"syndrome qubits of the {ec} complex lack a weight or a bit value")
        w = np.where(np.isnan(w), 1.0, w)  # other complexes: not part of this graph
        b = (b > 0).astype(np.uint8)
        eng = code.graph_engine()
        batch = eng.run_weighted(w[None, :], b[None, :])
        self._populate(code, ec, eng, batch, 0)
        self._rerun = (w, b)
        return self

    @classmethod
    def from_batch(cls, code, ec, eng, batch, t, backend="native"):
        """Trial ``t`` of a batch the engine has just run (the batched drivers' shortcut)."""
        self = cls(ec, None, backend=backend)
        self.code = code
        self._populate(code, ec, eng, batch, t)
        return self

    def _populate(self, code, ec, eng, batch, t):
        self.code = code
        ct = code.lattice.complexes[ec]
        cb = batch.complexes[ec]
        self._eng, self._t, self._ci, self._run_id = eng, t, code.ec.index(ec), eng.run_id
        odd = cb.odd(t)
        self.K = len(odd)
        self._odd = odd
        stabs = getattr(code, ec + "_stabilizers")
        self._stabs = [stabs[int(c)] for c in odd]
        self._index = {s: i for i, s in enumerate(self._stabs)}
        self._pair = cb.pairs(t)
        self.virtual_points = []
        self._bnd = self._side = None
        if ct.has_boundary and self.K:
            self._bnd, self._side = cb.boundary(t)
            # the virtual node of cube i is named after the boundary point its connector path ends in
            self._conn = eng.paths_list([t] * self.K, [self._ci] * self.K, list(range(self.K)), [-1] * self.K)
            g = code.graph
            self.virtual_points = [(g.to_points[int(ct.qubits[p[-1]])], i) for i, p in enumerate(self._conn)]
            self._index.update({vp: self.K + i for i, vp in enumerate(self.virtual_points)})

    def _engine(self):
        """The engine with this graph's realisation as its last run (re-run if it moved on)."""
        if self._eng.run_id != self._run_id:
            if self._rerun is None:
                raise RuntimeError("the engine has run another batch since this graph was built")
            self._eng.run_weighted(self._rerun[0][None, :], self._rerun[1][None, :], fetch=False)
            self._run_id, self._t = self._eng.run_id, 0
        return self._eng

    # -- MatchingGraph interface ------------------------------------------------------
    def add_edge(self, edge, weight, path=()):
        self._extra[frozenset(edge)] = (weight, list(path))

    def _ij(self, edge):
        return self._index[edge[0]], self._index[edge[1]]

    def edge_weight(self, edge):
        if frozenset(edge) in self._extra:
            return self._extra[frozenset(edge)][0]
        i, j = sorted(self._ij(edge))
        K = self.K
        if j < K:
            w = float(self._pair[tri_index(i, j, K)])
        elif i >= K:
            return 0
        elif j == K + i:
            w = float(self._bnd[i])
        else:
            raise KeyError(edge)
        return int(w) if self._eng is not None and self._eng.lengths == "rustworkx" else w

    def edge_path(self, edge):
        """The path of stabilizer-graph nodes behind a matching-graph edge, in the form
        ``mwpm_decoder`` consumes (``algos.py:88-95``): ``[stab1, ..., stab2]`` between two odd
        cubes, ``[boundary point, cube, ..., cube_i]`` for cube ``i`` and its virtual node
        (``matching.py:157-160``: the connector path without 'low' / 'high'), ``[]`` between two
        virtual nodes."""
        if frozenset(edge) in self._extra:
            return self._extra[frozenset(edge)][1]
        i, j = sorted(self._ij(edge))
        if i >= self.K:
            return []
        code, ct = self.code, self.code.lattice.complexes[self.ec]
        stabs = getattr(code, self.ec + "_stabilizers")
        other = code.qubit_cubes(self.ec)
        if j >= self.K:
            if j != self.K + i:
                raise KeyError(edge)
            qs = self._conn[i]
            cube, nodes = int(self._odd[i]), []
            for q in qs[:-1]:
                nodes.append(stabs[cube])
                a, b = other[q]
                cube = int(b if a == cube else a)
            nodes.append(stabs[cube])
            nodes.append(code.graph.to_points[int(ct.qubits[qs[-1]])])
            return nodes[::-1]
        qs = self._engine().paths_list([self._t], [self._ci], [i], [j])[0]
        cube, nodes = int(self._odd[j]), []
        for q in qs:
            nodes.append(stabs[cube])
            a, b = other[q]
            cube = int(b if a == cube else a)
        nodes.append(stabs[cube])
        return nodes[::-1]

    def min_weight_perfect_matching(self):
        pairs = solve_matching(self.K, self._pair, self._bnd, backend=self.backend)
        out = []
        used = set()
        for u, v in pairs:
            if v >= 0:
                out.append((self._stabs[u], self._stabs[v]))
            else:
                out.append((self._stabs[u], self.virtual_points[u]))
                used.add(u)
        rest = [i for i in range(len(self.virtual_points)) if i not in used]
        out += [(self.virtual_points[a], self.virtual_points[b]) for a, b in zip(rest[0::2], rest[1::2])]
        return out

    def total_weight_of(self, matching):
        return sum(self.edge_weight(e) for e in matching)

    def number_of_edges(self):
        K = self.K
        nb = len(self.virtual_points)
        return K * (K - 1) // 2 + nb + nb * (nb - 1) // 2

    def to_nx(self):
        import networkx as nx

        G = nx.Graph()
        nodes = self._stabs + self.virtual_points
        for a, b in it.combinations(range(len(nodes)), 2):
            try:
                w = self.edge_weight((nodes[a], nodes[b]))
            except KeyError:
                continue
            G.add_edge(nodes[a], nodes[b], weight=w, inverse_weight=-w)
        return G
