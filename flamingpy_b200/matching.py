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


def solve_matching(K: int, pair_len: np.ndarray, bnd_len: Optional[np.ndarray], backend="networkx"):
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
    """Reference ``MatchingGraph`` interface over GPU-built lengths."""

    def __init__(self, ec, code=None):
        self.ec = ec
        self.code = code
        self.virtual_points: List = []
        self.K = 0
        self._pair = np.empty(0)
        self._bnd = None
        self._extra = {}
        self._backend = "networkx"
        self._stabs = []

    @classmethod
    def from_batch(cls, code, ec, eng, batch, t, backend="networkx"):
        self = cls(ec, code)
        ct = code.lattice.complexes[ec]
        cb = batch.complexes[ec]
        self._eng, self._t, self._ci = eng, t, code.ec.index(ec)
        odd = cb.odd(t)
        self.K = len(odd)
        self._odd = odd
        stabs = getattr(code, ec + "_stabilizers")
        self._stabs = [stabs[int(c)] for c in odd]
        self._index = {s: i for i, s in enumerate(self._stabs)}
        self._pair = cb.pairs(t)
        if ct.has_boundary and self.K:
            self._bnd, side = cb.boundary(t)
            sg = getattr(code, ec + "_stab_graph")
            self.virtual_points = [(("low", "high")[int(s)], i) for i, s in enumerate(side)]
            self._index.update({vp: self.K + i for i, vp in enumerate(self.virtual_points)})
        self._backend = backend if isinstance(backend, str) else "networkx"
        return self

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
            return float(self._pair[tri_index(i, j, K)])
        if i >= K:
            return 0
        if j == K + i:
            return float(self._bnd[i])
        raise KeyError(edge)

    def edge_path(self, edge):
        """Syndrome qubits (coordinates) crossed by the correction of this edge."""
        if frozenset(edge) in self._extra:
            return self._extra[frozenset(edge)][1]
        i, j = sorted(self._ij(edge))
        if i >= self.K:
            return []
        dst = j if j < self.K else -1
        flips = self._eng.paths_toggle([self._t], [self._ci], [i], [dst])[self._t]
        ct = self.code.lattice.complexes[self.ec]
        off = self._eng.q_off[self.ec]
        g = self.code.graph
        return [g.to_points[int(ct.qubits[q])] for q in np.nonzero(flips[off : off + ct.Q])[0]]

    def min_weight_perfect_matching(self):
        pairs = solve_matching(self.K, self._pair, self._bnd, backend=self._backend)
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
