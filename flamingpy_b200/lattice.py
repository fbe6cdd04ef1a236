"""Closed-form RHG lattice tables for the Monte-Carlo EC trial path.

The reference builds the lattice as a networkx graph of coordinate tuples
(``flamingpy/codes/surface_code.py:115-245``), indexes nodes by sorting the
tuples (``codes/graphs/egraph.py:133-155``), finds stabilizers by set
intersection (``surface_code.py:396-487``) and links them with an O(S^2)
pairwise test (``codes/graphs/stabilizer_graph.py:289-302``).  Here the same
structures are produced directly as flat integer arrays from coordinate
arithmetic, in the layouts the CUDA kernels stage into shared memory:

* nodes are the integer points with one (dual node) or two (primal node) odd
  coordinates inside the per-axis range the boundary type allows, indexed in
  lexicographic (x, y, z) order;
* every error complex ("primal" / "dual") is a 3-D box of cubes in
  ``itertools.product`` order; cube ``c`` has six face slots
  ``(-x, +x, -y, +y, -z, +z)`` holding the syndrome qubit on that face and the
  cube (or LOW / HIGH boundary connector) on the other side.  Each syndrome
  qubit is exactly one edge of the stabilizer graph.

Nothing here imports the reference; ``tests/test_lattice.py`` checks the
tables against the live reference where it is importable.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

BOUNDARY_NAMES = {
    "open_primal": ["primal", "dual", "primal"],
    "open_dual": ["primal", "dual", "dual"],
    "primal": ["primal", "primal", "primal"],
    "dual": ["dual", "dual", "dual"],
    "periodic": ["periodic", "periodic", "periodic"],
    "periodic_primal": ["periodic", "periodic", "primal"],
    "periodic_dual": ["periodic", "periodic", "dual"],
}

# Neighbour codes in ComplexTables.nbr
NBR_NONE = -1
# LOW = S, HIGH = S + 1 (per complex)

# Face kinds packed two bits per direction into ComplexTables.ninfo
FACE_REGULAR = 0  # neighbour cube at c +- stride
FACE_WRAP = 1  # neighbour cube across a periodic seam
FACE_NONE = 2  # no qubit on this face
FACE_BOUNDARY = 3  # qubit exists, other side is the LOW (-) / HIGH (+) connector


def bound_str_for(boundaries: str, ec: str) -> str:
    """Boundary name the reference derives in ``SurfaceCode.__init__``
    (``surface_code.py:333-339``)."""
    if boundaries == "open":
        return "open_dual" if ec in ("primal", "both") else "open_primal"
    if boundaries == "toric":
        return "periodic_dual" if ec in ("primal", "both") else "periodic_primal"
    return boundaries


@dataclass
class ComplexTables:
    """Static tables of one error complex (one stabilizer graph)."""

    ec: str
    cube_dims: Tuple[int, int, int]  # number of cubes per axis (nx, ny, nz)
    S: int
    qubits: np.ndarray  # int32[Q] node index of each syndrome qubit, ascending
    cube_qubit: np.ndarray  # int32[S, 6] local qubit id (into ``qubits``) per face, -1 if absent
    nbr: np.ndarray  # int32[S, 6] cube on the other side, S=LOW, S+1=HIGH, -1 none
    ninfo: np.ndarray  # uint16[S] six 2-bit face kinds
    bound_axis: int  # axis carrying LOW/HIGH boundary points, -1 if none
    low_points: np.ndarray  # int32[B/2] node index, reference order not preserved (sorted)
    high_points: np.ndarray
    planes: List[np.ndarray]  # per checked plane: local qubit ids (check_correction)
    plane_names: List[str]
    # slot layout of the per-trial weight array staged in shared memory
    slot_dims: Tuple[int, int, int] = (0, 0, 0)
    n_slots: int = 0
    slot_qubit: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))
    sbase: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))  # int32[S] 3*padded idx
    d_off: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))  # int32[6,2] cube offset
    s_off: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))  # int32[6,2] slot offset
    low_seed_cube: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))
    low_seed_slot: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))
    high_seed_cube: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))
    high_seed_slot: np.ndarray = field(default_factory=lambda: np.empty(0, np.int32))

    @property
    def Q(self) -> int:
        return int(self.qubits.shape[0])

    @property
    def has_boundary(self) -> bool:
        return self.bound_axis >= 0


@dataclass
class LatticeTables:
    """Everything static about one code instance, as flat arrays."""

    dims: Tuple[int, int, int]
    boundaries: List[str]
    bound_str: str
    ec: List[str]
    coords: np.ndarray  # int32[N, 3], lexicographically sorted
    node_type: np.ndarray  # uint8[N], 0 primal, 1 dual
    adj_indptr: np.ndarray  # int32[N + 1]
    adj_indices: np.ndarray  # int32[nnz], ascending within a row
    adj_vals: np.ndarray  # int8[nnz] (polarity sign, 1 by default)
    complexes: Dict[str, ComplexTables]

    @property
    def N(self) -> int:
        return int(self.coords.shape[0])

    @property
    def degree(self) -> np.ndarray:
        return np.diff(self.adj_indptr).astype(np.int32)

    def coord_index(self) -> Dict[Tuple[int, int, int], int]:
        return {tuple(int(v) for v in c): i for i, c in enumerate(self.coords)}


def _axis_range(d: int, typ: str) -> Tuple[int, int]:
    """Inclusive coordinate range of one axis (SURVEY Appendix A.1;
    ``surface_code.py:172-216``)."""
    if typ == "periodic":
        return 0, 2 * d - 1
    if typ == "primal":
        return 0, 2 * d - 2
    if typ == "dual":
        return 1, 2 * d - 1
    raise ValueError(f"unknown boundary type {typ!r}")


def _polarity_sign(p1: np.ndarray, p2: np.ndarray) -> np.ndarray:
    """Vectorised ``alternating_polarity`` (``surface_code.py:26-53``) for edges
    given as (primal vertex, neighbour) coordinate arrays."""
    diff = p2 - p1
    direction = np.argmax(diff != 0, axis=1)
    src = np.where(direction == 0, p1[:, 1], np.where(direction == 1, p1[:, 2], p1[:, 0]))
    return np.where(src % 2 == 0, 1, -1).astype(np.int8)


def _reference_graph_order(dims, btypes):
    """Insertion order of the reference's ``RHG_graph`` (``surface_code.py:172-243``): the rank of
    every node and, per node, its neighbours in the order their edges were added.  The reference fills
    its graph while iterating a Python *set* of primal vertices, so these orders are a property of
    CPython's tuple hashing; they are re-derived here with real sets.  Two things depend on them: which
    of the two qubits shared by a size-2 periodic cube pair becomes the edge (``_double_edge_keep``),
    and the order of the micronodes inside each macronode of ``CVMacroLayer`` (``macro_tables``).
    Returns (rank: {node: int}, nbrs: {node: [neighbour, ...]}, seam: set of frozenset edges that
    cross a periodic boundary)."""
    import itertools as it

    d = [int(v) for v in dims]
    rng = [range(d[a] - (1 if btypes[a] == "primal" else 0)) for a in range(3)]
    faces = ((0, 1, 1), (1, 0, 1), (1, 1, 0), (2, 1, 1), (1, 2, 1), (1, 1, 2))
    primal_vertices = {(2 * i + f[0], 2 * j + f[1], 2 * k + f[2]) for (i, j, k) in it.product(*rng) for f in faces}

    def outside(p):
        for a in range(3):
            if btypes[a] == "dual" and p[a] in (0, 2 * d[a]):
                return True
            if btypes[a] == "periodic" and p[a] == 2 * d[a]:
                return True
        return False

    def four_neighbours(p):
        x, y, z = p
        if z % 2:
            if x % 2:
                return [(x, y, z - 1), (x - 1, y, z), (x, y, z + 1), (x + 1, y, z)]
            return [(x, y, z - 1), (x, y + 1, z), (x, y, z + 1), (x, y - 1, z)]
        return [(x, y - 1, z), (x - 1, y, z), (x, y + 1, z), (x + 1, y, z)]

    rank, nbrs, seam = {}, {}, set()

    def node(p):
        if p not in rank:
            rank[p] = len(rank)
            nbrs[p] = []

    def edge(u, v):
        if v not in nbrs[u]:
            nbrs[u].append(v)
        if u not in nbrs[v]:
            nbrs[v].append(u)

    for v in primal_vertices:
        if outside(v):
            continue
        for nb in four_neighbours(v):
            if outside(nb):
                continue
            node(v)
            node(nb)
            edge(v, nb)
            for a in range(3):
                if btypes[a] == "periodic" and nb[a] == 0:
                    other = list(nb)
                    other[a] = 2 * d[a] - 1
                    other = tuple(other)
                    node(other)
                    edge(nb, other)
                    seam.add(frozenset((nb, other)))
    return rank, nbrs, seam


def _reference_node_order(dims, btypes):
    return _reference_graph_order(dims, btypes)[0]


def macro_tables(lat: "LatticeTables"):
    """Static tables of the macronode lattice of ``CVMacroLayer`` (``noise/cv.py:367-375``,
    ``codes/graphs/egraph.py:26-103,133-155``): every node of the canonical lattice becomes a
    macronode of exactly four micronodes, one per incident edge (each edge a two-micronode
    'dumbbell') plus padding micronodes without a partner at the boundary.  Micronode ``4 * i + j``
    is the ``j``-th micronode of canonical node ``i`` in the reference's order - the order in which
    the edges at that node come up in ``can_graph.edges`` (insertion order of ``RHG_graph``, see
    ``_reference_graph_order``), then the padding.  Returns (partner int32[4N]: the micronode at the
    other end of the dumbbell or -1, sign int8[4N]: the edge's polarity weight)."""
    rank, nbrs, _ = _reference_graph_order(lat.dims, lat.boundaries)
    index = lat.coord_index()
    if set(rank) != set(index):
        raise AssertionError("reference graph order: node set differs from the lattice tables")
    N = lat.N
    slots = [[] for _ in range(N)]  # per canonical node: the canonical neighbour behind each micronode
    seen = set()
    for n in sorted(rank, key=rank.get):  # networkx EdgeView: nodes in insertion order, skip seen ends
        for nb in nbrs[n]:
            if nb not in seen:
                slots[index[n]].append(index[nb])
                slots[index[nb]].append(index[n])
        seen.add(n)
    partner = np.full(4 * N, -1, np.int32)
    sign = np.zeros(4 * N, np.int8)
    # polarity weights of the canonical edges
    weight = {}
    for i in range(N):
        for e in range(lat.adj_indptr[i], lat.adj_indptr[i + 1]):
            weight[(i, int(lat.adj_indices[e]))] = int(lat.adj_vals[e])
    for i in range(N):
        if len(slots[i]) > 4:
            raise AssertionError("a lattice node with more than four neighbours")
        for j, nb in enumerate(slots[i]):
            partner[4 * i + j] = 4 * nb + slots[nb].index(i)
            sign[4 * i + j] = weight[(i, nb)]
    return partner, sign


def _double_edge_keep(ec, dims, btypes, cidx, start, pairs):
    """Which of the two qubits shared by a doubly-connected cube pair the reference keeps as the
    edge's ``common_vertex``: ``(set(stab1.coords()) & set(stab2.coords())).pop()``
    (``stabilizer_graph.py:295-298``) - whatever CPython's set iteration yields.  Reproduced by
    running the same set operations on the same coordinate tuples in the same insertion orders
    (``surface_code.py:413-451``; networkx ``subgraph`` keeps ``set(nodes)``).
    ``pairs``: iterable of (cube1 < cube2).  Returns {(cube1, cube2): kept coordinate tuple}."""
    d = [int(v) for v in dims]
    rank = _reference_node_order(dims, btypes)
    n_graph = len(rank)
    periodic_axes = [a for a in range(3) if btypes[a] == "periodic"]
    cache = {}

    def coords_of(s):
        if s in cache:
            return cache[s]
        i, j, k = (int(v) for v in (cidx[s] + start))
        if ec == "primal":
            body = [(2 * i, 2 * j + 1, 2 * k + 1), (2 * i + 1, 2 * j, 2 * k + 1), (2 * i + 1, 2 * j + 1, 2 * k),
                    (2 * i + 2, 2 * j + 1, 2 * k + 1), (2 * i + 1, 2 * j + 2, 2 * k + 1), (2 * i + 1, 2 * j + 1, 2 * k + 2)]
        else:
            body = [(2 * i + 1, 2 * j + 2, 2 * k + 2), (2 * i + 2, 2 * j + 1, 2 * k + 2), (2 * i + 2, 2 * j + 2, 2 * k + 1),
                    (2 * i + 3, 2 * j + 2, 2 * k + 2), (2 * i + 2, 2 * j + 3, 2 * k + 2), (2 * i + 2, 2 * j + 2, 2 * k + 3)]
        members = set(body)
        found = set()
        for p in sorted((p for p in members if p in rank), key=rank.get):  # intersection iterates the graph
            found.add(p)
        actual = list(found)
        if len(actual) < 6:
            for a in periodic_axes:
                if ec == "dual":
                    p = list(body[3 + a])
                    if p[a] == 1:
                        p[a] = 2 * d[a] - 1
                        actual.append(tuple(p))
                else:
                    p = list(body[a])
                    if p[a] == 2 * d[a] - 2:
                        p[a] = 0
                        actual.append(tuple(p))
        shown = set(p for p in actual if p in rank)
        if 2 * len(shown) < n_graph:
            out = [p for p in shown]
        else:
            out = sorted(shown, key=rank.get)
        cache[s] = out
        return out

    keep = {}
    for s1, s2 in pairs:
        common = set(coords_of(s1)) & set(coords_of(s2))
        keep[(s1, s2)] = common.pop()
    return keep


def build_lattice(
    distance,
    ec="primal",
    boundaries: str = "open",
    polarity: bool = False,
) -> LatticeTables:
    """Build all static tables for ``SurfaceCode(distance, ec, boundaries)``.

    ``ec`` may be "primal", "dual" or "both" (both = the extension the
    reference disables at ``surface_code.py:323-331``; only meaningful for
    periodic boundaries where no perfect-qubit handling is needed).
    ``boundaries`` is "open", "toric", "periodic", or any name in
    ``BOUNDARY_NAMES`` (used by tests to sweep all per-axis combinations).
    ``polarity=True`` applies the reference's ``alternating_polarity`` signs.
    """
    if np.ndim(distance) == 0:
        dims = (int(distance),) * 3
    else:
        dims = tuple(int(v) for v in distance)
        if len(dims) != 3:
            raise ValueError("dims must be an integer or a sequence of three integers.")
    ec_list = ["primal", "dual"] if ec == "both" else [ec]
    if isinstance(boundaries, str):
        if boundaries in ("open", "toric", "periodic"):
            bound_str = bound_str_for(boundaries, ec)
        else:
            bound_str = boundaries
        if bound_str not in BOUNDARY_NAMES:
            raise ValueError(f"unknown boundaries {boundaries!r}")
        btypes = list(BOUNDARY_NAMES[bound_str])
    else:
        btypes = list(boundaries)
        bound_str = ",".join(btypes)

    lo = np.array([_axis_range(dims[a], btypes[a])[0] for a in range(3)])
    hi = np.array([_axis_range(dims[a], btypes[a])[1] for a in range(3)])
    per = np.array([t == "periodic" for t in btypes])
    size2 = 2 * np.array(dims)

    # ---- candidate points in range, classified by number of odd coordinates ----
    axes = [np.arange(lo[a], hi[a] + 1) for a in range(3)]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    n_odd = (grid % 2).sum(axis=1)
    primal_pts = grid[n_odd == 2]

    # ---- edges: each primal point to +-1 along its two odd axes ----
    e_src, e_dst = [], []
    for a in range(3):
        for sgn in (-1, 1):
            sel = primal_pts[:, a] % 2 == 1
            src = primal_pts[sel]
            dst = src.copy()
            dst[:, a] += sgn
            if per[a]:
                dst[:, a] %= size2[a]
            ok = (dst[:, a] >= lo[a]) & (dst[:, a] <= hi[a])
            e_src.append(src[ok])
            e_dst.append(dst[ok])
    e_src = np.concatenate(e_src)
    e_dst = np.concatenate(e_dst)

    def key(p):
        return (p[:, 0].astype(np.int64) * (size2[1] + 1) + p[:, 1]) * (size2[2] + 1) + p[:, 2]

    k_src, k_dst = key(e_src), key(e_dst)
    node_keys = np.unique(np.concatenate([k_src, k_dst]))  # sorted == lexicographic order
    N = node_keys.shape[0]
    coords = np.empty((N, 3), np.int32)
    coords[:, 2] = node_keys % (size2[2] + 1)
    coords[:, 1] = (node_keys // (size2[2] + 1)) % (size2[1] + 1)
    coords[:, 0] = node_keys // ((size2[2] + 1) * (size2[1] + 1))
    node_type = np.where((coords % 2).sum(axis=1) == 2, 0, 1).astype(np.uint8)

    i_src = np.searchsorted(node_keys, k_src).astype(np.int32)
    i_dst = np.searchsorted(node_keys, k_dst).astype(np.int32)
    if polarity:
        # the reference evaluates polarity((vertex, neighbour)) for in-range
        # neighbours and polarity((neighbour, other_side)) across a periodic seam
        # (surface_code.py:218-241)
        seam = np.zeros(len(e_src), bool)
        for a in range(3):
            if per[a]:
                seam |= (np.abs(e_dst[:, a] - e_src[:, a]) > 1)
        sign = _polarity_sign(e_src, e_dst)
        if seam.any():
            # across the seam the reference passes (dual node at 0, primal node at 2d-1)
            sign_seam = _polarity_sign(e_dst[seam], e_src[seam])
            sign = sign.copy()
            sign[seam] = sign_seam
    else:
        sign = np.ones(len(e_src), np.int8)
    rows = np.concatenate([i_src, i_dst])
    cols = np.concatenate([i_dst, i_src])
    vals = np.concatenate([sign, sign])
    # drop duplicate edges (d=1-style degenerate wraps) keeping one
    rc = rows.astype(np.int64) * N + cols
    rc_u, first = np.unique(rc, return_index=True)
    rows, cols, vals = rows[first], cols[first], vals[first]
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    indptr = np.zeros(N + 1, np.int32)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr).astype(np.int32)

    def lookup(pts: np.ndarray) -> np.ndarray:
        """Node index of each coordinate row, -1 where no such node."""
        inr = np.all((pts >= lo) & (pts <= hi), axis=1)
        k = key(np.where(inr[:, None], pts, 0))
        pos = np.searchsorted(node_keys, k)
        pos = np.minimum(pos, N - 1)
        found = inr & (node_keys[pos] == k)
        return np.where(found, pos, -1).astype(np.int32)

    complexes: Dict[str, ComplexTables] = {}
    for e in ec_list:
        complexes[e] = _build_complex(e, dims, btypes, bound_str, coords, lookup)

    return LatticeTables(
        dims=dims,
        boundaries=btypes,
        bound_str=bound_str,
        ec=ec_list,
        coords=coords,
        node_type=node_type,
        adj_indptr=indptr,
        adj_indices=cols.astype(np.int32),
        adj_vals=vals.astype(np.int8),
        complexes=complexes,
    )


def _build_complex(ec, dims, btypes, bound_str, coords, lookup) -> ComplexTables:
    dims = np.array(dims)
    per = np.array([t == "periodic" for t in btypes])
    size2 = 2 * dims
    if ec == "primal":
        # surface_code.py:172-188: i < d - [axis has primal boundary]; centre 2i+1
        n_cubes = dims - np.array([1 if t == "primal" else 0 for t in btypes])
        start = np.zeros(3, int)
        centre0 = 1
    else:
        # surface_code.py:466-487: i in [-1 (primal/periodic) or 0 (dual), d-2]; centre 2i+2
        start = np.array([0 if t == "dual" else -1 for t in btypes])
        n_cubes = (dims - 1) - start
        centre0 = 2
    nx, ny, nz = (int(v) for v in n_cubes)
    S = nx * ny * nz
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cidx = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)  # product order
    centres = 2 * (cidx + start) + centre0

    face_node = np.full((S, 6), -1, np.int32)
    nbr = np.full((S, 6), NBR_NONE, np.int32)
    strides = np.array([ny * nz, nz, 1])
    for a in range(3):
        for s, sgn in enumerate((-1, 1)):
            d = 2 * a + s
            pts = centres.copy()
            pts[:, a] += sgn
            if per[a]:
                pts[:, a] %= size2[a]
            face_node[:, d] = lookup(pts)
            # neighbour cube on the other side of that face
            ni = cidx[:, a] + sgn
            if per[a]:
                wrapped = ni % n_cubes[a]
                nb = np.arange(S) + (wrapped - cidx[:, a]) * strides[a]
                nbr[:, d] = nb
            else:
                inside = (ni >= 0) & (ni < n_cubes[a])
                nbr[:, d] = np.where(inside, np.arange(S) + sgn * strides[a], NBR_NONE)
    # no qubit on a face -> no edge
    nbr[face_node < 0] = NBR_NONE

    # boundary points (surface_code.py:489-525): only for non-periodic codes, on
    # the first axis whose boundary type equals ec
    bound_axis = -1
    low_pts = np.empty(0, np.int32)
    high_pts = np.empty(0, np.int32)
    if "periodic" not in bound_str:
        where = [a for a in range(3) if btypes[a] == ec]
        if where:
            bound_axis = where[0]
            low_c = 0 if ec == "primal" else 1
            high_c = 2 * dims[bound_axis] - 2 if ec == "primal" else 2 * dims[bound_axis] - 1
            dl, dh = 2 * bound_axis, 2 * bound_axis + 1
            is_low = (face_node[:, dl] >= 0) & (coords[np.maximum(face_node[:, dl], 0), bound_axis] == low_c)
            is_high = (face_node[:, dh] >= 0) & (
                coords[np.maximum(face_node[:, dh], 0), bound_axis] == high_c
            )
            nbr[is_low, dl] = S
            nbr[is_high, dh] = S + 1
            low_pts = np.sort(face_node[is_low, dl])
            high_pts = np.sort(face_node[is_high, dh])

    qubits = np.unique(face_node[face_node >= 0]).astype(np.int32)
    cube_qubit = np.where(face_node >= 0, np.searchsorted(qubits, np.maximum(face_node, 0)), -1)
    cube_qubit = cube_qubit.astype(np.int32)

    # every qubit is in one or two cubes; exactly one <=> boundary point
    counts = np.bincount(cube_qubit[cube_qubit >= 0], minlength=len(qubits))
    dangling = set(qubits[counts == 1].tolist())
    bset = set(low_pts.tolist()) | set(high_pts.tolist())
    standard = bound_str in ("open_dual", "open_primal", "periodic", "periodic_dual", "periodic_primal")
    if standard and not (S > 0 and counts.max() <= 2 and dangling == bset):
        if not (min(dims) <= 2 and per.any()):  # d=2 periodic: cubes share two faces
            raise AssertionError("lattice invariant violated: dangling qubits != boundary points")

    # size-2 periodic axes: the two cubes of such an axis share two qubits, but the reference's
    # stabilizer graph holds one edge per cube pair (stabilizer_graph.py:295-298) - the other
    # qubit still counts for the parities, it just is no edge
    for a in range(3):
        if per[a] and n_cubes[a] == 2:
            first = np.nonzero(cidx[:, a] == 0)[0]
            pairs = [(int(c), int(c + strides[a])) for c in first]
            keep = _double_edge_keep(ec, dims, btypes, cidx, start, pairs)
            for c, c2 in pairs:
                inner, seam = face_node[c, 2 * a + 1], face_node[c, 2 * a]
                kept = tuple(int(v) for v in keep[(c, c2)])
                if kept == tuple(int(v) for v in coords[inner]):
                    nbr[c, 2 * a] = nbr[c2, 2 * a + 1] = NBR_NONE
                elif kept == tuple(int(v) for v in coords[seam]):
                    nbr[c, 2 * a + 1] = nbr[c2, 2 * a] = NBR_NONE
                else:
                    raise AssertionError("double edge: kept qubit is neither shared face")

    # face kinds
    ninfo = np.zeros(S, np.uint16)
    for a in range(3):
        for s in range(2):
            d = 2 * a + s
            kind = np.full(S, FACE_NONE, np.uint16)
            has = face_node[:, d] >= 0
            nb = nbr[:, d]
            regular = has & (nb >= 0) & (nb < S) & (nb - np.arange(S) == (1 if s else -1) * strides[a])
            wrap = has & (nb >= 0) & (nb < S) & ~regular
            bnd = has & (nb >= S)
            kind[regular] = FACE_REGULAR
            kind[wrap] = FACE_WRAP
            kind[bnd] = FACE_BOUNDARY
            ninfo |= (kind << np.uint16(2 * d)).astype(np.uint16)

    # check_correction planes (decoder.py:186-214)
    if bound_str == "periodic":
        planes_to_check = [0, 1, 2]
    elif bound_str in ("periodic_primal", "periodic_dual"):
        planes_to_check = [0, 1]
    elif bound_str.startswith("open"):
        planes_to_check = [0] if ec == "primal" else [1]
    else:
        planes_to_check = []
    sheet = 0 if ec == "primal" else 1
    qcoords = coords[qubits]
    planes = [np.nonzero(qcoords[:, a] == sheet)[0].astype(np.int32) for a in planes_to_check]

    ct = ComplexTables(
        ec=ec,
        cube_dims=(nx, ny, nz),
        S=S,
        qubits=qubits,
        cube_qubit=cube_qubit,
        nbr=nbr,
        ninfo=ninfo,
        bound_axis=bound_axis,
        low_points=low_pts.astype(np.int32),
        high_points=high_pts.astype(np.int32),
        planes=planes,
        plane_names=["xyz"[a] for a in planes_to_check],
    )
    _build_slots(ct)
    return ct


def _build_slots(ct: ComplexTables) -> None:
    """Weight-slot layout used by the shortest-path kernel.

    Slot ``3*p + a`` holds the weight of the +a face of the cube with padded
    linear index ``p``; the grid is padded by one layer on the low side of the
    boundary axis so that LOW-connector faces (the -a0 face of cubes at index 0)
    get a slot of their own.  The -a face of cube c is the +a face of the cube
    one step down, so every face kind reduces to ``sbase[c] + const``.
    """
    nx, ny, nz = ct.cube_dims
    S = ct.S
    pad = [0, 0, 0]
    if ct.bound_axis >= 0:
        pad[ct.bound_axis] = 1
    PX, PY, PZ = nx + pad[0], ny + pad[1], nz + pad[2]
    ct.slot_dims = (PX, PY, PZ)
    ct.n_slots = 3 * PX * PY * PZ
    pstr = np.array([PY * PZ, PZ, 1])
    cstr = np.array([ny * nz, nz, 1])
    n = np.array([nx, ny, nz])
    c = np.arange(S)
    i, j, k = c // (ny * nz), (c // nz) % ny, c % nz
    pidx = ((i + pad[0]) * PY + (j + pad[1])) * PZ + (k + pad[2])
    ct.sbase = (3 * pidx).astype(np.int32)

    d_off = np.zeros((6, 2), np.int32)
    s_off = np.zeros((6, 2), np.int32)
    for a in range(3):
        # - direction: regular -> c - stride, slot of the lower cube's + face
        d_off[2 * a, 0] = -cstr[a]
        s_off[2 * a, 0] = -3 * pstr[a] + a
        # - direction across the seam: neighbour at index n-1
        d_off[2 * a, 1] = (n[a] - 1) * cstr[a]
        s_off[2 * a, 1] = 3 * (n[a] - 1) * pstr[a] + a
        # + direction: own + face either way
        d_off[2 * a + 1, 0] = cstr[a]
        s_off[2 * a + 1, 0] = a
        d_off[2 * a + 1, 1] = -(n[a] - 1) * cstr[a]
        s_off[2 * a + 1, 1] = a
    ct.d_off, ct.s_off = d_off, s_off

    slot_qubit = np.full(ct.n_slots, -1, np.int32)
    for d in range(6):
        a = d // 2
        kind = (ct.ninfo >> np.uint16(2 * d)) & 3
        has = kind != FACE_NONE
        typ = np.where(kind == FACE_WRAP, 1, 0)
        slot = ct.sbase + s_off[d][typ]
        q = ct.cube_qubit[:, d]
        sl, qq = slot[has], q[has]
        prev = slot_qubit[sl]
        if np.any((prev >= 0) & (prev != qq)):
            raise AssertionError("slot collision with different qubits")
        slot_qubit[sl] = qq
    ct.slot_qubit = slot_qubit
    # consistency: walking a regular/wrap face lands on nbr
    for d in range(6):
        kind = (ct.ninfo >> np.uint16(2 * d)) & 3
        for typ, code in ((0, FACE_REGULAR), (1, FACE_WRAP)):
            m = kind == code
            if np.any(c[m] + d_off[d, typ] != ct.nbr[m, d]):
                raise AssertionError("implicit neighbour arithmetic disagrees with nbr table")

    if ct.bound_axis >= 0:
        dl, dh = 2 * ct.bound_axis, 2 * ct.bound_axis + 1
        lm = ct.nbr[:, dl] == S
        hm = ct.nbr[:, dh] == S + 1
        ct.low_seed_cube = c[lm].astype(np.int32)
        ct.low_seed_slot = (ct.sbase[lm] + s_off[dl, 0]).astype(np.int32)
        ct.high_seed_cube = c[hm].astype(np.int32)
        ct.high_seed_slot = (ct.sbase[hm] + s_off[dh, 0]).astype(np.int32)
