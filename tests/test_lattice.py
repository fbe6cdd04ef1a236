"""Closed-form lattice tables versus the reference's networkx construction.

Mirrors the size invariants of the reference's
``tests/qubit_codes/test_surface_code.py:265-304,406-459`` and, where the
reference is importable, checks node order, adjacency, stabilizer order and
membership, boundary points, stabilizer-graph edges and check planes exactly.
"""

import itertools as it

import numpy as np
import pytest

from flamingpy_b200.lattice import build_lattice, FACE_NONE

CODES = [
    (d, ec, b)
    for d in (2, 3, 4, (2, 3, 4), (4, 2, 3), (2, 2, 3), (3, 2, 2), (2, 4, 2))
    for ec in ("primal", "dual")
    for b in ("open", "toric", "periodic")
]


@pytest.mark.parametrize("d", [2, 3, 5, 9, 13])
def test_size_formulas_periodic(d):
    lat = build_lattice(d, "both", "periodic")
    assert lat.N == 6 * d**3
    assert len(lat.adj_indices) == 2 * 12 * d**3
    for ec in ("primal", "dual"):
        ct = lat.complexes[ec]
        assert ct.S == d**3
        assert ct.Q == 3 * d**3
        assert not ct.has_boundary
        assert len(ct.planes) == 3


@pytest.mark.parametrize("d", [2, 3, 5, 9, 13])
@pytest.mark.parametrize("ec", ["primal", "dual"])
def test_size_formulas_open(d, ec):
    lat = build_lattice(d, ec, "open")
    ct = lat.complexes[ec]
    assert ct.S == d**2 * (d - 1)
    assert ct.Q == d**3 + 2 * d * (d - 1) ** 2
    assert len(ct.low_points) == d**2 and len(ct.high_points) == d**2
    # one stabilizer-graph edge per syndrome qubit
    n_faces = int(((ct.ninfo[:, None] >> (2 * np.arange(6))) & 3 != FACE_NONE).sum())
    n_bnd = len(ct.low_points) + len(ct.high_points)
    assert (n_faces + n_bnd) // 2 == ct.Q


def test_survey_table_sizes():
    # SURVEY.md section 8 size table
    lat = build_lattice(13, "primal", "open")
    assert (lat.N, len(lat.adj_indices) // 2) == (11725, 22512)
    lat = build_lattice(9, "primal", "open")
    assert (lat.N, len(lat.adj_indices) // 2) == (3689, 6944)
    ct = lat.complexes["primal"]
    assert (ct.S, ct.Q) == (648, 1881)


@pytest.mark.reference
@pytest.mark.parametrize("dims", [2, 3, (2, 3, 4)])
def test_rhg_graph_all_boundary_combinations(reference, dims):
    from flamingpy.codes.surface_code import RHG_graph, alternating_polarity

    for b in it.product(["primal", "dual", "periodic"], repeat=3):
        for pol in (None, alternating_polarity):
            G = RHG_graph(dims, np.array(b), polarity=pol)
            G.index_generator()
            lat = build_lattice(dims, "primal", list(b), polarity=pol is not None)
            pts = np.array([G.to_points[i] for i in range(len(G))])
            assert np.array_equal(pts, lat.coords)
            types = np.array([0 if G.nodes[tuple(p)]["type"] == "primal" else 1 for p in pts])
            assert np.array_equal(types, lat.node_type)
            A = G.adj_generator(sparse=True).tocsr()
            A.sort_indices()
            assert np.array_equal(A.indptr, lat.adj_indptr)
            assert np.array_equal(A.indices, lat.adj_indices)
            assert np.array_equal(A.data, lat.adj_vals)


@pytest.mark.reference
@pytest.mark.parametrize("d,ec,b", CODES)
def test_surface_code_tables(reference, d, ec, b):
    from oracle import ref_shim

    code = ref_shim.make_code(d, ec, b)
    lat = build_lattice(d, ec, b)
    ct = lat.complexes[ec]
    idx = code.graph.to_indices
    assert lat.N == len(code.graph)
    stabs = getattr(code, ec + "_stabilizers")
    assert len(stabs) == ct.S
    # stabilizers in order, same member sets
    for s, stab in enumerate(stabs):
        ref_members = sorted(idx[p] for p in stab.coords())
        mine = sorted(int(ct.qubits[q]) for q in ct.cube_qubit[s] if q >= 0)
        assert ref_members == mine, (s, ref_members, mine)
    # syndrome qubits
    assert sorted(set(getattr(code, ec + "_syndrome_inds"))) == ct.qubits.tolist()
    # boundary points: first half low, second half high
    bp = getattr(code, ec + "_bound_points")
    mid = len(bp) // 2
    assert sorted(idx[p] for p in bp[:mid]) == ct.low_points.tolist()
    assert sorted(idx[p] for p in bp[mid:]) == ct.high_points.tolist()
    # stabilizer graph: cube-cube edges with the same common vertex
    sg = getattr(code, ec + "_stab_graph")
    pos = {id(stab): s for s, stab in enumerate(stabs)}
    ref_edges = {}
    for u, v in sg.graph.edges():
        cv = sg.graph.edges[u, v]["common_vertex"]
        if id(u) in pos and id(v) in pos:
            ref_edges[frozenset((pos[id(u)], pos[id(v)]))] = idx[cv]
    mine = {}
    for c in range(ct.S):
        for dd in range(6):
            n = ct.nbr[c, dd]
            if 0 <= n < ct.S:
                k = frozenset((c, int(n)))
                q = int(ct.qubits[ct.cube_qubit[c, dd]])
                # size-2 periodic axes: two faces join the same cube pair, and the tables keep the one
                # qubit the reference's set.pop() keeps (stabilizer_graph.py:295-298)
                assert mine.setdefault(k, q) == q
    assert ref_edges == mine
    # cube - boundary point edges
    for c in range(ct.S):
        for dd in range(6):
            n = ct.nbr[c, dd]
            if n >= ct.S:
                pt = lat.coords[ct.qubits[ct.cube_qubit[c, dd]]]
                assert sg.graph.has_edge(stabs[c], tuple(int(v) for v in pt))
                assert sg.graph.has_edge("low" if n == ct.S else "high", tuple(int(v) for v in pt))
    # edge count: one per cube pair, plus two (cube-point, point-connector) per boundary point
    assert sg.graph.number_of_edges() == len(mine) + 2 * (len(ct.low_points) + len(ct.high_points))
    # check planes (decoder.py:197-214)
    names = {"periodic": "xyz", "toric": "xy", "open": "x" if ec == "primal" else "y"}[b]
    assert "".join(ct.plane_names) == names
    sheet = 0 if ec == "primal" else 1
    synd = set(getattr(code, ec + "_syndrome_coords"))
    for name, plane in zip(ct.plane_names, ct.planes):
        ref_plane = set(code.graph.slice_coords(name, sheet)) & synd
        assert sorted(idx[p] for p in ref_plane) == sorted(int(ct.qubits[q]) for q in plane)
    # slot table: every existing face maps to its qubit
    for c in range(ct.S):
        for dd in range(6):
            kind = (int(ct.ninfo[c]) >> (2 * dd)) & 3
            if kind == FACE_NONE:
                assert ct.cube_qubit[c, dd] < 0 or ct.nbr[c, dd] < 0  # no qubit, or the dropped double edge
                continue
            typ = 1 if kind == 1 else 0
            slot = ct.sbase[c] + ct.s_off[dd, typ]
            assert ct.slot_qubit[slot] == ct.cube_qubit[c, dd]


@pytest.mark.reference
def test_both_complexes_periodic(reference):
    from oracle import ref_shim

    code = ref_shim.make_code(3, "both", "periodic")
    lat = build_lattice(3, "both", "periodic")
    idx = code.graph.to_indices
    for ec in ("primal", "dual"):
        ct = lat.complexes[ec]
        stabs = getattr(code, ec + "_stabilizers")
        for s, stab in enumerate(stabs):
            assert sorted(idx[p] for p in stab.coords()) == sorted(
                int(ct.qubits[q]) for q in ct.cube_qubit[s] if q >= 0
            )
    allq = np.concatenate([lat.complexes[e].qubits for e in lat.ec])
    assert sorted(allq.tolist()) == list(range(lat.N))
