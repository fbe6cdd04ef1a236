"""Reference-compatible ``SurfaceCode`` backed by flat lattice tables.

Mirrors the public surface of ``flamingpy.codes.SurfaceCode``
(``flamingpy/codes/surface_code.py:248-371``): same constructor, ``ec`` list,
``bound_str``, ``graph`` with ``to_indices`` / ``to_points`` / ``nodes[...]`` /
``slice_coords``, ``{ec}_stabilizers``, ``{ec}_syndrome_coords`` / ``_inds``,
``all_syndrome_*``, ``{ec}_bound_points`` and ``{ec}_stab_graph``.  Node
attributes (``state``, ``hom_val_p``, ``bit_val``, ``weight``) - the
reference's hand-off between noise layer and decoder - live in NumPy arrays
and are exposed through dict-like views, so reference-style code keeps
working while the CUDA pipeline reads and writes the arrays in bulk.

Differences kept on purpose: construction is vectorised (no O(S^2) pairwise
intersection), ``ec="both"`` is accepted for periodic boundaries (the
reference disables it at ``surface_code.py:323-331``), and the only
``backend`` is the B200 one.
"""

from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .lattice import LatticeTables, build_lattice

_STATE_NAMES = {0: "GKP", 1: "GKP", 2: "p"}  # silent nodes keep their stale "GKP" label


class _NodeData:
    """Dict-like view of one node's attributes."""

    __slots__ = ("_g", "_i")

    def __init__(self, graph, index):
        self._g, self._i = graph, index

    def __getitem__(self, key):
        g, i = self._g, self._i
        if key == "type":
            return "primal" if g.lattice.node_type[i] == 0 else "dual"
        if key == "state":
            if g.state is None:
                raise KeyError(key)
            return _STATE_NAMES[int(g.state[i])]
        arr = {"hom_val_p": g.hom_val_p, "bit_val": g.bit_val, "weight": g.weight,
               "p_phase_cond": g.p_phase_cond}.get(key)
        if arr is None or (key != "bit_val" and np.isnan(arr[i])) or (key == "bit_val" and arr[i] < 0):
            raise KeyError(key)
        return int(arr[i]) if key == "bit_val" else arr[i]

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __setitem__(self, key, value):
        g, i = self._g, self._i
        if key == "state":
            g._ensure_arrays()
            g.state[i] = 2 if value == "p" else 1
        elif key == "bit_val":
            g._ensure_arrays()
            g.bit_val[i] = int(value)
        elif key in ("hom_val_p", "weight", "p_phase_cond"):
            g._ensure_arrays()
            if key == "p_phase_cond" and g.p_phase_cond is None:
                g.p_phase_cond = np.full(g.lattice.N, np.nan)
            getattr(g, key)[i] = value
        else:
            raise KeyError(key)

    def __contains__(self, key):
        return self.get(key) is not None


class _NodesView:
    def __init__(self, graph):
        self._g = graph

    def __getitem__(self, point):
        return _NodeData(self._g, self._g.to_indices[tuple(point)])

    def __iter__(self):
        return iter(self._g.to_points[i] for i in range(len(self._g)))

    def __len__(self):
        return len(self._g)

    def __contains__(self, point):
        return tuple(point) in self._g.to_indices


class LatticeGraph:
    """The slice of ``EGraph`` the Monte-Carlo path touches
    (``codes/graphs/egraph.py:133-194``)."""

    def __init__(self, lattice: LatticeTables):
        self.lattice = lattice
        pts = [tuple(int(v) for v in c) for c in lattice.coords]
        self.to_points = dict(enumerate(pts))
        self.to_indices = {p: i for i, p in enumerate(pts)}
        self.graph: Dict = {}
        self.adj_mat = None
        self.state = None
        self.hom_val_p = None
        self.bit_val = None
        self.weight = None
        self.p_phase_cond = None  # CVMacroLayer: conditional phase error probability of the reduced nodes

    def _ensure_arrays(self):
        N = self.lattice.N
        if self.state is None:
            self.state = np.ones(N, np.uint8)
        if self.hom_val_p is None:
            self.hom_val_p = np.full(N, np.nan)
        if self.bit_val is None:
            self.bit_val = np.full(N, -1, np.int8)
        if self.weight is None:
            self.weight = np.full(N, np.nan)

    def __len__(self):
        return self.lattice.N

    def order(self):
        return self.lattice.N

    @property
    def nodes(self):
        return _NodesView(self)

    def __iter__(self):
        return iter(self.nodes)

    def __contains__(self, point):
        return tuple(point) in self.to_indices

    def __getitem__(self, point):
        """Neighbours of a node (``G[node]`` in networkx)."""
        lat = self.lattice
        i = self.to_indices[tuple(point)]
        return [self.to_points[int(j)] for j in lat.adj_indices[lat.adj_indptr[i] : lat.adj_indptr[i + 1]]]

    def index_generator(self):
        return self.to_indices

    def adj_generator(self, sparse=True):
        import scipy.sparse as sp

        lat = self.lattice
        adj = sp.csr_array((lat.adj_vals, lat.adj_indices, lat.adj_indptr), shape=(lat.N, lat.N))
        self.adj_mat = adj if sparse else adj.toarray()
        return self.adj_mat

    def slice_coords(self, plane, number):
        a = {"x": 0, "y": 1, "z": 2}[plane]
        return [self.to_points[int(i)] for i in np.nonzero(self.lattice.coords[:, a] == number)[0]]

    def number_of_edges(self):
        return len(self.lattice.adj_indices) // 2


class Stabilizer:
    """One cube (``flamingpy/codes/stabilizer.py``): its member qubits and parity."""

    def __init__(self, code, ec, index):
        self._code, self.ec, self.index = code, ec, index
        ct = code.lattice.complexes[ec]
        self._nodes = [int(ct.qubits[q]) for q in ct.cube_qubit[index] if q >= 0]

    def coords(self):
        g = self._code.graph
        return [g.to_points[i] for i in self._nodes]

    @property
    def parity(self):
        """Sum of ``bit_val`` over the cube mod 2 (``stabilizer.py:43-55``)."""
        bits = self._code.graph.bit_val
        return int(np.sum(bits[self._nodes]) % 2)

    def __repr__(self):
        return f"Stabilizer at nodes {self.coords()}"

    def __hash__(self):
        return hash((self.ec, self.index))

    def __eq__(self, other):
        return isinstance(other, Stabilizer) and (self.ec, self.index) == (other.ec, other.index)


class StabilizerGraphView:
    """What decoders read from ``{ec}_stab_graph``
    (``codes/graphs/stabilizer_graph.py:27-329``): the stabilizer list, boundary
    points and odd-parity filter.  Shortest paths are not computed here - the CUDA
    pipeline produces the matching-graph lengths directly."""

    def __init__(self, code, ec):
        self._code, self.ec = code, ec
        ct = code.lattice.complexes[ec]
        self.stabilizers = getattr(code, ec + "_stabilizers")
        g = code.graph
        self.low_bound_points = [g.to_points[int(i)] for i in ct.low_points]
        self.high_bound_points = [g.to_points[int(i)] for i in ct.high_points]

    def bound_points(self):
        return self.low_bound_points + self.high_bound_points

    def has_bound_points(self):
        return bool(self.low_bound_points or self.high_bound_points)

    def odd_parity_stabilizers(self):
        return filter(lambda s: s.parity % 2 == 1, self.stabilizers)

    def real_nodes(self):
        return list(self.stabilizers) + self.bound_points()

    def assign_weights(self, code):
        """Edge weights are read straight from the qubit weights by the kernels."""

    def edge_data(self, node1, node2):
        """``{"common_vertex": coords, "weight": w}`` of a stabilizer-graph edge
        (``stabilizer_graph.py:72-88,289-302,321-329``): between two adjacent cubes the shared
        syndrome qubit, between a cube and a boundary point the point itself."""
        code, ct = self._code, self._code.lattice.complexes[self.ec]
        g = code.graph
        if isinstance(node1, Stabilizer) and isinstance(node2, Stabilizer):
            faces = np.nonzero(ct.nbr[node1.index] == node2.index)[0]
            if node1.ec != self.ec or node2.ec != self.ec or len(faces) == 0:
                raise KeyError((node1, node2))
            node = int(ct.qubits[ct.cube_qubit[node1.index, faces[0]]])
        else:
            stab, point = (node1, node2) if isinstance(node1, Stabilizer) else (node2, node1)
            node = g.to_indices.get(tuple(point)) if not isinstance(point, str) else None
            if not isinstance(stab, Stabilizer) or node is None or node not in stab._nodes or (
                    node not in ct.low_points and node not in ct.high_points):
                raise KeyError((node1, node2))
        weight = None if g.weight is None or np.isnan(g.weight[node]) else g.weight[node]
        return {"common_vertex": g.to_points[node], "weight": weight}


class SurfaceCode:
    """``SurfaceCode(distance, ec="primal", boundaries="open", polarity=None, backend="b200")``."""

    def __init__(self, distance, ec="primal", boundaries="open", polarity=None, backend="b200"):
        self.distance = distance
        if np.ndim(distance) == 0:
            self.dims = (int(distance),) * 3
        elif np.size(distance) == 3:
            self.dims = tuple(int(v) for v in distance)
        else:
            raise ValueError("distance must be an integer or a sequence of three integers.")
        allowed_ec = ("primal", "dual", "both")
        if not (boundaries in ("open", "toric", "periodic") and ec in allowed_ec):
            raise ValueError(
                f"The combination `ec={ec}` and `boundaries={boundaries} is not valid."
                "Allowed choices for boundaries are 'open', 'toric' or 'periodic'"
                "and `ec` should be either 'primal' or 'dual'."
            )
        if ec == "both" and boundaries != "periodic":
            raise ValueError("ec='both' is only supported with periodic boundaries.")
        if backend not in ("b200", "rustworkx", "networkx"):
            raise ValueError("Invalid backend; options are 'networkx' and 'rustworkx'.")
        self.backend = "b200"
        self.ec = ["primal", "dual"] if ec == "both" else [ec]
        use_polarity = False
        if polarity is not None:
            if getattr(polarity, "__name__", "") != "alternating_polarity" and polarity is not True:
                raise ValueError("only the reference's alternating_polarity edge signs are supported")
            use_polarity = True
        self.polarity = polarity
        self.lattice = build_lattice(self.dims, ec, boundaries, polarity=use_polarity)
        self.bound_str = self.lattice.bound_str
        self.boundaries = np.array(self.lattice.boundaries)
        self.graph = LatticeGraph(self.lattice)
        g = self.graph
        for e in self.ec:
            ct = self.lattice.complexes[e]
            setattr(self, e + "_stabilizers", [Stabilizer(self, e, s) for s in range(ct.S)])
            setattr(self, e + "_syndrome_inds", [int(i) for i in ct.qubits])
            setattr(self, e + "_syndrome_coords", [g.to_points[int(i)] for i in ct.qubits])
            setattr(self, e + "_bound_points",
                    [g.to_points[int(i)] for i in ct.low_points] + [g.to_points[int(i)] for i in ct.high_points])
        self.all_syndrome_inds = sum((getattr(self, e + "_syndrome_inds") for e in self.ec), start=[])
        self.all_syndrome_coords = sum((getattr(self, e + "_syndrome_coords") for e in self.ec), start=[])
        for e in self.ec:
            setattr(self, e + "_stab_graph", StabilizerGraphView(self, e))
        self._engines: Dict[Tuple, object] = {}
        self._trial = None  # (x, state) of the last apply_noise, injected into correct()

    def alternating_polarity(self, edge):  # pragma: no cover - convenience
        raise NotImplementedError

    def q_offset(self, ec):
        """Column of the first syndrome qubit of complex ``ec`` in the engine's qubit order."""
        off = 0
        for e in self.ec:
            if e == ec:
                return off
            off += self.lattice.complexes[e].Q
        raise KeyError(ec)

    def qubit_cubes(self, ec):
        """int32[Q, 2]: the one or two cubes of complex ``ec`` that contain each of its syndrome
        qubits (a boundary point lies in one cube: both entries equal)."""
        cache = self.__dict__.setdefault("_qubit_cubes", {})
        if ec not in cache:
            ct = self.lattice.complexes[ec]
            qc = np.full((ct.Q, 2), -1, np.int32)
            for f in range(6):
                q = ct.cube_qubit[:, f]
                for c in np.nonzero(q >= 0)[0]:
                    col = 0 if qc[q[c], 0] in (-1, c) else 1
                    qc[q[c], col] = c
            qc[qc[:, 1] < 0, 1] = qc[qc[:, 1] < 0, 0]
            cache[ec] = qc
        return cache[ec]

    def release_engines(self):
        """Close every cached engine (device buffers, streams).  Sweeps over noise / weight
        parameters call this after each point: the cache is keyed by those parameters and would
        otherwise hold every point's batch buffers until the code object dies."""
        for eng in self._engines.values():
            eng.close()
        self._engines.clear()

    def graph_engine(self, device=0):
        """The engine that builds matching graphs from the ``weight`` / ``bit_val`` node attributes
        (``fpb_run_weighted``); its own handle, never switched to the static-weight table."""
        return self.engine(1.0, 0.0, {"method": "uniform"}, device=device, lane="weighted")

    # ------------------------------------------------------------------ engine cache
    def engine(self, delta, p_swap, weight_opts=None, device=0, mode="compat", lane=0, reserve_sms=0,
               lengths="float"):
        """The ``TrialEngine`` for these noise / weight options (created once).  ``lane`` > 0
        gives further engines with the same options (own stream and buffers) for pipelining."""
        from .engine import TrialEngine

        wo = dict(weight_opts or {})
        key = (float(delta), float(p_swap or 0.0), tuple(sorted(wo.items())), device, mode, lane, reserve_sms, lengths)
        eng = self._engines.get(key)
        if eng is None:
            eng = TrialEngine(self.lattice, delta, p_swap, wo, mode=mode, device=device, reserve_sms=reserve_sms,
                              lengths=lengths)
            if lane and lane != "weighted" and self.engine(delta, p_swap, weight_opts, device, mode,
                                                           reserve_sms=reserve_sms).uses_table:
                eng.build_table()
            self._engines[key] = eng
        return eng


def alternating_polarity(edge):
    """Same signs as the reference's ``alternating_polarity``
    (``surface_code.py:26-53``); pass it as ``polarity=`` to ``SurfaceCode``."""
    p1, p2 = np.array(edge[0]), np.array(edge[1])
    direction = np.where(p2 - p1)[0][0]
    src = {0: p1[1], 1: p1[2], 2: p1[0]}[int(direction)]
    return int((-1) ** int(src))
