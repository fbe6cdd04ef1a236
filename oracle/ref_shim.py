"""TEST INFRASTRUCTURE ONLY - loader for the live reference (XanaduAI/flamingpy).

Only ``tests/``, ``oracle/gen_golden.py`` and ad-hoc validation scripts may use
this.  It imports the unmodified reference from ``/root/reference`` (present in
the build container only, never on the GPU box) after registering empty stub
modules for the optional dependencies that are not installed here
(``rustworkx``, ``matplotlib``, ``thewalrus``); the networkx code path of the
reference needs none of them.  Nothing in the product
package imports this file.
"""

import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("FLAMINGPY_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "flamingpy"))


def load():
    """Import and return the reference ``flamingpy`` package (networkx path)."""
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("rustworkx", "matplotlib", "thewalrus", "thewalrus.symplectic"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__version__ = "0.0-stub"
            sys.modules[name] = mod
    rx = sys.modules["rustworkx"]
    if getattr(rx, "PyGraph", None) is None:
        # rustworkx is not installed here.  The reference's Union-Find decoder (decoders/unionfind/)
        # uses a small part of its public PyGraph API for the erasure graph and the spanning forest;
        # a plain-Python stand-in of those calls lets the unmodified decoder run.
        rx.PyGraph = _PyGraph
        rx.connected_components = _connected_components
        rx.minimum_spanning_tree = _minimum_spanning_tree
    sym = sys.modules["thewalrus.symplectic"]
    if getattr(sym, "expand", None) is None:
        # thewalrus is not installed here.  The reference uses two of its functions, for the fixed
        # four-splitter of CVMacroLayer only (cv/ops.py:104-127); both have short closed forms
        # (thewalrus.symplectic, "xxpp" ordering), restated here so that the unmodified reference
        # class runs: a lossless two-mode beamsplitter and the embedding of a symplectic into N modes.
        sym.beam_splitter = _beam_splitter
        sym.expand = _expand
    sys.modules["thewalrus"].symplectic = sym
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import flamingpy  # noqa: F401
        import flamingpy.codes  # noqa: F401
        import flamingpy.noise  # noqa: F401
        import flamingpy.decoders.decoder  # noqa: F401
    return sys.modules["flamingpy"]


class _PyGraph:
    """The calls of rustworkx.PyGraph the reference's Union-Find decoder makes (undirected multigraph,
    integer node indices in insertion order, arbitrary payloads)."""

    def __init__(self):
        self._nodes = []
        self._adj = []  # per node: {neighbour: [payload, ...]}

    def add_node(self, obj):
        self._nodes.append(obj)
        self._adj.append({})
        return len(self._nodes) - 1

    def add_edge(self, a, b, data):
        self._adj[a].setdefault(b, []).append(data)
        if a != b:
            self._adj[b].setdefault(a, []).append(data)

    def nodes(self):
        return list(self._nodes)

    def out_edges(self, node):
        return [(node, nb, d) for nb, ds in self._adj[node].items() for d in ds]

    def get_edge_data(self, a, b):
        return self._adj[a][b][0]

    def remove_edge(self, a, b):
        self._adj[a][b].pop()
        if not self._adj[a][b]:
            del self._adj[a][b]
        if a != b:
            self._adj[b][a].pop()
            if not self._adj[b][a]:
                del self._adj[b][a]

    def degree(self, node):
        return sum(len(ds) for ds in self._adj[node].values())


def _connected_components(graph):
    seen, out = set(), []
    for start in range(len(graph._nodes)):
        if start in seen:
            continue
        comp, stack = {start}, [start]
        seen.add(start)
        while stack:
            u = stack.pop()
            for v in graph._adj[u]:
                if v not in seen:
                    seen.add(v)
                    comp.add(v)
                    stack.append(v)
        out.append(comp)
    return out


def _minimum_spanning_tree(graph):
    """All weights equal: any spanning forest (same node indices, payloads kept)."""
    forest = _PyGraph()
    for obj in graph._nodes:
        forest.add_node(obj)
    parent = list(range(len(graph._nodes)))

    def find(u):
        while parent[u] != u:
            parent[u] = parent[parent[u]]
            u = parent[u]
        return u

    for a in range(len(graph._nodes)):
        for b, ds in graph._adj[a].items():
            if a < b and find(a) != find(b):
                parent[find(a)] = find(b)
                forest.add_edge(a, b, ds[0])
    return forest


def _beam_splitter(theta, phi):
    """Symplectic of the beamsplitter U = [[cos t, -e^{-i phi} sin t], [e^{i phi} sin t, cos t]] in
    (q1, q2, p1, p2) ordering: [[Re U, -Im U], [Im U, Re U]]."""
    import numpy as np

    ct, st = np.cos(theta), np.sin(theta)
    eip = np.cos(phi) + 1j * np.sin(phi)
    U = np.array([[ct, -np.conj(eip) * st], [eip * st, ct]])
    X, Y = U.real, U.imag
    return np.block([[X, -Y], [Y, X]])


def _expand(S, modes, N):
    """Embed the symplectic S of len(modes) modes into N modes (identity elsewhere), xxpp ordering."""
    import numpy as np

    M = len(S) // 2
    S2 = np.identity(2 * N, dtype=S.dtype)
    w = np.array([modes]) if isinstance(modes, int) else np.array(modes)
    S2[w.reshape(-1, 1), w.reshape(1, -1)] = S[:M, :M].copy()
    S2[(w + N).reshape(-1, 1), (w + N).reshape(1, -1)] = S[M:, M:].copy()
    S2[w.reshape(-1, 1), (w + N).reshape(1, -1)] = S[:M, M:].copy()
    S2[(w + N).reshape(-1, 1), w.reshape(1, -1)] = S[M:, :M].copy()
    return S2


def make_code(distance, ec="primal", boundaries="open", polarity=None):
    """``SurfaceCode(..., backend="networkx")``; ``ec="both"`` bypasses the
    constructor guard at ``surface_code.py:323-331`` the way SURVEY.md finding
    0.2 describes (periodic boundaries only)."""
    load()
    from flamingpy.codes import SurfaceCode
    from flamingpy.codes.graphs import NxStabilizerGraph

    if ec != "both":
        return SurfaceCode(distance, ec=ec, boundaries=boundaries, polarity=polarity, backend="networkx")
    if boundaries != "periodic":
        raise ValueError("ec='both' oracle is only defined for periodic boundaries")
    code = SurfaceCode(distance, ec="primal", boundaries=boundaries, polarity=polarity, backend="networkx")
    code.ec = ["primal", "dual"]
    code.identify_stabilizers()
    code.identify_boundary()
    for error_type in code.ec:
        setattr(code, error_type + "_stab_graph", NxStabilizerGraph(error_type, code))
    return code
