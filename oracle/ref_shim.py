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
