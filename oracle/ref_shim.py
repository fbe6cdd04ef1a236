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
    if not hasattr(sym, "expand"):
        sym.expand = None
        sym.beam_splitter = None
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
