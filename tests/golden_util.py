"""Helpers to read tests/golden/*.npz (see oracle/gen_golden.py)."""

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    """CVLayer chains (oracle/gen_golden.py); the iid_*.npz files of the i.i.d. arm have their own
    layout and tests (tests/test_iid.py)."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("iid_", "macro_", "uf_")))


def macro_files():
    """CVMacroLayer chains (oracle/gen_golden_macro.py)."""
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "macro_*.npz")))


def uf_files():
    """Union-Find decoder chains (oracle/gen_golden_uf.py)."""
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "uf_*.npz")))


def load_uf(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    dims = tuple(int(v) for v in g["dims"])
    g["dims"] = dims
    g["distance"] = dims[0] if len(set(dims)) == 1 else dims
    for k in ("ec", "boundaries"):
        g[k] = str(g[k])
    g["delta"] = float(g["delta"])
    ps = float(g["p_swap"])
    g["p_swap"] = int(ps) if ps in (0.0, 1.0) else ps
    g["trials"] = int(g["trials"])
    return g


def load_macro(path):
    g = load(path)
    wo = {"method": g["method"]}
    if bool(g["prob_precomputed"]):
        wo["prob_precomputed"] = True
    if g["integer"]:
        wo.update(integer=True, multiplier=g["multiplier"])
    g["weight_opts"] = wo
    return g


def load(path):
    z = np.load(path)
    g = {k: z[k] for k in z.files}
    dims = tuple(int(v) for v in g["dims"])
    g["dims"] = dims
    g["distance"] = dims[0] if len(set(dims)) == 1 else dims
    for k in ("ec", "boundaries", "method"):
        g[k] = str(g[k])
    g["delta"] = float(g["delta"])
    ps = float(g["p_swap"])
    g["p_swap"] = int(ps) if ps in (0.0, 1.0) else ps
    g["trials"] = int(g["trials"])
    g["integer"] = bool(g["integer"])
    g["multiplier"] = int(g["multiplier"])
    g["weight_opts"] = {"method": g["method"], "delta": g["delta"], "integer": g["integer"],
                        "multiplier": g["multiplier"]}
    return g


def ragged(g, key, t):
    off = g[key + "_off"]
    return g[key][off[t]:off[t + 1]]
