"""TEST INFRASTRUCTURE ONLY - generate tests/golden/*.npz from the live reference.

Run in the build container (needs /root/reference):

    python -m oracle.gen_golden

Each file holds, for a chain of consecutive trials on one reused reference
``CVLayer`` (so the carry-over of SURVEY.md finding 0.3 is included), the raw
draws and labels to inject plus every output of the reference for them:
homodyne values, bits, weights, parities, odd cubes, matching-graph lengths and
the result of ``correct``.  The reference's own tests hold no vectors for this
path, so these files are the committed known answers.
"""

import os

import numpy as np

from flamingpy_b200.lattice import build_lattice
from oracle import capture, ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, d, ec, boundaries, delta, p_swap, weight opts, trials, seed
CASES = [
    ("d3_open_primal_ps025", 3, "primal", "open", 0.09, 0.25, {"method": "blueprint", "delta": 0.09}, 16, 101),
    ("d3_open_dual_ps025", 3, "dual", "open", 0.09, 0.25, {"method": "blueprint", "delta": 0.09}, 8, 102),
    ("d3_periodic_both_ps0", 3, "both", "periodic", 0.09, 0, {"method": "blueprint", "delta": 0.09}, 8, 103),
    ("d4_toric_primal_ps05", 4, "primal", "toric", 0.1, 0.5, {"method": "blueprint", "delta": 0.1}, 6, 104),
    ("d5_open_primal_ps05", 5, "primal", "open", 0.08, 0.5, {"method": "blueprint", "delta": 0.08}, 6, 105),
    ("d5_periodic_dual_ps0", 5, "dual", "periodic", 0.09, 0, {"method": "blueprint", "delta": 0.09}, 4, 106),
    ("d4_open_primal_ps1_uniform", 4, "primal", "open", 0.09, 1, {"method": "uniform"}, 4, 107),
    ("d4_open_primal_ps1_blueprint", 4, "primal", "open", 0.09, 1, {"method": "blueprint", "delta": 0.09}, 4, 108),
    ("d3_open_primal_int100", 3, "primal", "open", 0.09, 0.25,
     {"method": "blueprint", "delta": 0.09, "integer": True, "multiplier": 100}, 6, 109),
    ("d234_open_primal_ps01", (2, 3, 4), "primal", "open", 0.12, 0.1, {"method": "blueprint", "delta": 0.12}, 6, 110),
    ("d7_open_primal_ps0", 7, "primal", "open", 0.105, 0, {"method": "blueprint", "delta": 0.105}, 3, 111),
    # size-2 periodic axes (two cubes share two qubits, the reference keeps one edge per pair)
    ("d2_periodic_both_ps0", 2, "both", "periodic", 0.12, 0, {"method": "blueprint", "delta": 0.12}, 8, 112),
    ("d223_toric_dual_ps025", (2, 2, 3), "dual", "toric", 0.12, 0.25, {"method": "blueprint", "delta": 0.12}, 6, 113),
]


def generate(case):
    name, d, ec, b, delta, p_swap, wopts, trials, seed = case
    code = ref_shim.make_code(d, ec, b)
    layer = capture.make_layer(code, delta, p_swap)
    lat = build_lattice(d, ec, b)
    rng = np.random.default_rng(seed)
    data = {
        "dims": np.array(lat.dims), "ec": np.array(ec), "boundaries": np.array(b),
        "delta": np.array(delta), "p_swap": np.array(float(p_swap)), "trials": np.array(trials),
        "seed": np.array(seed), "method": np.array(wopts["method"]),
        "integer": np.array(bool(wopts.get("integer", False))),
        "multiplier": np.array(wopts.get("multiplier", 1)),
    }
    xs, states, succ = [], [], []
    per = {e: {k: [] for k in ("hom", "bits", "weights", "parity", "odd", "pair_len", "bnd_len", "bnd_side")}
           for e in lat.ec}
    for _ in range(trials):
        r = capture.capture_trial(code, layer, lat, rng, wopts, run_correct=True)
        xs.append(r["x"])
        states.append(r["state"])
        succ.append(r["success"])
        for e in lat.ec:
            for k in per[e]:
                per[e][k].append(r["complex"][e][k])
    data["x"] = np.stack(xs)
    data["state"] = np.stack(states)
    data["success"] = np.array(succ)
    for e in lat.ec:
        for k in ("hom", "bits", "weights", "parity"):
            data[f"{e}_{k}"] = np.stack(per[e][k])
        for k in ("odd", "pair_len", "bnd_len", "bnd_side"):
            arrs = per[e][k]
            data[f"{e}_{k}_off"] = np.cumsum([0] + [len(a) for a in arrs]).astype(np.int64)
            data[f"{e}_{k}"] = np.concatenate(arrs) if arrs else np.empty(0)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **data)
    return name


if __name__ == "__main__":
    import sys

    for c in CASES:
        if len(sys.argv) > 1 and c[0] not in sys.argv[1:]:
            continue  # regenerate only the named sets
        print("wrote", generate(c))
