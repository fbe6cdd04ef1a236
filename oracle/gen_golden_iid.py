"""TEST INFRASTRUCTURE ONLY - golden vectors of the i.i.d. Pauli-noise arm from the live reference.

    python -m oracle.gen_golden_iid        # build container only (needs /root/reference)

Each tests/golden/iid_*.npz holds consecutive trials of the reference's ``IidNoise.apply_noise``
(``flamingpy/noise/iid.py``) decoded by ``correct(..., weight_options={"method": "uniform"},
decoder_opts={"backend": "networkx"})``: the bit of every lattice node (``flamingpy_b200.lattice``
index order), and per complex the stabilizer parities, odd cubes, matching-graph lengths
(``NxMatchingGraph``) and the verdict.  The bits are what gets injected; the reference's mapping of
its RNG stream to nodes is a set-iteration artefact and is not restated (see oracle/restate.py).
"""

import os

import numpy as np

from flamingpy_b200.lattice import build_lattice
from oracle import capture, ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, d, ec, boundaries, error probability, trials, seed
CASES = [
    ("iid_d3_open_primal_p05", 3, "primal", "open", 0.05, 24, 201),
    ("iid_d4_open_dual_p08", 4, "dual", "open", 0.08, 12, 202),
    ("iid_d3_periodic_both_p04", 3, "both", "periodic", 0.04, 12, 203),
    ("iid_d5_open_primal_p03", 5, "primal", "open", 0.03, 8, 204),
    ("iid_d4_toric_primal_p06", 4, "primal", "toric", 0.06, 8, 205),
]


def generate(case):
    name, d, ec, b, p, trials, seed = case
    code = ref_shim.make_code(d, ec, b)
    lat = build_lattice(d, ec, b)
    rng = np.random.default_rng(seed)
    data = {"dims": np.array(lat.dims), "ec": np.array(ec), "boundaries": np.array(b), "p": np.array(p),
            "trials": np.array(trials), "seed": np.array(seed)}
    bits, succ = [], []
    per = {e: {k: [] for k in ("parity", "odd", "pair_len", "bnd_len", "bnd_side")} for e in lat.ec}
    for _ in range(trials):
        r = capture.capture_iid_trial(code, lat, rng, p)
        bits.append(r["bit_val"])
        succ.append(r["success"])
        for e in lat.ec:
            for k in per[e]:
                per[e][k].append(r["complex"][e][k])
    data["bit_val"] = np.stack(bits)
    data["success"] = np.array(succ)
    for e in lat.ec:
        data[f"{e}_parity"] = np.stack(per[e]["parity"])
        for k in ("odd", "pair_len", "bnd_len", "bnd_side"):
            arrs = per[e][k]
            data[f"{e}_{k}_off"] = np.cumsum([0] + [len(a) for a in arrs]).astype(np.int64)
            data[f"{e}_{k}"] = np.concatenate(arrs) if arrs else np.empty(0)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **data)
    return name, int(np.sum(succ)), trials


if __name__ == "__main__":
    for c in CASES:
        print(generate(c))
