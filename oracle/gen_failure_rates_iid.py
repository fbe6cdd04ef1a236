"""TEST INFRASTRUCTURE ONLY - reference logical-failure rates of the i.i.d. Pauli-noise arm.

    python -m oracle.gen_failure_rates_iid     # writes tests/golden/failure_rates_iid.json

CPU restatement (oracle/restate.py: iid_bits -> trial_from_bits -> networkx matching -> logical
check; pinned trial-for-trial against the live reference on injected bits, tests/golden/iid_*.npz)
with NumPy RNG on 8 processes.
"""

import json
import os
import sys
from multiprocessing import Pool

import numpy as np

from flamingpy_b200.lattice import build_lattice
from oracle import restate

CASES = [  # d, ec, boundaries, error probability, trials per process
    (3, "primal", "open", 0.05, 5000),
    (5, "primal", "open", 0.03, 2000),
    (3, "primal", "periodic", 0.04, 3000),
]


def work(args):
    d, ec, b, p, n, seed = args
    lat = build_lattice(d, ec, b)
    rng = np.random.default_rng(seed)
    fails = 0
    for _ in range(n):
        res = restate.trial_from_bits(lat, restate.iid_bits(rng, lat.N, p), want_graph=False)
        fails += not restate.correct_trial(lat, res)
    return fails


if __name__ == "__main__":
    out = []
    with Pool(8) as pool:
        for d, ec, b, p, n in CASES:
            fs = pool.map(work, [(d, ec, b, p, n, 300 + s) for s in range(8)])
            rec = {"distance": d, "ec": ec, "boundaries": b, "p": p, "failures": int(sum(fs)), "trials": 8 * n,
                   "seeds": "300..307"}
            print(rec, file=sys.stderr)
            out.append(rec)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "failure_rates_iid.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
