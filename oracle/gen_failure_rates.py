"""TEST INFRASTRUCTURE ONLY - reference logical-failure rates for the statistical parity test.

    python -m oracle.gen_failure_rates        # writes tests/golden/failure_rates.json

Runs the CPU restatement (oracle/restate.py, pinned trial-for-trial against the live
reference including its networkx matching and the reused-CVLayer carry-over) with NumPy RNG
on 8 processes.  SURVEY.md section 6 lists the same quantities measured on the reference
itself with 8 x 1000 trials (seeds 200..207); the restatement reproduces those counts exactly
(86/8000 and 181/8000 at d=3, p_swap=0) - this script just buys smaller error bars.
"""

import json
import os
import sys
from multiprocessing import Pool

import numpy as np

from flamingpy_b200.lattice import build_lattice
from oracle import restate

CASES = [
    # d, boundaries, delta, p_swap, trials per process
    (3, "open", 0.09, 0.25, 5000),
    (3, "open", 0.09, 0.0, 5000),
    (5, "open", 0.09, 0.0, 1500),
]


def work(args):
    d, b, delta, ps, n, fresh, seed = args
    lat = build_lattice(d, "primal", b)
    chain = restate.TrialChain(lat, delta, ps, {"method": "blueprint", "delta": delta}, fresh=fresh)
    rng = np.random.default_rng(seed)
    fails = 0
    for _ in range(n):
        res = chain.next(rng, want_graph=False)
        fails += not chain.correct(res)
    return fails


if __name__ == "__main__":
    out = []
    with Pool(8) as pool:
        for d, b, delta, ps, n in CASES:
            for fresh in (False, True):
                fs = pool.map(work, [(d, b, delta, ps, n, fresh, 200 + s) for s in range(8)])
                rec = {"distance": d, "boundaries": b, "delta": delta, "p_swap": ps,
                       "mode": "fresh" if fresh else "compat", "failures": int(sum(fs)), "trials": 8 * n,
                       "seeds": "200..207", "trials_per_chain": n}
                print(rec, file=sys.stderr)
                out.append(rec)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "failure_rates.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
