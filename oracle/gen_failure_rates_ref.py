"""TEST INFRASTRUCTURE ONLY - logical-failure rates measured on the LIVE reference (build container only).

    python -m oracle.gen_failure_rates_ref     # writes tests/golden/failure_rates_reference.json

Eight processes stand in for the reference's mpi4py ranks; each runs the reference's own loop body
(flamingpy/simulations.py:46-59: `noise.apply_noise(rng)`, `correct(code, "MWPM", weight_options)` with the
networkx backend) on one reused noise layer, exactly as `ec_monte_carlo` does per rank - so these are "compat"
rates (SURVEY.md finding 0.3).  CVMacroLayer runs with thewalrus' two functions restated (oracle/ref_shim.py).
`oracle/gen_failure_rates.py` makes the larger samples from the pinned restatement; this file pins the
device-RNG path to the reference itself.
"""

import json
import os
import sys
import time
from multiprocessing import Pool

# noise, d, boundaries, delta, p_swap, weight opts, trials per process
CASES = [
    ("CVLayer", 3, "open", 0.09, 0.25, {"method": "blueprint", "delta": 0.09}, 2500),
    ("CVLayer", 3, "open", 0.09, 0.0, {"method": "blueprint", "delta": 0.09}, 2500),
    ("CVLayer", 5, "open", 0.09, 0.0, {"method": "blueprint", "delta": 0.09}, 300),
    ("CVMacroLayer", 3, "open", 0.1, 0.25, {"method": "blueprint", "prob_precomputed": True}, 1500),
    ("CVMacroLayer", 3, "open", 0.12, 0.0, {"method": "blueprint", "prob_precomputed": True}, 1500),
]


def work(args):
    noise, d, b, delta, ps, wo, n, seed = args
    import numpy as np

    from oracle import ref_shim

    ref_shim.load()
    import flamingpy.noise as fn
    from flamingpy.decoders.decoder import correct

    code = ref_shim.make_code(d, "primal", b)
    layer = getattr(fn, noise)(code, delta=delta, p_swap=ps)
    rng = np.random.default_rng(seed)
    fails = 0
    for _ in range(n):
        layer.apply_noise(rng)
        fails += not correct(code, "MWPM", weight_options=dict(wo), decoder_opts={"backend": "networkx"})
    return fails


if __name__ == "__main__":
    out = []
    with Pool(8) as pool:
        for noise, d, b, delta, ps, wo, n in CASES:
            t0 = time.time()
            fs = pool.map(work, [(noise, d, b, delta, ps, wo, n, 300 + s) for s in range(8)])
            rec = {"noise": noise, "distance": d, "boundaries": b, "delta": delta, "p_swap": ps, "weight_opts": wo,
                   "mode": "compat", "failures": int(sum(fs)), "trials": 8 * n, "seeds": "300..307",
                   "trials_per_chain": n, "source": "live reference (networkx backend)", "seconds": round(time.time() - t0, 1)}
            print(rec, file=sys.stderr, flush=True)
            out.append(rec)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "failure_rates_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
