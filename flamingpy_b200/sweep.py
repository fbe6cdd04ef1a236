"""Threshold-sweep driver (BASELINE.json configs[4]): logical failure counts over a grid of code
distances and noise parameters, one CSV row per point with the reference's columns
(``flamingpy/simulations.py:151-194``) plus throughput columns.

    python -m flamingpy_b200.sweep -distances 9,13 -deltas 0.06,0.09,0.12 -trials 1000 -fname sweep.csv

Defaults follow the reference's (`surface_code.py:312-313`): open boundaries, primal complex,
p_swap = 0, blueprint weights.  With ``-graph_only`` the host matcher is skipped and only the
device stage (noise -> weights -> syndrome -> matching graph) is timed.
"""

from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

from . import _lib
from .codes import SurfaceCode
from .noise import CVLayer
from .simulations import _dist_info, ec_monte_carlo, trial_range

COLUMNS = ["code", "distance", "ec", "boundaries", "noise", "delta", "p_swap", "decoder", "weight_opts", "errors",
           "trials", "current_time", "simulation_time", "mpi_size", "mode", "gpus", "device_trials_per_s"]


def sweep(distances, deltas, trials, p_swap=0.0, boundaries="open", ec="primal", mode="fresh", fname=None,
          graph_only=False, seed=20261019, batch=1024, backend="native"):
    _, rank, size = _dist_info()
    rows = []
    for d in distances:
        code = SurfaceCode(d, ec, boundaries)
        for delta in deltas:
            wo = {"method": "blueprint", "delta": delta}
            noise = CVLayer(code, delta=delta, p_swap=p_swap)
            t0 = time.perf_counter()
            if graph_only:
                eng = code.engine(delta, p_swap, wo, mode=mode)
                lo, hi = trial_range(trials, rank, size)
                dev_ms = 0.0
                for first in range(lo, hi, batch):
                    n = min(batch, hi - first)
                    eng.run_rng(seed, first, n, stage=_lib.STAGE_GRAPH, fetch=False)
                    dev_ms += eng.timings()["total_ms"]
                errors, rate = -1, (hi - lo) / (dev_ms * 1e-3) if dev_ms else 0.0
            else:
                errors, st = ec_monte_carlo(trials, code, noise, "MWPM", {"weight_opts": wo}, seed=seed, mode=mode,
                                            batch=batch, backend=backend, return_stats=True)
                rate = st["trials_done"] / size / (st["device_ms"] * 1e-3) if st["device_ms"] else 0.0
            dt = time.perf_counter() - t0
            row = ["SurfaceCode", d, ec, boundaries, "CVLayer", delta, p_swap, "MWPM", '"%s"' % wo, errors, trials,
                   time.strftime("%H:%M:%S"), "%f" % dt, size, mode, size, "%.1f" % rate]
            rows.append(row)
            code.release_engines()  # a finished (delta, weights) point frees its device buffers
            if rank == 0 and fname:
                new = not os.path.exists(fname)
                with open(fname, "a", encoding="utf8") as f:
                    if new:
                        f.write(",".join(COLUMNS) + "\n")
                    f.write(",".join(str(v) for v in row) + "\n")
    return rows


def main(argv=None):
    ap = argparse.ArgumentParser(description="Threshold sweep over distances and deltas.")
    ap.add_argument("-distances", type=str, default="9,13,17,25")
    ap.add_argument("-deltas", type=str, default=",".join("%.3f" % v for v in np.linspace(0.06, 0.12, 13)))
    ap.add_argument("-trials", type=int, default=1000)
    ap.add_argument("-p_swap", type=float, default=0.0)
    ap.add_argument("-boundaries", type=str, default="open")
    ap.add_argument("-ec", type=str, default="primal")
    ap.add_argument("-mode", type=str, default="fresh",
                    help="fresh (default): a new CVLayer per trial - what a threshold estimate needs; compat: the "
                         "reference's reused-layer carry-over (with p_swap=0 every second trial is noise-free)")
    ap.add_argument("-fname", type=str, default=None)
    ap.add_argument("-graph_only", action="store_true")
    ap.add_argument("-batch", type=int, default=1024)
    a = ap.parse_args(argv)
    rows = sweep([int(v) for v in a.distances.split(",")], [float(v) for v in a.deltas.split(",")], a.trials,
                 a.p_swap, a.boundaries, a.ec, a.mode, a.fname, a.graph_only, batch=a.batch)
    if _dist_info()[1] == 0:
        for r in rows:
            print(",".join(str(v) for v in r))


if __name__ == "__main__":
    main(sys.argv[1:])
