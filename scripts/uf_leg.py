"""Union-Find leg of bench.py, run in a process of its own (its own CUDA context) so that nothing it does can disturb
the headline measurement: complete error-correction trials per second with ``decoder="UF"`` on the bench's lattice,
first with the host decoder (csrc/uf.cpp), then with the device decoder (k_uf, csrc/fpb_uf.cuh), which is first checked
against the host decoder on one batch (equal grown sets, syndrome cleared).  One JSON line per leg on stdout.
    python scripts/uf_leg.py --distance 13 --ec both --boundaries periodic --delta 0.09 --p-swap 0 --mode compat --trials 32768"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--distance", type=int, default=13)
    ap.add_argument("--ec", default="both")
    ap.add_argument("--boundaries", default="periodic")
    ap.add_argument("--delta", type=float, default=0.09)
    ap.add_argument("--p-swap", type=float, default=0.0)
    ap.add_argument("--mode", default="compat")
    ap.add_argument("--trials", type=int, default=32768)
    ap.add_argument("--batch", type=int, default=4096)
    a = ap.parse_args()

    import numpy as np

    from flamingpy_b200 import _lib, uf_native
    from flamingpy_b200.codes import SurfaceCode
    from flamingpy_b200.decoder import uf_erasures
    from flamingpy_b200.noise import CVLayer
    from flamingpy_b200.simulations import ec_monte_carlo

    code = SurfaceCode(a.distance, a.ec, a.boundaries)
    ps = int(a.p_swap) if a.p_swap in (0.0, 1.0) else a.p_swap
    layer = CVLayer(code, delta=a.delta, p_swap=ps)
    args = {"weight_opts": {"method": "uniform"}}
    common = {"api": "simulations.ec_monte_carlo(trials, SurfaceCode, CVLayer, 'UF', uniform weights): complete "
                     "error-correction trials incl. cluster growth, peeling and logical check", "unit": "trials/s",
              "trials": a.trials, "batch": a.batch, "mode": a.mode}

    def timed(backend):
        ec_monte_carlo(2 * a.batch, code, layer, "UF", args, seed=1, batch=a.batch, mode=a.mode, backend=backend)  # warm-up
        t0 = time.perf_counter()
        fails, st = ec_monte_carlo(a.trials, code, layer, "UF", args, seed=2, batch=a.batch, mode=a.mode, backend=backend,
                                   return_stats=True)
        dt = time.perf_counter() - t0
        return {"value": a.trials / dt, "failures": int(fails), "device_ms": round(float(st["device_ms"]), 1), **common}

    try:
        out = {"leg": "host", "decoder": "csrc/uf.cpp on %d host threads" % uf_native.load().fpb_uf_max_threads(), **timed("native")}
    except Exception as exc:  # noqa: BLE001 - reported, not raised: this is a side measurement
        out = {"leg": "host", "error": repr(exc)[:300]}
    print(json.dumps(out), flush=True)

    try:
        eng = code.engine(a.delta, ps, {"method": "uniform"}, mode=a.mode)
        T = 256
        eng.run_rng(3, 0, T, stage=_lib.STAGE_SYNDROME, fetch=False)
        flips, grown = eng.uf_decode(use_erasures=bool(ps), want_grown=True)
        _, parity, st = eng.fetch_syndrome(want_state=bool(ps))
        er = uf_erasures(code, st)
        same_grown = cleared = True
        off = 0
        for i, e in enumerate(code.ec):
            ct = code.lattice.complexes[e]
            _, hg = uf_native.decode_batch(ct, parity[i], None if er is None else er[:, off:off + ct.Q], want_grown=True)
            same_grown &= bool(np.array_equal(grown[:, off:off + ct.Q], hg))
            par = np.zeros((T, ct.S), np.uint8)
            for f in range(6):
                q = ct.cube_qubit[:, f]
                has = q >= 0
                par[:, has] ^= flips[:, off:off + ct.Q][:, q[has]]
            cleared &= bool(np.array_equal(par, parity[i]))
            off += ct.Q
        out = {"leg": "device", "decoder": "k_uf (csrc/fpb_uf.cuh), one CTA per trial and complex",
               "grown_edges_equal_host_decoder": same_grown, "recovery_clears_syndrome": cleared, "checked_trials": T}
        if same_grown and cleared:
            out.update(timed("device"))
    except Exception as exc:  # noqa: BLE001
        out = {"leg": "device", "error": repr(exc)[:300]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
