"""TEST / BASELINE TOOL (build container only: needs /root/reference) - the reference's own host path on
points of the threshold sweep (BASELINE.json configs[4]), timed beside the B200 sweep at a stated, reduced trial
count.  P worker processes stand in for the reference's mpi4py ranks (mpirun is not installed): each runs the
reference's loop body - `noise.apply_noise(rng)`, `correct(code, "MWPM", weight_options, networkx backend)`
(flamingpy/simulations.py:46-59,98-110) - on its share of the trials.
    python scripts/ref_host_sweep.py 9 8 0.06,0.09,0.12        # distance, trials per process, deltas"""
import multiprocessing as mp, os, sys, time
sys.path.insert(0, ".")


def work(args):
    d, delta, n, seed = args
    import numpy as np
    from oracle import ref_shim
    ref_shim.load()
    from flamingpy.decoders.decoder import correct
    from flamingpy.noise import CVLayer
    t0 = time.perf_counter()
    code = ref_shim.make_code(d, "primal", "open")
    t_setup = time.perf_counter() - t0
    layer = CVLayer(code, delta=delta, p_swap=0)
    rng = np.random.default_rng(seed)
    wo = {"method": "blueprint", "integer": False, "multiplier": 1, "delta": delta}
    t0 = time.perf_counter()
    ok = 0
    for _ in range(n):
        layer.apply_noise(rng)
        ok += bool(correct(code, "MWPM", weight_options=wo, decoder_opts={"backend": "networkx"}))
    return n - ok, time.perf_counter() - t0, t_setup


if __name__ == "__main__":
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    deltas = [float(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0.06,0.09,0.12").split(",")]
    P = os.cpu_count() or 1
    print(f"reference host path (networkx backend), d={d} open primal p_swap=0 blueprint, {P} processes x {n} trials per point")
    with mp.Pool(P) as pool:
        for delta in deltas:
            t0 = time.perf_counter()
            res = pool.map(work, [(d, delta, n, 1000 + r) for r in range(P)])
            wall = time.perf_counter() - t0
            errs = sum(r[0] for r in res)
            loop = max(r[1] for r in res)
            setup = max(r[2] for r in res)
            print(f"REFHOST d={d} delta={delta:.3f} trials={P * n} errors={errs} loop_wall={loop:.1f}s "
                  f"code_setup={setup:.1f}s -> {P * n / loop:.3f} trials/s on {P} cores ({loop / n:.2f} s per trial per core)",
                  flush=True)
