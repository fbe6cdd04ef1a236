#!/usr/bin/env python
"""Benchmark of the Monte-Carlo EC trial path (noise -> weights -> syndrome -> matching graph).

    python bench.py --gpus N --steps K --warmup W            # B200 arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port on host cores)

One step = one batch of `--batch` independent trials per GPU pushed through the whole hot path.
Prints one JSON line (rank 0).  See DESIGN.md "Measurement" for what every field means.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20261019

WORKLOADS = {
    # BASELINE.json configs[2]: the d=13 configuration the metric is quoted on
    "d13_periodic_both": dict(distance=13, ec="both", boundaries="periodic", delta=0.09, p_swap=0.0,
                              weight_opts={"method": "blueprint", "delta": 0.09}),
    # BASELINE.json configs[1]
    "d9_open_primal": dict(distance=9, ec="primal", boundaries="open", delta=0.08, p_swap=0.5,
                           weight_opts={"method": "blueprint", "delta": 0.08}),
    "d13_open_primal": dict(distance=13, ec="primal", boundaries="open", delta=0.09, p_swap=0.0,
                            weight_opts={"method": "blueprint", "delta": 0.09}),
    # BASELINE.json configs[3]: all nodes p-squeezed -> trial-independent weights -> table gather
    "d21_open_allp_blueprint": dict(distance=21, ec="primal", boundaries="open", delta=0.09, p_swap=1.0,
                                    weight_opts={"method": "blueprint", "delta": 0.09}),
    "d21_open_allp_uniform": dict(distance=21, ec="primal", boundaries="open", delta=0.09, p_swap=1.0,
                                  weight_opts={"method": "uniform"}),
    "d17_periodic_primal": dict(distance=17, ec="primal", boundaries="periodic", delta=0.09, p_swap=0.0,
                                weight_opts={"method": "blueprint", "delta": 0.09}),
    "d3_open_primal": dict(distance=3, ec="primal", boundaries="open", delta=0.09, p_swap=0.25,
                           weight_opts={"method": "blueprint", "delta": 0.09}),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [s.strip() for s in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def algorithmic_bytes(lat, n_trials, total_odd, total_pairs):
    """B_trial = 32 Q + (2N + 2Q + 2S)/8 + 4 K + 8 [K(K-1)/2 + K_b],
    summed exactly over the trials of a run (Q, S, K over both complexes, N once)."""
    Q = sum(lat.complexes[e].Q for e in lat.ec)
    S = sum(lat.complexes[e].S for e in lat.ec)
    static = n_trials * (32 * Q + (2 * lat.N + 2 * Q + 2 * S) / 8.0)
    dyn = 0.0
    for i, e in enumerate(lat.ec):
        kb = total_odd[i] if lat.complexes[e].has_boundary else 0
        dyn += 4 * total_odd[i] + 8 * (total_pairs[i] + kb)
    return static + dyn


def paths_kernel_bytes(lat, n_trials, total_odd, total_pairs):
    """Algorithmic bytes of the matching-graph kernel alone: it reads each trial's weights
    (8 Q) and odd list (4 K) once and writes the lengths 8 [K(K-1)/2 + K_b] once."""
    Q = sum(lat.complexes[e].Q for e in lat.ec)
    b = n_trials * 8.0 * Q
    for i, e in enumerate(lat.ec):
        kb = total_odd[i] if lat.complexes[e].has_boundary else 0
        b += 4 * total_odd[i] + 8 * (total_pairs[i] + kb)
    return b


def recorded_traffic(workload, mode, batch):
    """DRAM bytes per k_paths launch from the committed ncu capture, if it was taken on this
    very configuration (profiles/r2_k_paths_traffic.json); None otherwise."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_k_paths_traffic.json")) as f:
            t = json.load(f)
        if t["workload"] == workload and t["mode"] == mode and t["trials_per_gpu_per_step"] == batch:
            return t["traffic_bytes_per_launch"]
    except Exception:
        pass
    return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores
# --------------------------------------------------------------------------------------------
def host_inputs(lat, cfg, n_trials, seed, fresh):
    from flamingpy_b200.noise import sample_batch

    rng = np.random.default_rng(seed)
    return sample_batch(rng, lat.N, cfg["delta"], cfg["p_swap"], n_trials, fresh=fresh)


def time_cpu_port(lat, cfg, n_trials, fresh, threads=0, seed=SEED):
    """Time oracle/fpb_oracle.c (the CPU restatement of the reference path) on `n_trials` trials whose
    draws and labels are already in host memory - the same starting point as the B200 arm's `e2e`
    (host sampling is outside the timed region in both arms).  Returns (seconds, threads used)."""
    from oracle import c_oracle  # checker / baseline leg only

    c_oracle.load()
    if threads <= 0:
        threads = os.cpu_count() or 1  # explicit: torchrun exports OMP_NUM_THREADS=1
    x, st = host_inputs(lat, cfg, n_trials, seed, fresh)
    t0 = time.perf_counter()
    c_oracle.run(lat, cfg["delta"], cfg["p_swap"], cfg["weight_opts"], x, st, n_threads=threads, copy_out=False)
    dt = time.perf_counter() - t0
    return dt, threads


def workload_config(args, cfg):
    """The `config` object: the workload and nothing else, identical in both arms."""
    return {"workload": args.workload, **{k: v for k, v in cfg.items() if k != "weight_opts"},
            "weights": cfg["weight_opts"]["method"], "mode": args.mode,
            "span": "noise->weights->syndrome->matching graph (MWPM excluded)"}


def run_reference_arm(args, cfg, lat):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    fresh = args.mode == "fresh"
    cores = os.cpu_count() or 1
    per_step = args.cpu_trials if args.cpu_trials > 0 else max(8, 32 * cores)
    for _ in range(args.warmup):
        time_cpu_port(lat, cfg, max(2, per_step // 4), fresh)
    t_tot = 0.0
    threads = cores
    for s in range(args.steps):
        dt, threads = time_cpu_port(lat, cfg, per_step, fresh, seed=SEED + s)
        t_tot += dt
    value = per_step * args.steps / t_tot
    line = {
        "impl": "reference", "metric": "EC Monte Carlo trials/sec at d=%d" % lat.dims[0], "value": value,
        "unit": "trials/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, cfg),
        "cpu_baseline": {"value": value, "unit": "trials/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step} trials/step x {args.steps} steps, oracle/fpb_oracle.c (C + OpenMP "
                                   f"restatement of the reference path; the Python reference itself cannot run on "
                                   f"this box), draws and labels pre-sampled in host memory (as for the B200 arm's e2e)"},
        "e2e": {"value": value, "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def run_gpu_arm(args, cfg, lat):
    import torch

    from flamingpy_b200 import _lib
    from flamingpy_b200.engine import TrialEngine

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fresh = args.mode == "fresh"
    B = args.batch
    eng = TrialEngine(lat, cfg["delta"], cfg["p_swap"], cfg["weight_opts"], mode=args.mode, device=local,
                      delta_step=args.delta_step)
    static = cfg["weight_opts"]["method"] == "uniform" or cfg["p_swap"] == 1.0
    if static and not args.no_table:
        eng.build_table()  # trial-independent weights: the matching graph becomes a table gather

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    def first_trial(step):  # contiguous block per (step, rank): independent of the GPU count
        return (step * world + rank) * B

    # ---------------- device-resident throughput (`value`) ----------------
    for s in range(args.warmup):
        eng.run_rng(SEED, first_trial(s), B, fetch=False)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    dev_ms = graph_ms = 0.0
    launches = 0
    tot_odd = [0, 0]
    tot_pairs = [0, 0]
    t0 = time.perf_counter()
    for s in range(args.steps):
        eng.mark(0)
        eng.run_rng(SEED, first_trial(args.warmup + s), B, fetch=False)
        eng.mark(1)
        dev_ms += eng.elapsed_ms()
        tm = eng.timings()
        graph_ms += tm["graph_ms"]
        launches += tm["launches"]
        sz = eng.sizes()
        for i in range(sz.n_complex):
            tot_odd[i] += int(sz.total_odd[i])
            tot_pairs[i] += int(sz.total_pairs[i])
    barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None
    tmax = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    counts = torch.tensor([B * args.steps, sum(tot_odd)], dtype=torch.int64, device="cuda")
    if use_dist:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)  # the path's one collective (simulations.py:110)
    dev_ms_max, wall_ms_max = (float(v) for v in tmax.tolist())
    total_trials = int(counts[0].item())
    value = total_trials / (dev_ms_max * 1e-3)

    # ---------------- end to end through the public API with host buffers (`e2e`) ----------------
    # Every step copies its inputs from pinned host memory (H2D) and all contract outputs back to
    # pinned host buffers (D2H).  Two engines on two streams, driven by two host threads, let one
    # step's copies overlap the other's kernels (ctypes releases the GIL inside the calls).
    Be = args.e2e_batch
    n_lanes = max(1, args.e2e_lanes)
    n_sets = min(args.steps + args.warmup, 3)
    N, Q = lat.N, eng.Qtot
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()  # noqa: E731
    xs = [pin((Be, 2 * N), torch.float64) for _ in range(n_sets)]
    sts = [pin((Be, N), torch.uint8) for _ in range(n_sets)]
    from flamingpy_b200.noise import LabelSampler, sample_batch

    sampler = LabelSampler(N, cfg["p_swap"], fresh)
    rng = np.random.default_rng(SEED + 1000 * rank)
    for i in range(n_sets):
        sample_batch(rng, N, cfg["delta"], cfg["p_swap"], Be, sampler=sampler, out_x=xs[i], out_state=sts[i])
    engines = [eng]
    for _ in range(n_lanes - 1):
        e2 = TrialEngine(lat, cfg["delta"], cfg["p_swap"], cfg["weight_opts"], mode=args.mode, device=local,
                         delta_step=args.delta_step)
        if eng.uses_table:
            e2.build_table()
        engines.append(e2)
    # host result buffers sized from a probe run
    eng.run_injected(xs[0], sts[0], fetch=False)
    sz = eng.sizes()

    KNN = 12  # candidate partners per odd cube in the sparse hand-off (mwpm_native.CAND_K)

    def alloc_bufs(product):
        # "sparse": what the host matcher and the logical check need - bits, parities, odd lists, boundary
        # lengths and each cube's nearest partners with lengths; hom / weights / the K(K-1)/2 pair lengths stay
        # in HBM (the correction-path stage reads them there).  "contract": every output of fpb_fetch.
        bufs = {"bits": pin((Be * eng.bit_words,), torch.int32).view(np.uint32)}
        if product == "contract":
            bufs.update(hom=pin((Be * Q,), torch.float64), weights=pin((Be * Q,), torch.float64))
        for i, e in enumerate(lat.ec):
            ct = lat.complexes[e]
            mo = int(sz.total_odd[i] * 1.25) + 1024
            mp = int(sz.total_pairs[i] * 1.25) + 1024
            bufs[f"parity{i}"] = pin((Be * ((ct.S + 31) // 32),), torch.int32).view(np.uint32)
            bufs[f"odd_count{i}"] = pin((Be,), torch.int32)
            bufs[f"odd_ids{i}"] = pin((mo,), torch.int32)
            if product == "contract":
                bufs[f"pair_len{i}"] = pin((mp,), torch.float64)
            else:
                bufs[f"_knn_ids{i}"] = pin((mo * KNN,), torch.int32)
                bufs[f"_knn_len{i}"] = pin((mo * KNN,), torch.float64)
                bufs[f"_wmax{i}"] = pin((Be,), torch.float64)
            if ct.has_boundary:
                bufs[f"bnd_len{i}"] = pin((mo,), torch.float64)
                bufs[f"bnd_side{i}"] = pin((mo,), torch.uint8)
        return bufs

    lane_bufs = {prod: [alloc_bufs(prod) for _ in engines] for prod in ("sparse", "contract")}

    # A step streams its Be trials through the lanes in `chunks` sub-batches (host buffer slices), so
    # that the only copies that cannot overlap a kernel are the first sub-batch's H2D and the last one's
    # D2H (with whole steps as the pipeline unit that fill / drain is 11 % of a 5-step run)
    if args.e2e_chunks > 0:
        chunks = max(1, min(args.e2e_chunks, Be))
    else:  # auto: sub-batches of at least ~3 ms of device work (launch + size round trips stay small), at most 8
        ms_per_trial = dev_ms_max / max(1, B * args.steps)
        chunks = max(1, min(8, int(Be * ms_per_trial / 3.0)))
    bounds = [Be * c // chunks for c in range(chunks + 1)]

    def e2e_chunk(product, lane, i, c):
        en = engines[lane]
        bufs = lane_bufs[product][lane]
        lo_, hi_ = bounds[c], bounds[c + 1]
        nb = hi_ - lo_
        en.run_injected(xs[i % n_sets][lo_:hi_], sts[i % n_sets][lo_:hi_], fetch=False)  # H2D inside
        s_ = en.fetch_into({k: v for k, v in bufs.items() if not k.startswith("_")})  # D2H inside
        per_cx = lambda j, e: (nb * (4 * ((lat.complexes[e].S + 31) // 32) + 4) + 4 * int(s_.total_odd[j])  # noqa: E731
                               + (9 * int(s_.total_odd[j]) if lat.complexes[e].has_boundary else 0))
        if product == "contract":
            return nb * (16 * Q + 4 * eng.bit_words) + sum(per_cx(j, e) + 8 * int(s_.total_pairs[j])
                                                           for j, e in enumerate(lat.ec))
        for j in range(len(lat.ec)):  # candidates with lengths, D2H inside
            en.knn_len_into(j, KNN, bufs[f"_knn_ids{j}"], bufs[f"_knn_len{j}"], bufs[f"_wmax{j}"])
        return nb * 4 * eng.bit_words + sum(per_cx(j, e) + 12 * KNN * int(s_.total_odd[j]) + 8 * nb
                                            for j, e in enumerate(lat.ec))

    def run_steps(product, first, count):
        done = [0] * len(engines)

        def work(lane):
            for g in range(first * chunks + lane, (first + count) * chunks, len(engines)):
                done[lane] += e2e_chunk(product, lane, g // chunks, g % chunks)

        ths = [threading.Thread(target=work, args=(ln,)) for ln in range(len(engines))]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return sum(done)

    def time_e2e(product):
        run_steps(product, 0, max(len(engines), args.warmup))
        barrier()
        t0 = time.perf_counter()
        nbytes = run_steps(product, args.warmup, args.steps)
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if use_dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return Be * args.steps * world / float(t.item()), nbytes // max(1, args.steps)

    e2e_value, d2h_bytes = time_e2e("sparse")
    e2e_contract, d2h_contract = time_e2e("contract")
    h2d_per_step = Be * (2 * N * 8 + N)

    # ---------------- the other label mode, device-resident (compat halves the noisy trials) ----------------
    other = "fresh" if args.mode == "compat" else "compat"
    eng_o = TrialEngine(lat, cfg["delta"], cfg["p_swap"], cfg["weight_opts"], mode=other, device=local,
                        delta_step=args.delta_step)
    if eng.uses_table:
        eng_o.build_table()
    n_o = max(2, min(args.steps, 5))
    eng_o.run_rng(SEED, first_trial(0), B, fetch=False)
    barrier()
    o_ms = 0.0
    for s_i in range(n_o):
        eng_o.mark(0)
        eng_o.run_rng(SEED, first_trial(1 + s_i), B, fetch=False)
        eng_o.mark(1)
        o_ms += eng_o.elapsed_ms()
    eng_o.close()
    o_t = torch.tensor([o_ms], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.all_reduce(o_t, op=dist.ReduceOp.MAX)
    other_value = B * n_o * world / (float(o_t.item()) * 1e-3)

    # ---------------- complete EC trials through the public API (`e2e_full`) ----------------
    # simulations.ec_monte_carlo: device stage + host matcher on this rank's share of the host cores +
    # correction paths + logical check, the reference's Monte-Carlo loop (simulations.py:62-116)
    e2e_full = None
    if not args.no_e2e_full:
        from flamingpy_b200.codes import SurfaceCode
        from flamingpy_b200.noise import CVLayer
        from flamingpy_b200.simulations import ec_monte_carlo

        for en in engines[1:]:
            en.close()
        engines = engines[:1]
        code = SurfaceCode(cfg["distance"], cfg["ec"], cfg["boundaries"])
        noise = CVLayer(code, delta=cfg["delta"], p_swap=cfg["p_swap"])
        dargs = {"weight_opts": cfg["weight_opts"]}
        n_full = args.full_trials * world  # ec_monte_carlo splits the global count over the ranks
        # warm-up: enough batches to create every lane's engine and pinned buffers and to fill the matcher's pools
        ec_monte_carlo(8 * args.full_batch * world, code, noise, "MWPM", dargs, seed=SEED, batch=args.full_batch,
                       mode=args.mode)
        barrier()
        t0 = time.perf_counter()
        fails, st_full = ec_monte_carlo(n_full, code, noise, "MWPM", dargs, seed=SEED + 1, batch=args.full_batch,
                                        mode=args.mode, return_stats=True)
        barrier()
        f_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if use_dist:
            dist.all_reduce(f_t, op=dist.ReduceOp.MAX)
        from flamingpy_b200 import mwpm_native

        e2e_full = {"value": n_full / float(f_t.item()), "unit": "trials/s", "trials": n_full, "failures": int(fails),
                    "api": "simulations.ec_monte_carlo(trials, SurfaceCode, CVLayer, 'MWPM', weight_opts): complete "
                           "error-correction trials incl. host MWPM, correction paths and logical check",
                    "handoff": st_full["handoff"], "host_cores_per_rank": mwpm_native.default_threads(),
                    "batch": args.full_batch}
        code.release_engines()

    if rank == 0:
        peak, peak_src = load_peaks()
        kb = paths_kernel_bytes(lat, B * args.steps, tot_odd, tot_pairs)
        if eng.uses_table:
            # gather: one length written per pair / connector, the odd list read once.  The entries of
            # the static all-pairs table come through L2 and count 0 like every static table
            kb = sum(4 * tot_odd[i] + 8 * tot_pairs[i] + (9 * tot_odd[i] if lat.complexes[e].has_boundary else 0)
                     for i, e in enumerate(lat.ec))
        achieved = kb / (graph_ms * 1e-3) / 1e9 if graph_ms > 0 else 0.0
        pipe_b = algorithmic_bytes(lat, B * args.steps, tot_odd, tot_pairs)
        line = {
            "metric": "EC Monte Carlo trials/sec at d=%d" % lat.dims[0], "value": value, "unit": "trials/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, cfg),
            "run": {"trials_per_gpu_per_step": B, "rng": "Philox4x32-10 keyed by (seed, global trial, node)",
                    "l2": "per-step inputs and outputs exceed L2 (%.0f MB written per step)" % (
                        pipe_b / args.steps / 1e6),
                    "odd_cubes_per_trial": sum(tot_odd) / max(1, B * args.steps),
                    "binding": eng.binding,
                    "host_sampling": "outside the timed region in both arms (draws and labels start in host memory)"},
            "value_other_mode": {"mode": other, "value": other_value, "unit": "trials/s",
                                 "note": "compat = the reference's reused-CVLayer behaviour: with p_swap = 0 every "
                                         "second trial of a chain is noise-free; fresh = a new layer per trial"},
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": {"value": e2e_value, "unit": "trials/s", "h2d_bytes_per_step": h2d_per_step,
                    "d2h_bytes_per_step": d2h_bytes, "trials_per_gpu_per_step": Be,
                    "chunks_per_step": chunks,
                    "handoff": "sparse: per odd cube its %d nearest partners with lengths (fpb_knn_len) + odd lists, "
                               "boundary lengths, bits, parities; hom / weights / the K(K-1)/2 lengths stay in HBM for "
                               "fpb_price / fpb_pairs_at / fpb_paths_toggle" % KNN,
                    "api": "TrialEngine.run_injected(host x, host labels) + fetch_into + knn_len_into(pinned host "
                           "buffers), per sub-batch",
                    "host_threads": n_lanes},
            "e2e_contract": {"value": e2e_contract, "unit": "trials/s", "h2d_bytes_per_step": h2d_per_step,
                             "d2h_bytes_per_step": d2h_contract,
                             "handoff": "full: every output of fpb_fetch incl. hom, weights and all pair lengths "
                                        "(round 1's e2e definition)"},
            "e2e_full": e2e_full,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_gather" if eng.uses_table else "k_paths", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": None if eng.uses_table else recorded_traffic(args.workload, args.mode, B),
                         "algorithmic_bytes_per_launch": kb / args.steps, "peak_source": peak_src,
                         "share_of_step": graph_ms / dev_ms if dev_ms else None,
                         "note": ("static-weight table gather: bound by the HBM writes of the lengths and the L2 reads of "
                                  "the table entries (not counted)" if eng.uses_table else
                                  "k_paths is bound by shared-memory latency/issue, not HBM (DESIGN.md)"),
                         "pipeline_achieved": pipe_b / (dev_ms * 1e-3) / 1e9,
                         "pipeline_frac": pipe_b / (dev_ms * 1e-3) / 1e9 / peak},
            "clocks": clk,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n_cpu = args.cpu_trials if args.cpu_trials > 0 else max(8, 64 * cores)
            dt, threads = time_cpu_port(lat, cfg, n_cpu, fresh)
            line["cpu_baseline"] = {
                "value": n_cpu / dt, "unit": "trials/s", "cores": threads, "kind": "port",
                "sample": f"{n_cpu} trials of the same workload, oracle/fpb_oracle.c (C + OpenMP restatement of "
                          f"the reference path), draws and labels pre-sampled in host memory, {dt:.1f} s"}
        if not args.no_uf_leg and world == 1:
            line["uf"] = run_uf_leg(args, cfg)
        print(json.dumps(line), flush=True)
    for en in engines:
        en.close()
    if use_dist:
        dist.destroy_process_group()


def run_uf_leg(args, cfg):
    """Complete EC trials per second with decoder="UF" on the bench's lattice (uniform weights), host decoder and
    device decoder, measured by scripts/uf_leg.py in a child process: the device decoder (k_uf) was written after the
    last GPU session of its round, so it runs in a CUDA context of its own where a fault cannot touch the numbers
    above; a side measurement - errors are reported in the object, never raised."""
    import subprocess

    cmd = [sys.executable, os.path.join(ROOT, "scripts", "uf_leg.py"), "--distance", str(cfg["distance"]), "--ec", cfg["ec"],
           "--boundaries", cfg["boundaries"], "--delta", str(cfg["delta"]), "--p-swap", str(cfg["p_swap"]),
           "--mode", args.mode, "--trials", str(args.uf_trials)]
    out = {}
    try:
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
        for ln in r.stdout.splitlines():
            try:
                d = json.loads(ln)
            except ValueError:
                continue
            if isinstance(d, dict) and "leg" in d:
                out[d.pop("leg")] = d
        if r.returncode != 0:
            out["exit"] = {"returncode": r.returncode, "stderr_tail": r.stderr[-300:]}
    except Exception as exc:  # noqa: BLE001
        out["error"] = repr(exc)[:300]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="d13_periodic_both", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="compat", choices=["compat", "fresh"],
                    help="compat = one reused CVLayer per chain like simulations.ec_monte_carlo; fresh = new layer per trial")
    ap.add_argument("--batch", type=int, default=2048, help="trials per GPU per step (device-resident arm)")
    ap.add_argument("--e2e-batch", type=int, default=512, help="trials per GPU per step (end-to-end arm)")
    ap.add_argument("--e2e-chunks", type=int, default=0,
                    help="sub-batches a step's trials are streamed in (end-to-end arm); 1 = whole steps, 0 = auto")
    ap.add_argument("--e2e-lanes", type=int, default=2, help="engines / host threads overlapping copies and kernels")
    ap.add_argument("--delta-step", type=float, default=0.0)
    ap.add_argument("--cpu-trials", type=int, default=0, help="CPU sample size in trials (0: 64 x cores for cpu_baseline, 32 x cores per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-full", action="store_true", help="skip the complete-EC-trial measurement (ec_monte_carlo)")
    ap.add_argument("--full-trials", type=int, default=32768, help="complete EC trials per GPU for e2e_full")
    ap.add_argument("--full-batch", type=int, default=512)
    ap.add_argument("--no-uf-leg", action="store_true", help="skip the Union-Find side measurement (scripts/uf_leg.py, N = 1 only)")
    ap.add_argument("--uf-trials", type=int, default=32768, help="complete EC trials of each Union-Find leg")
    ap.add_argument("--no-table", action="store_true", help="static-weight workloads: search instead of the table gather")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: note: fewer than 3 warm-up steps requested", file=sys.stderr)

    from flamingpy_b200.lattice import build_lattice

    cfg = WORKLOADS[args.workload]
    lat = build_lattice(cfg["distance"], cfg["ec"], cfg["boundaries"])
    if args.impl == "reference":
        run_reference_arm(args, cfg, lat)
    else:
        run_gpu_arm(args, cfg, lat)


if __name__ == "__main__":
    main()
