"""Monte-Carlo error-correction driver (``flamingpy/simulations.py``).

``ec_monte_carlo`` keeps the reference's signature
(``simulations.py:62-116``) but runs the trials in device batches: noise,
weights, syndromes and matching graphs come from the CUDA pipeline (Philox
stream keyed by the global trial index), the host solves the matchings, the
GPU extracts the correction paths, and failure counts are summed - across
ranks with one ``all_reduce`` when ``torch.distributed`` is initialised (the
reference's ``MPI.Reduce``, ``simulations.py:110``).

``mode="compat"`` (default) reproduces the reference's reused-``CVLayer``
carry-over (SURVEY.md finding 0.3); ``mode="fresh"`` draws every trial from a
fresh layer.
"""

from __future__ import annotations

import argparse
import sys
from ast import literal_eval as l_eval
from datetime import datetime
from time import perf_counter

import numpy as np

from . import _lib
from .codes import SurfaceCode
from .decoder import candidates, correct, logical_failures, match_batch
from .engine import TrialBatch
from .noise import CVLayer, CVMacroLayer, IidNoise

int_time = int(str(datetime.now().timestamp()).replace(".", ""))


def ec_mc_trial(code_instance, noise_instance, decoder, weight_opts, rng=None, decoder_opts=None):
    """One trial through the reference-style objects (``simulations.py:46-59``)."""
    noise_instance.apply_noise(np.random.default_rng() if rng is None else rng)
    return correct(code=code_instance, decoder=decoder, weight_options=weight_opts, decoder_opts=decoder_opts)


def _dist_info():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except ImportError:  # pragma: no cover
        pass
    return None, 0, 1


def trial_range(trials: int, rank: int, size: int):
    """Contiguous block of the global trial range owned by ``rank`` (the reference deals
    trials round-robin, ``simulations.py:98-99``; blocks keep a compat chain on one rank)."""
    base, rem = divmod(trials, size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def ec_monte_carlo(trials, code_instance, noise_instance, decoder, decoder_args, world_comm=None, mpi_rank=0,
                   mpi_size=1, *, seed=None, batch=1024, mode="compat", backend="native", device=None,
                   return_stats=False, handoff=None, match_workers=None, reserve_sms=None):
    """Run ``trials`` Monte-Carlo trials of the complete error-correction procedure and
    return the number of logical errors.

    ``handoff``: how the matching graphs reach the host matcher.  "full": every length is copied to
    the host (the contract outputs of ``fpb_fetch``).  "sparse" (native matcher only): the K(K-1)/2
    lengths stay in HBM, the matcher receives candidate partners with lengths and prices the complete
    graph on the device (``fpb_knn_len`` / ``fpb_pairs_at`` / ``fpb_price``) - the same exact optimum
    for ~7 % of the bytes over the bus, and no host pass over the triangle.  Default: "sparse" on
    lattices without boundary points, "full" otherwise (measured on one B200 + 16 cores, complete EC
    trials per second: d=13 periodic 9.3 k sparse / 8.8 k full; d=13 open 7.3 k / 8.4 k - the boundary
    chain and the tight connector pairs cost the sparse path more round trips).
    ``match_workers``: host threads that match different batches at the same time, each with its
    share of the cores (default 3 for the sparse hand-off, whose pricing rounds are batch-synchronous,
    2 for the full one).  ``reserve_sms``: SMs the persistent shortest-path kernel leaves free for the
    matcher's pricing kernels (default 4 with the sparse hand-off, else 0).
    ``decoder="UF"``: ``backend="native"`` (default) grows and peels the clusters on the host cores
    (``csrc/uf.cpp``); ``backend="device"`` does it on the GPU (``fpb_uf_decode``, one CTA per trial
    and complex), so that only bits and flips cross the bus.

    Trial split: with ``torch.distributed`` initialised, over its ranks with one ``all_reduce`` of
    the counts (one process per GPU; NCCL).  Otherwise, if an mpi4py-style ``world_comm`` is given
    with ``mpi_size > 1``, over the MPI ranks with ``world_comm.allreduce`` (the reference's
    ``Reduce``, ``simulations.py:98-110``; every rank gets the total).  Without either, ``mpi_rank``
    / ``mpi_size`` are ignored and this process runs all ``trials``.
    """
    if decoder not in ("MWPM", "UF"):
        raise KeyError(decoder)  # the reference's decoder_dict lookup (decoders/decoder.py:279-286)
    if backend not in (("native", "device") if decoder == "UF" else ("native", "networkx")):
        raise ValueError(f"unknown backend {backend!r} for decoder {decoder!r}")
    if not isinstance(noise_instance, (CVLayer, CVMacroLayer, IidNoise)):
        raise NotImplementedError("the batched driver accelerates CVLayer, CVMacroLayer and IidNoise")
    weight_opts = dict(decoder_args.get("weight_opts") or {})
    iid = isinstance(noise_instance, IidNoise)
    macro = isinstance(noise_instance, CVMacroLayer)
    if decoder == "UF" and weight_opts.get("method", "uniform") == "blueprint":
        raise Exception("Incompatible decoder & weight options combination.")  # noqa: TRY002 - decoder.py:94-95
    if macro and weight_opts.get("method", "uniform") == "blueprint" and not weight_opts.get("prob_precomputed"):
        raise ValueError("CVMacroLayer leaves no homodyne values on the reduced lattice: use weight_opts "
                         "{'method': 'blueprint', 'prob_precomputed': True} (or 'uniform')")
    if not macro and weight_opts.get("prob_precomputed"):
        raise ValueError("prob_precomputed weights need the p_phase_cond attribute that CVMacroLayer computes")
    if iid:
        if weight_opts.get("method", "uniform") != "uniform":
            raise ValueError("IidNoise has no homodyne outcomes: use weight_opts {'method': 'uniform'}")
        n_delta, n_pswap, mode = 1.0, 0.0, "fresh"  # engine parameters that bit-level trials never read
    else:
        n_delta, n_pswap = noise_instance.delta, noise_instance.p_swap
        if not macro and not n_pswap and len(np.atleast_1d(noise_instance.states.get("p", []))):
            # the device Philox path draws labels from p_swap only: a layer with hand-picked
            # p-squeezed nodes would silently run as all-GKP
            raise NotImplementedError(
                "ec_monte_carlo does not batch a CVLayer built with explicit states={'p': [...]}: "
                "use p_swap, or loop ec_mc_trial (apply_noise + correct) for such a layer")
    dist, rank, size = _dist_info()
    use_mpi = dist is None and world_comm is not None and mpi_size > 1
    if use_mpi:
        rank, size = mpi_rank, mpi_size
    if device is None:
        device = 0
        try:
            import torch

            if torch.cuda.is_available():
                device = torch.cuda.current_device()
        except ImportError:  # pragma: no cover
            pass
    seed = int_time if seed is None else int(seed)
    if decoder == "UF":
        # Union-Find: the device pipeline stops after the syndrome (no matching graph); clusters are
        # grown and peeled on the host cores (csrc/uf.cpp) or by k_uf on the device, one batch at a time
        from .decoder import uf_decode_batch, uf_erasures

        eng = code_instance.engine(n_delta, n_pswap, {"method": "uniform"}, device=device, mode=mode)
        lo, hi = trial_range(trials, rank, size)
        failures = done = 0
        t_dev = 0.0
        for first in range(lo, hi, batch):
            n = min(batch, hi - first)
            if iid:
                eng.run_iid(seed, first, n, noise_instance.error_probability, stage=_lib.STAGE_SYNDROME, fetch=False)
            elif macro:
                eng.run_macro_rng(seed, first, n, stage=_lib.STAGE_SYNDROME, fetch=False)
            else:
                eng.run_rng(seed, first, n, stage=_lib.STAGE_SYNDROME, fetch=False)
            if backend == "device":  # growth, peeling and the logical check on the GPU: one count crosses the bus
                eng.uf_decode(use_erasures=not iid and (macro or bool(n_pswap)), fetch=False)
                failures += eng.logical_check(want_trials=False)
                done += n
                t_dev += eng.timings()["total_ms"]
                continue
            else:
                bits, parity, st = eng.fetch_syndrome(want_state=not iid and not macro and bool(n_pswap), packed_bits=True)
                if macro:
                    st = eng.fetch_macro()["red_state"]
                flips = uf_decode_batch(code_instance, parity, uf_erasures(code_instance, st))
            failures += int(logical_failures(code_instance, TrialBatch(n, _lib.STAGE_SYNDROME, bits=bits), flips).sum())
            done += n
            t_dev += eng.timings()["total_ms"]
        return _reduce_counts(failures, done, trials, dist, rank, size, use_mpi, world_comm, device, return_stats,
                              {"device_ms": t_dev, "handoff": "none", "uf_backend": "device" if backend == "device" else "host"})
    if handoff is None:
        periodic = not any(code_instance.lattice.complexes[e].has_boundary for e in code_instance.ec)
        handoff = "sparse" if (backend == "native" and periodic) else "full"
    if match_workers is None:
        match_workers = 3 if handoff == "sparse" else 2
    if handoff not in ("sparse", "full") or (handoff == "sparse" and backend != "native"):
        raise ValueError("handoff must be 'sparse' (native matcher only) or 'full'")
    code = code_instance
    # sparse hand-off: the matcher's pricing kernels run while the next batch's persistent shortest-path
    # kernel holds the SMs; leaving it a few SMs short costs the graph stage ~3 % and the matcher no wait
    # (measured at 0 vs 4; with the faster matcher of the end of round 2 the waits for pricing rounds are as long as
    # the solves, profiles/r2_final_checks.md - `reserve_sms` is the knob to sweep next)
    rsv = (4 if handoff == "sparse" else 0) if reserve_sms is None else int(reserve_sms)
    eng = code.engine(n_delta, n_pswap, weight_opts, device=device, mode=mode, reserve_sms=rsv)
    if iid and not eng.uses_table:
        eng.build_table()  # uniform weights: every matching graph is a gather from one all-pairs table
    lo, hi = trial_range(trials, rank, size)
    # Three stages on three threads, three engines in rotation: (1) device stage of batch k + fetch of
    # what the decoder needs + candidate partners; (2) host matching of batch k - 1 on all cores;
    # (3) correction paths of batch k - 2 on the device + logical check.  An engine starts its next
    # batch only after stage 3 has released it (its last run is what paths_toggle works on).
    import queue
    import threading

    starts = list(range(lo, hi, batch))
    # `match_workers` host threads match different batches at the same time, each with its share of the
    # host cores: the serial parts of one batch's matching (hand-over, late pricing rounds, stragglers)
    # overlap the parallel part of the other's.  One lane (engine + buffers) per batch in flight.
    match_workers = max(1, min(int(match_workers), len(starts))) if backend == "native" else 1
    n_eng = min(2 + match_workers, max(1, len(starts)))
    engines = [eng] + [code.engine(n_delta, n_pswap, weight_opts, device=device, mode=mode, lane=i, reserve_sms=rsv)
                       for i in range(1, n_eng)]
    # pinned host buffers, one set per lane, kept on the code object: reused by every batch, every
    # later call and every other noise / weight setting of a sweep (pinning memory is slow)
    lane_bufs = [code.__dict__.setdefault("_decode_bufs", {}).setdefault((device, i), {}) for i in range(n_eng)]
    knn_bufs = [code.__dict__.setdefault("_knn_bufs", {}).setdefault((device, i), {}) for i in range(n_eng)]
    free = [threading.Semaphore(1) for _ in engines]
    q_graph, q_match = queue.Queue(maxsize=2), queue.Queue(maxsize=2)
    state = {"failures": 0, "done": 0, "t_dev": 0.0, "error": None, "match": {}}

    def produce():
        try:
            for k, first in enumerate(starts):
                lane = k % n_eng
                free[lane].acquire()
                if state["error"] is not None:
                    break
                en, n = engines[lane], min(batch, hi - first)
                if iid:
                    en.run_iid(seed, first, n, noise_instance.error_probability, stage=_lib.STAGE_GRAPH, fetch=False)
                elif macro:
                    en.run_macro_rng(seed, first, n, stage=_lib.STAGE_GRAPH, fetch=False)
                else:
                    en.run_rng(seed, first, n, stage=_lib.STAGE_GRAPH, fetch=False)
                b = en.fetch_decode(lane_bufs[lane], want_pairs=handoff == "full", packed_bits=True)
                if backend != "native":
                    cands = None
                elif handoff == "full":
                    cands = [candidates(en, b, ci, e) for ci, e in enumerate(code.ec)]
                else:  # candidates with lengths: what the sparse hand-off starts from
                    from . import mwpm_native

                    cands = [en.knn_len_pinned(ci, mwpm_native.CAND_K, knn_bufs[lane]) for ci in range(len(code.ec))]
                q_graph.put((lane, b, cands))
        except BaseException as exc:  # noqa: BLE001 - re-raised by the caller's thread
            state["error"] = exc
        q_graph.put(None)

    def finish():
        while True:
            item = q_match.get()
            if item is None:
                return
            lane, b, queries = item
            try:
                if state["error"] is None:
                    flips = engines[lane].paths_toggle(*queries, packed=True)  # words in, words out: nothing unpacked
                    state["failures"] += int(logical_failures(code, b, flips).sum())
                    state["done"] += b.n_trials
                    state["t_dev"] += b.timings["total_ms"]
            except BaseException as exc:  # noqa: BLE001
                state["error"] = exc
            free[lane].release()

    from . import mwpm_native

    threads_each = max(1, mwpm_native.default_threads() // match_workers)
    stats_lock = threading.Lock()

    def match():
        while True:
            item = q_graph.get()
            if item is None:
                q_graph.put(None)  # pass the end marker on to the other workers
                return
            lane, b, cands = item
            queries = None
            try:
                if state["error"] is None:
                    st = {}
                    queries = match_batch(code, engines[lane], b, backend=backend, cands=cands, handoff=handoff,
                                          stats=st, threads=threads_each)
                    with stats_lock:
                        for k, v in st.items():
                            state["match"][k] = state["match"].get(k, 0) + v
            except BaseException as exc:  # noqa: BLE001
                state["error"] = exc
            if queries is None:
                free[lane].release()
            else:
                q_match.put((lane, b, queries))

    t_prod = threading.Thread(target=produce)
    t_fin = threading.Thread(target=finish)
    t_match = [threading.Thread(target=match) for _ in range(match_workers - 1)]
    for th in [t_prod, t_fin] + t_match:
        th.start()
    match()  # this thread is a matching worker, too
    for th in t_match:
        th.join()
    q_match.put(None)
    t_prod.join()
    t_fin.join()
    if state["error"] is not None:
        raise state["error"]
    return _reduce_counts(state["failures"], state["done"], trials, dist, rank, size, use_mpi, world_comm, device,
                          return_stats, {"device_ms": state["t_dev"], "handoff": handoff, **state["match"]})


def _reduce_counts(failures, done, trials, dist, rank, size, use_mpi, world_comm, device, return_stats, stats):
    """Sum (failures, trials done) over the ranks and return the failure count (with ``stats``)."""
    counts = np.array([failures, done], np.int64)
    if dist is not None and size > 1:
        import torch

        dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
        tc = torch.tensor(counts, dtype=torch.int64, device=dev)
        dist.all_reduce(tc, op=dist.ReduceOp.SUM)
        counts = tc.cpu().numpy()
    elif use_mpi:
        counts = np.array(world_comm.allreduce([int(failures), int(done)]) if hasattr(world_comm, "allreduce")
                          else [failures, done], np.int64).reshape(-1, 2).sum(axis=0)
    if int(counts[1]) != trials and (dist is not None or use_mpi or size == 1):
        raise RuntimeError(f"trial accounting: {int(counts[1])} trials completed, {trials} requested")
    errors = int(counts[0])
    if return_stats:
        return errors, {"trials_done": int(counts[1]), "rank": rank, "world_size": size, **stats}
    return errors


def write_csv_row(file_name, code_name, code_args, noise_name, noise_args, decoder, decoder_args, errors, trials,
                  seconds, mpi_size):
    """Append one result row in the reference's CSV format (``simulations.py:151-194``, pinned
    by its ``tests/frontend/test_simulations.py:124-174``): columns ``code``, the code argument names,
    ``noise``, the noise argument names, ``decoder``, the decoder argument names, then ``errors,
    trials, current_time, simulation_time, mpi_size``; the header is written when the file is new."""
    import csv
    import os

    header = (["code"] + list(code_args) + ["noise"] + list(noise_args) + ["decoder"] + list(decoder_args)
              + ["errors", "trials", "current_time", "simulation_time", "mpi_size"])
    row = ([code_name] + list(code_args.values()) + [noise_name] + list(noise_args.values()) + [decoder]
           + list(decoder_args.values())
           + [int(errors), int(trials), datetime.now().time().strftime("%H:%M:%S"), "%f" % seconds, int(mpi_size)])
    new_file = not os.path.exists(file_name)
    with open(file_name, "a", newline="", encoding="utf8") as f:
        w = csv.writer(f, lineterminator="\n")  # quotes the weight_opts dict, which contains commas
        if new_file:
            w.writerow(header)
        w.writerow(row)


def run_ec_simulation(trials, code, code_args, noise, noise_args, decoder, decoder_args, fname=None, **mc_kwargs):
    """``run_ec_simulation`` (``simulations.py:119-195``): build the objects, run the Monte
    Carlo, append one CSV row with the reference's columns."""
    start = perf_counter()
    code_instance = code(**code_args)
    noise_instance = noise(code_instance, **noise_args)
    _, rank, size = _dist_info()
    errors = ec_monte_carlo(trials, code_instance, noise_instance, decoder, decoder_args, None, rank, size,
                            **mc_kwargs)
    stop = perf_counter()
    if rank == 0:
        write_csv_row(fname or "./sims_results.csv", code.__name__, code_args, noise.__name__, noise_args, decoder,
                      decoder_args, errors, trials, stop - start, size)
    return errors


def main(argv=None):
    """Same flags as the reference CLI (``simulations.py:198-248``)."""
    parser = argparse.ArgumentParser(description="Arguments for Monte Carlo FT simulations.")
    parser.add_argument("-code", type=str, default="SurfaceCode")
    parser.add_argument("-code_args", type=str, default="{'distance': 3, 'ec': 'primal', 'boundaries': 'open'}")
    parser.add_argument("-noise", type=str, default="CVLayer")
    parser.add_argument("-noise_args", type=str, default="{'delta': 0.09, 'p_swap': 0.25}")
    parser.add_argument("-decoder", type=str, default="MWPM")
    parser.add_argument("-decoder_args", type=str, default="{'weight_opts': {'method': 'blueprint', 'delta': 0.09}}")
    parser.add_argument("-trials", type=int, default=100)
    parser.add_argument("-fname", type=str, default=None)
    parser.add_argument("-mode", type=str, default="compat")
    args = parser.parse_args(argv)
    if args.decoder.lower() in ["mwpm", "minimum weight perfect matching"]:
        dec = "MWPM"
    elif args.decoder.lower() in ["unionfind", "uf", "union-find", "union find"]:
        dec = "UF"
    else:
        raise ValueError(f"Decoder {args.decoder} is either invalid or not yet implemented.")
    classes = {"SurfaceCode": SurfaceCode, "CVLayer": CVLayer, "CVMacroLayer": CVMacroLayer, "IidNoise": IidNoise}
    return run_ec_simulation(args.trials, classes[args.code], l_eval(args.code_args), classes[args.noise],
                             l_eval(args.noise_args), dec, l_eval(args.decoder_args), fname=args.fname,
                             mode=args.mode)


if __name__ == "__main__":
    main(sys.argv[1:])
