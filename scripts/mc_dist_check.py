"""ec_monte_carlo over torch.distributed/NCCL: the failure count must not depend on the number of
GPUs (Philox counters are keyed by the global trial index).
    python scripts/mc_dist_check.py                      # one process
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/mc_dist_check.py     # two ranks, one all_reduce of (failures, trials)"""
import os, sys, time; sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.noise import CVLayer
from flamingpy_b200.simulations import ec_monte_carlo

multi = "RANK" in os.environ
if multi:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
rank = dist.get_rank() if multi else 0
for d, bnd, ec, trials in [(9, "open", "primal", 4096), (13, "periodic", "both", 2048)]:
    code = SurfaceCode(d, ec, bnd)
    noise = CVLayer(code, delta=0.09, p_swap=0)
    wo = {"weight_opts": {"method": "blueprint", "delta": 0.09}}
    t0 = time.perf_counter()
    e, st = ec_monte_carlo(trials, code, noise, "MWPM", wo, seed=20261019, mode="fresh", batch=256, return_stats=True)
    if rank == 0:
        print(f"MC d={d} {bnd} ec={ec} world={st['world_size']} trials={st['trials_done']} errors={e} "
              f"in {time.perf_counter() - t0:.2f}s", flush=True)
if multi:
    dist.destroy_process_group()
