"""BASELINE.json configs[2]: SurfaceCode(d=13, periodic, ec='both'), CVLayer delta=0.09 p_swap=0, blueprint
weights, complete error-correction trials split over the GPUs of one box with one NCCL all-reduce of the counts.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        scripts/mc_config2.py 200000"""
import os, sys, time; sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from flamingpy_b200 import mwpm_native
from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.noise import CVLayer
from flamingpy_b200.simulations import ec_monte_carlo

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
multi = "RANK" in os.environ
if multi:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
rank = dist.get_rank() if multi else 0
world = dist.get_world_size() if multi else 1
code = SurfaceCode(13, "both", "periodic")
noise = CVLayer(code, delta=0.09, p_swap=0)
wo = {"weight_opts": {"method": "blueprint", "delta": 0.09}}
ec_monte_carlo(1024 * world, code, noise, "MWPM", wo, seed=1, batch=256)  # warm-up: engines, pinned buffers, NCCL
if multi:
    dist.barrier()
t0 = time.perf_counter()
errors, st = ec_monte_carlo(trials, code, noise, "MWPM", wo, seed=20261019, batch=256, return_stats=True)
dt = time.perf_counter() - t0
if rank == 0:
    print(f"CONFIG2 gpus={world} host_threads_per_rank={mwpm_native.default_threads()} trials={st['trials_done']} "
          f"errors={errors} rate={errors / st['trials_done']:.5f} wall={dt:.1f}s throughput={st['trials_done'] / dt:.0f} trials/s "
          f"(compat mode: every second trial of a chain is noise-free)", flush=True)
if multi:
    dist.destroy_process_group()
