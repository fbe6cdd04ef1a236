"""BASELINE.json configs[4]: threshold sweep over distances and deltas, trials split over the GPUs of one box
(one process per GPU, one NCCL all-reduce of the failure counts per point).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        scripts/sweep_dist.py -distances 9,13 -trials 100000 -fname gpurun_out/sweep.csv"""
import os, sys, time; sys.path.insert(0, ".")
import torch
import torch.distributed as dist
from flamingpy_b200 import sweep

multi = "RANK" in os.environ
if multi:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
t0 = time.perf_counter()
sweep.main(sys.argv[1:])
if not multi or dist.get_rank() == 0:
    print(f"SWEEP wall={time.perf_counter() - t0:.1f}s", flush=True)
if multi:
    dist.destroy_process_group()
