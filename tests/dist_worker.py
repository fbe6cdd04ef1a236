"""Worker for tests/test_distributed_cpu.py: world_size-2 gloo run of the host-side
multi-GPU logic (trial partition + the single count all-reduce)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flamingpy_b200.simulations import _dist_info, trial_range  # noqa: E402


def main():
    dist.init_process_group("gloo")
    d, rank, size = _dist_info()
    assert d is not None and size == int(os.environ["WORLD_SIZE"])
    trials = 1001
    lo, hi = trial_range(trials, rank, size)
    # stand-in for the per-rank Monte-Carlo loop: "fail" every trial whose global index is 0 mod 7
    failures = sum(1 for t in range(lo, hi) if t % 7 == 0)
    counts = torch.tensor([failures, hi - lo], dtype=torch.int64)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    assert int(counts[1]) == trials
    assert int(counts[0]) == len(range(0, trials, 7))
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    assert float(tmax) == float(size)
    dist.barrier()
    if rank == 0:
        print("DIST_OK", int(counts[0]), int(counts[1]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
