"""Worker for tests/test_distributed_cpu.py: the real ``ec_monte_carlo`` driver (Union-Find decoder, whose
host side needs no GPU) under world_size-2 gloo, with the device calls answered by the C-oracle engine
stand-in of tests/test_uf.py.  Each rank decodes its half of the trial range; the one all-reduce returns
the whole-job failure count on every rank."""
import os
import sys

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flamingpy_b200.codes import SurfaceCode  # noqa: E402
from flamingpy_b200.noise import CVLayer  # noqa: E402
from flamingpy_b200.simulations import ec_monte_carlo  # noqa: E402
from tests.test_uf import OracleEngine  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    code = SurfaceCode(3, "primal", "open")
    stub = OracleEngine(code.lattice, 0.09, 0.25, mode="fresh")
    code.engine = lambda *a, **k: stub
    layer = CVLayer(code, delta=0.09, p_swap=0.25)
    errors, st = ec_monte_carlo(1000, code, layer, "UF", {"weight_opts": {"method": "uniform"}}, seed=11, batch=250,
                                mode="fresh", return_stats=True)
    assert st["trials_done"] == 1000 and st["world_size"] == 2 and st["rank"] == rank
    dist.barrier()
    if rank == 0:
        print("DIST_UF_OK", errors)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
