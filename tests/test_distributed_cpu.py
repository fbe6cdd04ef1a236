"""The N>1 path on CPU: two gloo ranks run the trial partition and the one collective."""

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gloo_partition_and_allreduce():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "DIST_OK 143 1001" in out.stdout


def test_two_rank_gloo_monte_carlo_driver_with_uf():
    """ec_monte_carlo itself under two gloo ranks (decoder="UF": the host decoder runs anywhere; the device
    calls are answered by the C-oracle stand-in): the all-reduced failure count equals the single-process one."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "dist_worker_uf.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    two = int(out.stdout.split("DIST_UF_OK")[1].split()[0])

    sys.path.insert(0, ROOT)
    from flamingpy_b200.codes import SurfaceCode
    from flamingpy_b200.noise import CVLayer
    from flamingpy_b200.simulations import ec_monte_carlo
    from tests.test_uf import OracleEngine

    code = SurfaceCode(3, "primal", "open")
    stub = OracleEngine(code.lattice, 0.09, 0.25, mode="fresh")
    code.engine = lambda *a, **k: stub
    one = ec_monte_carlo(1000, code, CVLayer(code, delta=0.09, p_swap=0.25), "UF", {"weight_opts": {"method": "uniform"}},
                         seed=11, batch=250, mode="fresh")
    assert one == two and 150 < one < 400
