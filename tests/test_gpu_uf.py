"""The Union-Find decoder behind the reference's API on the GPU box: the device pipeline delivers
bits, syndrome and state labels (stage FPB_STAGE_SYNDROME), the host decoder (csrc/uf.cpp) grows and
peels.  Mirrors tests/decoders/test_uf_decoder.py:426-520 of the reference (weights, erasures,
"recovery leaves no unsatisfied stabilizer", end-to-end correct) on the reference's own vectors."""

import json
import os

import numpy as np
import pytest

from flamingpy_b200.codes import SurfaceCode
from flamingpy_b200.decoder import assign_weights, correct
from flamingpy_b200.noise import CVLayer, CVMacroLayer, IidNoise
from flamingpy_b200.simulations import ec_monte_carlo
from tests import golden_util as gu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", gu.uf_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_correct_uf_reproduces_reference_run(path):
    from tests.test_uf import replay_uf_golden

    g = gu.load_uf(path)
    code = SurfaceCode(g["distance"], g["ec"], g["boundaries"])
    replay_uf_golden(g, code, CVLayer(code, delta=g["delta"], p_swap=g["p_swap"]))


def test_incompatible_weights_raise_like_the_reference():
    code = SurfaceCode(3, "primal", "open")
    CVLayer(code, delta=0.09, p_swap=0.25).apply_noise(np.random.default_rng(0))
    with pytest.raises(Exception, match="Incompatible decoder"):
        correct(code, "UF", {"method": "blueprint", "delta": 0.09})
    with pytest.raises(Exception, match="Incompatible decoder"):
        assign_weights(code, "UF", method="blueprint")
    with pytest.raises(Exception, match="Incompatible decoder"):
        ec_monte_carlo(10, code, CVLayer(code, delta=0.09, p_swap=0.25), "UF",
                       {"weight_opts": {"method": "blueprint", "delta": 0.09}})


def _uf_cases():
    path = os.path.join(gu.GOLDEN_DIR, "failure_rates_uf_reference.json")
    if not os.path.exists(path):
        return []
    with open(path) as f:
        return [(r["distance"], r["boundaries"], r["delta"], r["p_swap"], r["failures"], r["trials"])
                for r in json.load(f) if r["boundaries"] == "open"]


@pytest.mark.parametrize("d,b,delta,ps,ref_fail,ref_n", _uf_cases())
def test_uf_failure_rate_consistent_with_live_reference(d, b, delta, ps, ref_fail, ref_n):
    """ec_monte_carlo(decoder="UF") with device RNG against the rate of the live reference's UF decoder
    (oracle/gen_failure_rates_uf_ref.py).  Two-sample binomial z test at 3.29 sigma (p = 1e-3)."""
    code = SurfaceCode(d, "primal", b)
    layer = CVLayer(code, delta=delta, p_swap=ps)
    n = 40000 if d == 3 else 8000
    errors = ec_monte_carlo(n, code, layer, "UF", {"weight_opts": {"method": "uniform"}}, seed=20261020, mode="compat",
                            batch=4000)
    p1, p2 = errors / n, ref_fail / ref_n
    pooled = (errors + ref_fail) / (n + ref_n)
    z = (p1 - p2) / np.sqrt(pooled * (1 - pooled) * (1 / n + 1 / ref_n))
    print(f"UF d={d} ps={ps}: gpu {errors}/{n}={p1:.5f} live reference {ref_fail}/{ref_n}={p2:.5f} z={z:+.2f}")
    assert abs(z) < 3.29, (errors, n, p1, p2, z)


def test_uf_monte_carlo_is_batch_invariant_and_runs_every_noise_model():
    code = SurfaceCode(5, "primal", "open")
    layer = CVLayer(code, delta=0.09, p_swap=0.2)
    args = {"weight_opts": {"method": "uniform"}}
    a = ec_monte_carlo(600, code, layer, "UF", args, seed=5, batch=64, mode="fresh")
    b, st = ec_monte_carlo(600, code, layer, "UF", args, seed=5, batch=600, mode="fresh", return_stats=True)
    assert a == b and st["trials_done"] == 600 and 0 < a < 600
    # fewer failures than no decoding at all would leave: the MWPM decoder on the same trials does not lose to it by much
    m = ec_monte_carlo(600, code, layer, "MWPM", args, seed=5, batch=600, mode="fresh")
    assert m <= a + 40
    periodic = SurfaceCode(4, "primal", "periodic")
    e = ec_monte_carlo(400, periodic, IidNoise(periodic, 0.03), "UF", args, seed=6, batch=128)
    assert 0 <= e < 200
    macro = ec_monte_carlo(300, code, CVMacroLayer(code, delta=0.08, p_swap=0.1), "UF", args, seed=7, batch=100)
    assert 0 <= macro < 300
    assert ec_monte_carlo(200, code, CVLayer(code, delta=0.001, p_swap=0), "UF", args, seed=8) == 0
