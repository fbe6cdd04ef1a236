"""Host-side logic that needs no GPU: the C ABI library loads and exports every
symbol the header declares, the reference-compatible facade objects, host
samplers, matching helpers and the trial partition."""

import ctypes
import os
import re

import numpy as np
import pytest

from flamingpy_b200 import _lib
from flamingpy_b200.codes import SurfaceCode, alternating_polarity
from flamingpy_b200.matching import dense_inverse_weights, solve_matching, tri_index
from flamingpy_b200.noise import LabelSampler, draw, sample_batch, sigma_vector
from flamingpy_b200.simulations import trial_range
from oracle import restate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """No compute calls: just that libfpb.so loads and has the header's entry points."""
    header = open(os.path.join(ROOT, "include", "fpb.h")).read()
    declared = set(re.findall(r"\b(fpb_[a-z_]+)\s*\(", header))
    assert {"fpb_create", "fpb_run_injected", "fpb_run_rng", "fpb_fetch", "fpb_paths_toggle"} <= declared
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert set(_lib.EXPORTS) == declared
    assert _lib.load().fpb_abi_version() == _lib.FPB_ABI_VERSION


def test_matcher_library_exports_every_declared_symbol():
    from flamingpy_b200 import mwpm_native

    header = open(os.path.join(ROOT, "include", "fpb_mwpm.h")).read()
    declared = set(re.findall(r"\b(fpb_mwpm[a-z_]*)\s*\(", header))
    assert {"fpb_mwpm", "fpb_mwpm_batch", "fpb_mwpm_batch_cand", "fpb_mwpm_configure", "fpb_mwpm_stats",
            "fpb_mwpm_max_threads", "fpb_mwpm_sparse_begin", "fpb_mwpm_sparse_requests", "fpb_mwpm_sparse_start",
            "fpb_mwpm_sparse_solve", "fpb_mwpm_sparse_apply", "fpb_mwpm_sparse_finish",
            "fpb_mwpm_sparse_free"} == declared
    lib = ctypes.CDLL(mwpm_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_create_fails_loudly_without_gpu():
    """The product path has no CPU fallback: without a device fpb_create errors out."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from flamingpy_b200.engine import TrialEngine
    from flamingpy_b200.lattice import build_lattice

    with pytest.raises(_lib.FpbError):
        TrialEngine(build_lattice(3, "primal", "open"), 0.09, 0.25)


def test_pybind_module_is_the_default_binding(monkeypatch):
    """flamingpy_b200.cpp.fpbpy (the reference's native-plugin pattern, lemonpy.cpp:50-84) imports,
    carries the header's ABI version and every Handle method the engine calls, and is what
    ``open_handle`` uses; without a device its constructor fails like the ctypes path."""
    import torch

    fpbpy = _lib.pybind_module()
    assert fpbpy is not None and fpbpy.abi_version() == _lib.FPB_ABI_VERSION
    assert _lib.binding_name() == "pybind11"
    monkeypatch.setenv("FPB_BINDING", "ctypes")
    assert _lib.binding_name() == "ctypes"
    monkeypatch.delenv("FPB_BINDING")
    methods = {n for n in dir(_lib.CtypesHandle) if not n.startswith("_")}
    assert methods <= set(dir(fpbpy.Handle)), methods - set(dir(fpbpy.Handle))
    if not torch.cuda.is_available():
        from flamingpy_b200.engine import config_dict
        from flamingpy_b200.lattice import build_lattice

        cfg, _, _ = config_dict(build_lattice(3, "primal", "open"), 0.09, 0.25, {"method": "uniform"})
        with pytest.raises(_lib.FpbError, match="no usable CUDA device"):
            fpbpy.Handle(cfg)
        with pytest.raises(_lib.FpbError, match="no usable CUDA device"):
            _lib.CtypesHandle(cfg)
        with pytest.raises((RuntimeError, KeyError)):
            fpbpy.Handle({"n_nodes": 3})


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "flamingpy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/ is", "").replace("``oracle/``", ""), f


@pytest.mark.parametrize("d,ec,b", [(3, "primal", "open"), (3, "dual", "open"), (4, "primal", "toric"),
                                    (3, "both", "periodic"), ((2, 3, 4), "primal", "open")])
def test_surface_code_facade_sizes(d, ec, b):
    code = SurfaceCode(d, ec, b)
    dims = code.dims
    assert code.ec == (["primal", "dual"] if ec == "both" else [ec])
    assert len(code.graph) == code.lattice.N
    for e in code.ec:
        ct = code.lattice.complexes[e]
        assert len(getattr(code, e + "_stabilizers")) == ct.S
        assert sorted(getattr(code, e + "_syndrome_inds")) == ct.qubits.tolist()
        sg = getattr(code, e + "_stab_graph")
        assert len(sg.low_bound_points) == len(sg.high_bound_points)
        assert bool(sg.has_bound_points()) == (b == "open")
    p = code.all_syndrome_coords[0]
    assert code.graph.to_points[code.graph.to_indices[p]] == p
    assert code.graph.nodes[p]["type"] in ("primal", "dual")
    assert all(np.count_nonzero(np.array(q) - np.array(p)) == 1 for q in code.graph[p])  # lattice neighbours differ along one axis
    if np.ndim(d) == 0 and b == "periodic":
        assert len(code.graph) == 6 * dims[0] ** 3


def test_surface_code_rejects_bad_arguments():
    with pytest.raises(ValueError):
        SurfaceCode(3, ec="primal", boundaries="closed")
    with pytest.raises(ValueError):
        SurfaceCode(3, ec="both", boundaries="open")
    SurfaceCode(3, polarity=alternating_polarity)


@pytest.mark.reference
def test_facade_matches_reference_attributes(reference):
    from oracle import ref_shim

    ref = ref_shim.make_code(3, "primal", "open")
    code = SurfaceCode(3, "primal", "open")
    assert code.bound_str == ref.bound_str
    assert list(code.boundaries) == list(ref.boundaries)
    assert code.graph.to_indices == ref.graph.to_indices
    assert set(code.primal_syndrome_coords) == set(ref.primal_syndrome_coords)
    assert set(code.primal_bound_points) == set(ref.primal_bound_points)
    for s_ref, s in zip(ref.primal_stabilizers, code.primal_stabilizers):
        assert set(s_ref.coords()) == set(s.coords())
    assert set(code.graph.slice_coords("x", 0)) == set(ref.graph.slice_coords("x", 0))
    for p in list(ref.graph.nodes)[:20]:
        assert set(code.graph[p]) == set(ref.graph[p])


def test_host_sampler_matches_oracle_chain():
    """Product sampler and oracle restatement are separate code; same seed, same draws."""
    N, delta, ps = 95, 0.09, 0.25
    a, b = np.random.default_rng(3), np.random.default_rng(3)
    mine = LabelSampler(N, ps)
    orc = restate.LabelChain(N, ps)
    for _ in range(6):
        st = mine.next(a)
        so = orc.next(b)
        assert np.array_equal(st, so)
        sig = sigma_vector(st, delta)
        assert np.array_equal(sig, restate.sigma_from_state(so, delta))
        assert np.array_equal(draw(a, sig), restate.draw_quadratures(b, sig))


def test_init_covs_exact_float32():
    """reference tests/noise_models/test_cv_noise.py:148-165"""
    N, delta = 50, 0.07
    all_gkp = sigma_vector(np.ones(N, np.uint8), delta)
    assert all_gkp.dtype == np.float32
    assert np.array_equal(all_gkp, np.full(2 * N, (delta / 2) ** 0.5, dtype=np.float32))
    all_p = sigma_vector(np.full(N, 2, np.uint8), delta)
    expect = np.concatenate([np.full(N, 1 / (2 * delta) ** 0.5, np.float32), np.full(N, (delta / 2) ** 0.5, np.float32)])
    assert np.array_equal(all_p, expect)


def test_sample_batch_compat_and_fresh():
    x, st = sample_batch(np.random.default_rng(0), 40, 0.09, 0, 4)
    assert [int(s.any()) for s in st] == [1, 0, 1, 0]
    assert not x[1].any() and x[0].all()
    x, st = sample_batch(np.random.default_rng(0), 40, 0.09, 0, 4, fresh=True)
    assert all(s.all() for s in st)
    fr = np.mean([LabelSampler(549, 0.3, fresh=True).next(np.random.default_rng(i)) == 2 for i in range(200)])
    assert abs(fr - 0.3) < 0.02


def test_tri_index_and_dense_matrix():
    K = 5
    pair = np.arange(10, dtype=float) + 1
    k = 0
    for a in range(K - 1):
        for b in range(a + 1, K):
            assert tri_index(a, b, K) == k
            k += 1
    A = dense_inverse_weights(K, pair, np.full(K, 0.5))
    assert A.shape == (10, 10) and np.allclose(A, A.T)
    assert A[1, 3] == -pair[tri_index(1, 3, K)] and A[2, 7] == -0.5 and A[0, 7] < -1e18 and A[6, 8] == 0
    assert dense_inverse_weights(4, np.ones(6), None).shape == (4, 4)


def test_solve_matching_small():
    # two close pairs far from each other
    K = 4
    pair = np.array([1.0, 9.0, 9.0, 9.0, 9.0, 1.0])  # (0,1)=1 (2,3)=1
    assert sorted(solve_matching(K, pair, None)) == [(0, 1), (2, 3)]
    # with cheap boundary legs every cube goes to its connector
    assert sorted(solve_matching(K, pair, np.full(K, 0.1))) == [(0, -1), (1, -1), (2, -1), (3, -1)]
    assert solve_matching(0, np.empty(0), None) == []


def test_trial_range_partitions_exactly():
    for trials in (0, 1, 7, 100, 1001):
        for size in (1, 2, 3, 8):
            spans = [trial_range(trials, r, size) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == trials
            assert all(spans[i][1] == spans[i + 1][0] for i in range(size - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bench_uf_leg_never_raises(monkeypatch):
    """bench.run_uf_leg collects the JSON lines of scripts/uf_leg.py; a crashed, silent or hung child ends up as an
    entry of the returned object, never as an exception in the bench line's path."""
    import subprocess
    import types

    import bench

    args = types.SimpleNamespace(mode="compat", uf_trials=64)
    cfg = dict(distance=3, ec="primal", boundaries="open", delta=0.09, p_swap=0.0)

    def fake(stdout="", returncode=0, stderr="", exc=None):
        def run(cmd, **kw):
            assert cmd[1].endswith("uf_leg.py") and "--trials" in cmd and kw["timeout"] > 0
            if exc:
                raise exc
            return types.SimpleNamespace(stdout=stdout, stderr=stderr, returncode=returncode)
        return run

    monkeypatch.setattr(subprocess, "run", fake('noise\n{"leg": "host", "value": 5.0}\n{"leg": "device", "error": "x"}\n'))
    assert bench.run_uf_leg(args, cfg) == {"host": {"value": 5.0}, "device": {"error": "x"}}
    monkeypatch.setattr(subprocess, "run", fake('{"leg": "host", "value": 5.0}\n', returncode=-11, stderr="Segmentation fault"))
    out = bench.run_uf_leg(args, cfg)
    assert out["host"] == {"value": 5.0} and out["exit"]["returncode"] == -11
    monkeypatch.setattr(subprocess, "run", fake(exc=subprocess.TimeoutExpired("uf_leg", 300)))
    assert "TimeoutExpired" in bench.run_uf_leg(args, cfg)["error"]
