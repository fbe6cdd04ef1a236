"""The kernel sources executed on the CPU (no GPU needed): ``tests/emu`` compiles flamingpy_b200/csrc/fpb.cu and its
.cuh files with g++ against a stand-in CUDA runtime and a SIMT machine (one fiber per CUDA thread, blocking warp
collectives and barriers), and this module runs the ``-m gpu`` parity tests against that library in a child pytest
(``FPB_LIB`` points the ctypes binding at it).  Same kernels, same host orchestration, same tests, same oracle - only
the execution engine differs, so a green run here says the kernel *logic* is right; timing, occupancy and real
hardware behaviour are what the GPU box is for.

Test infrastructure only: the product never loads the emulated library (``flamingpy_b200/_lib.py`` opens the in-tree
nvcc build unless ``FPB_LIB`` says otherwise) and has no CPU fallback."""
import os
import re
import subprocess
import sys

import platform
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the SIMT machine switches fibers with a few lines of x86-64 System V assembly (tests/emu/simt.cpp)
pytestmark = pytest.mark.skipif(platform.machine() not in ("x86_64", "AMD64") or shutil.which("g++") is None,
                                reason="the kernel emulation needs g++ on x86-64")

# not runnable on the emulated library: the pybind11 module links the nvcc build; torch views need a CUDA context
UNSUPPORTED = ["fpbpy", "device_tensors", "threshold_sweep"]
# correct under emulation but too slow for the CPU suite (tens of seconds to minutes each at 1e3-1e4 x GPU time)
SLOW = ["d17_to_d25", "d21", "failure_rate", "same_failures_with_either_handoff", "fetch_decode_matches",
        "macro_monte_carlo", "two_host_threads", "monte_carlo_d9", "uf_monte_carlo_is_batch", "native_matcher_end_to_end",
        "run_bits_full_size", "engine_reuse", "candidate_lists_for_the_matcher", "compat_int_lengths and 11-primal",
        "full_fetch_optimum and both-periodic", "full_fetch_optimum and 13-primal"]


@pytest.fixture(scope="module")
def emulated_library():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    try:
        import build_emu
    finally:
        sys.path.pop(0)
    return build_emu.build()


def _child_pytest(lib, args, timeout, **extra_env):
    env = dict(os.environ, FPB_LIB=lib, FPB_BINDING="ctypes", **extra_env)
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "--timeout=180"] + args
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    m = re.search(r"(\d+) passed", r.stdout)
    return r, int(m.group(1)) if m else 0


def test_simt_machine_semantics():
    """The emulator itself: shuffles, votes, scans, __syncthreads with exited threads, atomics, named barriers with a
    predicate OR, shared-window addresses and saturating conversions on small kernels with known answers
    (tests/emu/selftest/selftest.cu), also under a shuffled schedule."""
    import ctypes

    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    try:
        import build_emu
    finally:
        sys.path.pop(0)
    lib = ctypes.CDLL(build_emu.build_selftest())
    out = np.zeros(1024, np.int32)
    assert lib.selftest_run(out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0
    lane = np.arange(32)
    v = lane * 3 + 1
    assert np.array_equal(out[0:32], np.full(32, 16))
    assert np.array_equal(out[32:64], np.where(lane >= 2, v[np.maximum(lane - 2, 0)], v))
    assert np.array_equal(out[64:96], v[lane ^ 16])
    ballot = sum(1 << i for i in range(32) if i % 3 == 0)
    assert np.array_equal(out[96:128].view(np.uint32), np.full(32, ballot, np.uint32))
    assert np.array_equal(out[128:160], np.full(32, 1 + 2 * 0xff))
    assert np.array_equal(out[160:192], lane + 1)
    tid = np.arange(128)
    assert np.array_equal(out[192:320], 8128 + (127 - tid))
    live = tid[40:]
    assert np.array_equal(out[320 + 40:448], 7348 + (127 - (live - 40) % 88)) and not out[320:360].any()
    team, r = tid >> 6, tid & 63
    assert np.array_equal(out[448:576], ((team << 6) + (63 - r)) * 4 + (team == 1) * 2)
    assert out[576:583].tolist() == [2**31 - 1, -2**31, -3, -2, 0, 0xc0000000 - 2**32, 5 + 800]
    # the same answers when warps and lanes are visited in a shuffled order
    code = ("import ctypes, numpy as np, sys; lib = ctypes.CDLL(sys.argv[1]); o = np.zeros(1024, np.int32); "
            "lib.selftest_run(o.ctypes.data_as(ctypes.POINTER(ctypes.c_int))); sys.stdout.write(o.tobytes().hex())")
    r2 = subprocess.run([sys.executable, "-c", code, build_emu.build_selftest()], capture_output=True, text=True,
                        env=dict(os.environ, FPB_EMU_SHUFFLE="7"), timeout=120)
    assert r2.returncode == 0 and bytes.fromhex(r2.stdout) == out.tobytes(), r2.stderr[-500:]


def test_emulated_library_exports_the_whole_abi(emulated_library):
    """Same translation unit as libfpb.so, so the emulated build must export every entry point of include/fpb.h."""
    import ctypes

    from flamingpy_b200 import _lib

    lib = ctypes.CDLL(emulated_library)
    for name in _lib.EXPORTS:
        assert hasattr(lib, name), name
    assert lib.fpb_abi_version() == _lib.FPB_ABI_VERSION


def test_gpu_parity_tests_pass_on_the_emulated_kernels(emulated_library):
    """Golden vectors, seeded injected noise vs the oracle (d = 5..9, both complexes, integer weights), the device-RNG
    pipeline replayed through the oracle, bucket-width independence, d = 13 full-size properties, correction paths."""
    r, passed = _child_pytest(emulated_library, ["tests/test_gpu_parity.py", "-n", "4"], 900)
    assert r.returncode == 0 and passed >= 30, r.stdout[-3000:] + r.stderr[-1000:]


def test_remaining_gpu_tests_pass_on_the_emulated_kernels(emulated_library):
    """Edge cases (ragged batches, d = 13 team kernels against the C oracle, static-weight tables, rustworkx-compat
    lengths), the facade (reference runs reproduced trial by trial), the sparse hand-off kernels, the macronode
    kernels, the i.i.d. arm and the Union-Find path - minus the cases listed in UNSUPPORTED / SLOW."""
    deselect = " and ".join(f"not ({k})" for k in UNSUPPORTED + SLOW)
    files = ["tests/test_gpu_edge_cases.py", "tests/test_gpu_facade.py", "tests/test_gpu_handoff.py",
             "tests/test_gpu_uf.py", "tests/test_zz_gpu_uf_device.py", "tests/test_gpu_boundary.py", "tests/test_macro.py",
             "tests/test_iid.py"]
    r, passed = _child_pytest(emulated_library, files + ["-n", "4", "-k", deselect], 1500)
    assert r.returncode == 0 and passed >= 95, r.stdout[-3000:] + r.stderr[-1000:]


def test_results_do_not_depend_on_the_interleaving_of_warps(emulated_library):
    """``FPB_EMU_SHUFFLE``: warps and lanes are visited in a pseudo-random order that changes at every scheduling
    pass.  The searches (one-warp and two-warp team kernels, store-then-verify protocol) and the device Union-Find
    decoder (racing min-id hooks) must give the same results under any interleaving."""
    r, passed = _child_pytest(emulated_library, ["tests/test_gpu_parity.py", "tests/test_zz_gpu_uf_device.py", "-n", "4"], 900,
                              FPB_EMU_SHUFFLE="20261020")
    assert r.returncode == 0 and passed >= 45, r.stdout[-3000:] + r.stderr[-1000:]


VARIANTS = [  # environment switch -> (what it selects in csrc/fpb.cu, test that exercises it)
    ({"FPB_TEAM_W": "4"}, "k_paths_teamw<.., 4, ..>: four warps per search (the d >= 17 kernels) on the d = 13 lattice",
     "tests/test_gpu_edge_cases.py", "both_complexes_full_size"),
    ({"FPB_FORCE_WG": "1"}, "weights and face tables in global memory (the large-lattice variants)",
     "tests/test_gpu_edge_cases.py", "both_complexes_full_size"),
    ({"FPB_NO_TEAM": "1"}, "one warp per search at d = 13 (k_paths<5, 2, 384, ..>, two cubes per lane)",
     "tests/test_gpu_edge_cases.py", "both_complexes_full_size"),
    ({"FPB_TEAM_RM": "1"}, "team_search_rm: residue-major bucket rows, lean verify (the opt-in round-2 experiment)",
     "tests/test_gpu_edge_cases.py", "both_complexes_full_size"),
    ({"FPB_GATHER_ROWS": "1"}, "k_gather_rows: the row-driven static-weight gather (opt-in)",
     "tests/test_gpu_edge_cases.py", "static_weight_table_is_bit_identical"),
]


def test_every_search_kernel_variant_against_the_c_oracle(emulated_library):
    """The kernel variants that the default configuration does not pick at test sizes, forced through their environment
    switches: two fresh d = 13 periodic ``ec="both"`` trials against the C oracle (1e-9), and the static-weight gather
    against the search.  (FPB_TEAM_W=8 / 16 pass too but take 35 s / 2 min under emulation: run them by hand.)"""
    from concurrent.futures import ThreadPoolExecutor

    def run(v):
        env, _, path, key = v
        return _child_pytest(emulated_library, [path, "-k", key], 900, **env)

    with ThreadPoolExecutor(len(VARIANTS)) as ex:
        results = list(ex.map(run, VARIANTS))
    for v, (r, passed) in zip(VARIANTS, results):
        assert r.returncode == 0 and passed >= 1, (v[0], r.stdout[-2000:] + r.stderr[-500:])
