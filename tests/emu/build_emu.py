"""TEST INFRASTRUCTURE - builds ``tests/emu/_build/libfpb_emu.so``: the kernel sources of
``flamingpy_b200/csrc`` (fpb.cu and every .cuh it includes) compiled by g++ against the
stand-in ``tests/emu/include/cuda_runtime.h`` and the SIMT machine ``tests/emu/simt.cpp``, so that the parity
tests can run the real kernel code on the CPU (no GPU in the build container).

Two mechanical source transformations on copies of the files (the originals are what nvcc compiles):

* ``callee<<<grid, block, smem, stream>>>(args);`` -> ``fpb_emu::launch(callee, dim3(grid), dim3(block), smem, stream, args);``
* ``extern __shared__ ... T name[];`` -> ``T* name = (T*)fpb_emu::dyn_smem();``

Everything else (inline PTX of the shared-memory accessors and named barriers) is handled by the
``#ifdef FPB_EMU`` branches next to those few helpers in fpb_kernels.cuh.

    python tests/emu/build_emu.py        # builds if out of date, prints the library path
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "flamingpy_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libfpb_emu.so")
SOURCES = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def _match_back(s: str, i: int) -> int:
    """``s[i]`` is ')' : index of the matching '('."""
    depth = 0
    while i >= 0:
        if s[i] == ")":
            depth += 1
        elif s[i] == "(":
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced parentheses before a launch")


def _match_fwd(s: str, i: int, open_c: str = "(", close_c: str = ")") -> int:
    """``s[i]`` is the opening character: index of the matching closing one."""
    depth = 0
    while i < len(s):
        if s[i] == open_c:
            depth += 1
        elif s[i] == close_c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced parentheses after a launch")


def _split_top(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src: str) -> str:
    out = ""
    pos = 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            return out + src[pos:]
        # callee: identifier (with ::), optionally followed by one call's parentheses
        j = k - 1
        while src[j].isspace():
            j -= 1
        if src[j] == ")":
            j = _match_back(src, j) - 1
        while j >= 0 and (src[j].isalnum() or src[j] in "_:"):
            j -= 1
        callee = src[j + 1:k].strip()
        e = src.index(">>>", k)
        cfg = _split_top(src[k + 3:e])
        assert len(cfg) == 4, ("launch configuration must name grid, block, smem, stream", src[k - 40:e + 3])
        a0 = e + 3
        while src[a0].isspace():
            a0 += 1
        assert src[a0] == "("
        a1 = _match_fwd(src, a0)
        args = src[a0 + 1:a1].strip()
        call = f"fpb_emu::launch({callee}, dim3({cfg[0]}), dim3({cfg[1]}), {cfg[2]}, {cfg[3]}"
        call += (", " + args if args else "") + ")"
        out += src[pos:j + 1] + call
        pos = a1 + 1


_DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][A-Za-z0-9_ ]*?)\s+([A-Za-z_][A-Za-z0-9_]*)\s*\[\s*\]\s*;")


def rewrite_dynamic_smem(src: str) -> str:
    return _DYN.sub(lambda m: f"{m.group(1)}* {m.group(2)} = ({m.group(1)}*)fpb_emu::dyn_smem();", src)


def transform(name: str, src_dir: str = CSRC) -> str:
    with open(os.path.join(src_dir, name)) as f:
        src = f.read()
    src = rewrite_dynamic_smem(rewrite_launches(src))
    assert "<<<" not in src and "extern __shared__" not in src, name
    return src.replace('#include "../../include/fpb.h"', '#include "fpb.h"')


def build(force: bool = False) -> str:
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [
        os.path.join(HERE, "include", "cuda_runtime.h"), os.path.join(HERE, "simt.cpp"),
        os.path.join(ROOT, "include", "fpb.h"), os.path.abspath(__file__)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    for name in SOURCES:
        out = name[:-3] + ".cpp" if name.endswith(".cu") else name
        with open(os.path.join(BUILD, out), "w") as f:
            f.write(transform(name))
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing",
           "-Wno-unknown-pragmas", "-Wno-attributes",
           "-I" + os.path.join(HERE, "include"), "-I" + os.path.join(ROOT, "include"), "-I" + BUILD,
           os.path.join(BUILD, "fpb.cpp"), os.path.join(HERE, "simt.cpp"), "-o", LIB]
    subprocess.check_call(cmd, cwd=BUILD)
    return LIB


def build_selftest(force: bool = False) -> str:
    """``_build/libselftest_emu.so``: the kernels of tests/emu/selftest/selftest.cu on the SIMT machine."""
    src_dir = os.path.join(HERE, "selftest")
    lib = os.path.join(BUILD, "libselftest_emu.so")
    deps = [os.path.join(src_dir, "selftest.cu"), os.path.join(HERE, "include", "cuda_runtime.h"),
            os.path.join(HERE, "simt.cpp"), os.path.abspath(__file__)]
    if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, "selftest.cpp"), "w") as f:
        f.write(transform("selftest.cu", src_dir))
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                           "-Wno-attributes", "-I" + os.path.join(HERE, "include"), os.path.join(BUILD, "selftest.cpp"),
                           os.path.join(HERE, "simt.cpp"), "-o", lib], cwd=BUILD)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
