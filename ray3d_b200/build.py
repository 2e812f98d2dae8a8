"""Build libray3d_b200.so in-tree with nvcc for sm_100a (no torch dependency, no JIT cache).

    python -m ray3d_b200.build            # incremental
    python -m ray3d_b200.build --force

The shared library lives next to this file so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libray3d_b200.so")
SOURCES = ["r3d_plan.cpp", "r3d_stage_kernels.cu", "r3d_gemm_ffma.cu", "r3d_gemm_tc.cu", "r3d_tail_tc.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: ray3d_b200 needs the CUDA 12.9 toolchain to build its sm_100a library")
    return cand


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = nvcc_path()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "ray3d_b200.h"))
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            # experiment builds: R3D_BUILD_EXPERIMENTS=1 compiles the timing-experiment switches in (env R3D_TC_DEBUG etc.,
            # some give wrong results by design); R3D_BUILD_TRACE=1 adds the per-tile clock stamps (scripts/tile_trace.py)
            trace = ["-DR3D_TC_TRACE", "-DR3D_EXPERIMENTS"] if os.environ.get("R3D_BUILD_TRACE") == "1" else []
            if os.environ.get("R3D_BUILD_EXPERIMENTS") == "1" and not trace:
                trace = ["-DR3D_EXPERIMENTS"]
            cmd = [nvcc, *ARCH, *trace, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                   "-Xptxas", "-v" if verbose else "-warn-spills", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                   "-x", "cu", "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static", "-ldl"]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
