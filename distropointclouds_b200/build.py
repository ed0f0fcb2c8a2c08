"""Builds libdsc.so (the C-ABI CUDA engine) in-tree with nvcc for sm_100a.

    python -m distropointclouds_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so it travels to
the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdsc.so")
SOURCES = ["dsc_api.cu", "dsc_k1_build.cu", "dsc_k1_f64.cu", "dsc_k23_cells.cu", "dsc_k6_labels.cu", "dsc_k7_voxel.cu"]
HEADERS = ["dsc_internal.h", "dsc_math.cuh", "dsc_k1_common.cuh", "dsc_k1_scan.cuh", "dsc_k1_fallback.cuh", "dsc_k4.cuh", "dsc_math_f64.cuh",
           os.path.join("..", "..", "include", "dsc.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    # host code never evaluates the binning math, but keep contraction off everywhere anyway
    "-Xcompiler", "-ffp-contract=off",
    "-cudart", "static",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine has no CPU fallback and cannot be built without CUDA")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB + ".tmp"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout)
    if verbose:
        print(r.stdout, file=sys.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
