"""Build the CUDA shared library (sm_100a) in-tree: dynamont_b200/csrc/libdynamont_b200.so."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdynamont_b200.so")
SOURCES = [os.path.join(CSRC, "engine.cu")]
HEADERS = [os.path.join(CSRC, "dp_common.cuh"), os.path.join(CSRC, "dp_kernels.cuh"),
           os.path.join(HERE, "..", "include", "dynamont_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # every FMA in the kernels is written explicitly; keeps pass 1 and its recomputation bit-identical
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return p


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(f) <= t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + SOURCES
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + log)
    if verbose or r.returncode:
        print(log, file=sys.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed (see dynamont_b200/csrc/build.log)")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
