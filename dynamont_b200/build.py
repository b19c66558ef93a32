"""Build the CUDA shared library (sm_100a) in-tree: dynamont_b200/csrc/libdynamont_b200.so."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdynamont_b200.so")
STREAM_BIN = os.path.join(CSRC, "dynamont-NT-b200")
STREAM_SRC = os.path.join(CSRC, "stream_main.cpp")
SOURCES = [os.path.join(CSRC, "engine.cu")]
HEADERS = [os.path.join(CSRC, "dp_common.cuh"), os.path.join(CSRC, "dp_kernels.cuh"), os.path.join(CSRC, "dp_linear.cuh"),
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
    return os.path.exists(STREAM_BIN) and all(os.path.getmtime(f) <= t for f in SOURCES + HEADERS + [STREAM_SRC])


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
    build_stream()
    return LIB


def build_stream() -> str:
    """The streaming stdin/stdout front end: plain C++ host code linked against the C ABI library."""
    cmd = ["g++", "-O2", "-std=c++17", "-o", STREAM_BIN, STREAM_SRC, "-L" + CSRC, "-ldynamont_b200",
           "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return STREAM_BIN


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
