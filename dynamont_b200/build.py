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
HDR = [os.path.join(CSRC, f) for f in ("dp_common.cuh", "dp_kernels.cuh", "dp_linear.cuh", "ribbon.h")] + \
      [os.path.join(HERE, "..", "include", "dynamont_b200.h")]
# translation units -> the headers each one depends on (beyond HDR); compiled in parallel, each only when stale
UNITS = {
    "engine.cu": [os.path.join(CSRC, "ntk_kernels.cuh"), os.path.join(CSRC, "ntk_prepass.cuh")],
    "ribbon.cu": [os.path.join(CSRC, "dp_ribbon.cuh")],
}
OBJDIR = os.path.join(CSRC, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # every FMA in the kernels is written explicitly; keeps pass 1 and its recomputation bit-identical
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")
    return p


def _obj(unit: str) -> str:
    return os.path.join(OBJDIR, unit.replace(".cu", ".o"))


def _stale(unit: str) -> bool:
    o = _obj(unit)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(f) > t for f in [os.path.join(CSRC, unit)] + HDR + UNITS[unit])


def up_to_date() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STREAM_BIN):
        return False
    t = os.path.getmtime(LIB)
    return not any(_stale(u) for u in UNITS) and all(os.path.getmtime(_obj(u)) <= t for u in UNITS) and \
        os.path.getmtime(STREAM_SRC) <= os.path.getmtime(STREAM_BIN)


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = nvcc_path()
    todo = [u for u in UNITS if force or _stale(u)]
    procs = []
    for u in todo:
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", "-o", _obj(u), os.path.join(CSRC, u)]
        procs.append((u, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = ""
    failed = False
    for u, cmd, pr in procs:
        out, _ = pr.communicate()
        log += " ".join(cmd) + "\n" + out
        failed = failed or pr.returncode != 0
    if not failed:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [_obj(u) for u in UNITS]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + r.stdout + r.stderr
        failed = r.returncode != 0
    with open(os.path.join(CSRC, "build.log"), "a" if todo != list(UNITS) else "w") as fh:
        fh.write(log)
    if verbose or failed:
        print(log, file=sys.stderr)
    if failed:
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
