"""TEST INFRASTRUCTURE — compile the CUDA sources against the SIMT emulator (simt_host.h) with g++ so the
kernels can be unit-tested without a GPU.  Output: tests/emu/_build/libdynamont_emu.so (never loaded by the
product; dynamont_b200 only ever loads csrc/libdynamont_b200.so)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
CSRC = os.path.join(ROOT, "dynamont_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libdynamont_emu.so")
DEPS = [os.path.join(CSRC, f) for f in ("engine.cu", "dp_common.cuh", "dp_kernels.cuh", "dp_linear.cuh", "dp_ribbon.cuh", "ribbon.cu", "ribbon.h", "ntk_kernels.cuh")] + \
       [os.path.join(HERE, "simt_host.h"), os.path.join(ROOT, "include", "dynamont_b200.h")]


def build(force: bool = False) -> str:
    if os.environ.get("DYN_EMU_LIB"):  # tools/emu_asan.sh: a sanitizer build of the same sources
        return os.environ["DYN_EMU_LIB"]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DDYN_HOST_EMU=1"] + (["-DDYN_RIB_DEBUG=1"] if os.environ.get("DYN_RIB_DEBUG") else []) + [
           "-include", os.path.join(HERE, "simt_host.h"), "-x", "c++", os.path.join(CSRC, "engine.cu"),
           "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv))
