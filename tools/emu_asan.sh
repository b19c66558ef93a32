#!/bin/bash
# The kernel sources compiled for the CPU emulator (tests/emu) with AddressSanitizer + UBSan, then the emulator test-suite.
# The emulator runs every lane as a ucontext fibre, so ASan's stack-use-after-return detection is switched off.
set -e
cd "$(dirname "$0")/.."
OUT=tests/emu/_build/libdynamont_emu_asan.so
g++ -O1 -g -std=c++17 -fPIC -shared -ffp-contract=off -fsanitize=address,undefined -fno-omit-frame-pointer -DDYN_HOST_EMU=1 \
    -include tests/emu/simt_host.h -x c++ dynamont_b200/csrc/engine.cu -o $OUT
LIBASAN=$(g++ -print-file-name=libasan.so)
ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:verify_asan_link_order=0 UBSAN_OPTIONS=print_stacktrace=1 \
  LD_PRELOAD=$LIBASAN DYN_EMU_LIB=$PWD/$OUT python -m pytest tests/test_emu_kernels.py -x -q "$@"
