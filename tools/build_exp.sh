#!/bin/bash
# experiment build of ONE kernel variant with extra macros:  tools/build_exp.sh NAME VARIANT "-DDYN_X=1 ..."
# -> gpurun_out/exp/libdyn_NAME.so (travels to the GPU box inside gpurun_out? no: gpurun_out/ is not sent) -> build/exp/
NAME=$1; V=$2; shift; shift
mkdir -p build_exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC -shared -Xptxas -v \
  -DDYN_ONLY_VARIANT=$V $* -o build_exp/libdyn_$NAME.so dynamont_b200/csrc/engine.cu > build_exp/$NAME.log 2>&1
grep -A2 "k_alignIN3dyn3Cfg.*ELi1ELi[0-9]*ELb1" build_exp/$NAME.log | grep "registers\|spill" 
