#!/bin/bash
# soak on the final build: 24 reads per kind (216 reads) by the default tiers, then every read through the log2-domain ribbon as well
mkdir -p gpurun_out
python tools/gpu_soak.py 24 > gpurun_out/r3l_soak.log 2>&1; tail -11 gpurun_out/r3l_soak.log
DYN_SOAK_OPTS="rib_log=2" python tools/gpu_soak.py 12 > gpurun_out/r3l_soak_log2rib.log 2>&1; tail -11 gpurun_out/r3l_soak_log2rib.log
