#!/bin/bash
# ncu capture of the long-read ribbon kernel (c4: two-level checkpoints, records-free layout, gather sweep)
mkdir -p gpurun_out
CMD="python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_ribbon -c 1 -f -o gpurun_out/prof_r2f_c4 $CMD > gpurun_out/prof_r2f_c4_ncu.log 2>&1
tail -2 gpurun_out/prof_r2f_c4_ncu.log | cut -c1-300
