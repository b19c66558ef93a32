#!/bin/bash
# one ncu --set full capture of k_align (mode 1) for a build variant, on c1 x 1184 reads (one read per resident warp at
# 8 CTAs/SM), after the same command has exited 0 without ncu.   usage: tools/gpu_prof.sh NAME VARIANT [extra bench args]
NAME=$1; V=$2; shift; shift
CMD="python bench.py --config c1 --reads 1184 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --variant $V $*"
$CMD > gpurun_out/${NAME}_plain.json 2> gpurun_out/${NAME}_plain.err || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_align -c 1 -f -o gpurun_out/$NAME $CMD > gpurun_out/${NAME}_ncu.log 2>&1
tail -2 gpurun_out/${NAME}_ncu.log
