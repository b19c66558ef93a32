#!/bin/bash
# one ncu --set full capture of the ribbon kernel (mode 1) on c2 x READS reads, after the same command has exited 0
# without ncu.   usage: tools/gpu_prof_rib.sh NAME [READS] [extra bench args]
NAME=$1; READS=${2:-3552}; shift; shift
CMD="python bench.py --config c2 --reads $READS --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $*"
mkdir -p gpurun_out
$CMD > gpurun_out/${NAME}_plain.json 2> gpurun_out/${NAME}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${NAME}_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_ribbon -c 1 -f -o gpurun_out/$NAME $CMD > gpurun_out/${NAME}_ncu.log 2>&1
tail -2 gpurun_out/${NAME}_ncu.log
