#!/bin/bash
# training soak (M-step metric on occupied kmers) + ncu captures of HEAD (align, train) for the final counters
mkdir -p gpurun_out
python tools/gpu_train_soak.py 24 > gpurun_out/r3p_train_soak.log 2>&1; tail -7 gpurun_out/r3p_train_soak.log
CMD="python bench.py --config c2 --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/prof_r2f_align_plain.json 2> gpurun_out/prof_r2f_align_plain.err
ncu --set full --clock-control none --import-source on -k regex:k_ribbon -c 1 -f -o gpurun_out/prof_r2f_align $CMD > gpurun_out/prof_r2f_align_ncu.log 2>&1
tail -1 gpurun_out/prof_r2f_align_ncu.log
CMD="python bench.py --config c5 --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/prof_r2f_train_plain.json 2> gpurun_out/prof_r2f_train_plain.err
ncu --set full --clock-control none --import-source on -k regex:k_ribbon -c 1 -f -o gpurun_out/prof_r2f_train $CMD > gpurun_out/prof_r2f_train_ncu.log 2>&1
tail -1 gpurun_out/prof_r2f_train_ncu.log
CMD="python bench.py --config c2 --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2f.csv $CMD > gpurun_out/launches_r2f.log 2>&1
tail -1 gpurun_out/launches_r2f.log | cut -c1-200
