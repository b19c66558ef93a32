#!/bin/bash
# long noisy reads against the oracle (16-row groups under noise), training soak with the M-step metric
mkdir -p gpurun_out
DYN_SOAK_KINDS="noisy_long noisy2_long" timeout 1200 python tools/gpu_soak.py 10 > gpurun_out/r3o_soak_long.log 2>&1; tail -4 gpurun_out/r3o_soak_long.log
python tools/gpu_train_soak.py 24 > gpurun_out/r3o_train_soak.log 2>&1; tail -7 gpurun_out/r3o_train_soak.log
