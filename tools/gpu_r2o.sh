#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_ntk_stages.py -m gpu -x -q -s --durations=8 > gpurun_out/r2o_ntk.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_ntk.log
tail -25 gpurun_out/r2o_ntk.log
python tools/ntk_timing.py > gpurun_out/r2o_ntk_timing.log 2>&1; tail -5 gpurun_out/r2o_ntk_timing.log
