#!/bin/bash
# parity suite incl. the BASELINE-size tests (verbose timing) + timing breakdown of one device-resident / host batch
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s --durations=15 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
grep -E "passed|failed|rc=|borders identical|Error|error" gpurun_out/r2c_pytest.log | tail -20
DYN_TIMING=1 timeout 900 python bench.py --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -c 1500 gpurun_out/r2c_bench.json; grep "dyn timing" gpurun_out/r2c_bench.err | tail -12
