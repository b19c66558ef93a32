#!/bin/bash
# async lanes: GPU tests + the default bench (c2, 100 000 reads per step)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2d_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1), "ribbon", r.get("ribbon_reads"), r.get("ribbon_fault_reads"))
PY
tail -3 gpurun_out/r2d_bench.err
