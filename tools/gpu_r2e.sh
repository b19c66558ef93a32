#!/bin/bash
mkdir -p gpurun_out
DYN_TIMING=1 timeout 900 python bench.py --reads 60000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1), d["e2e"])
PY
grep "dyn timing" gpurun_out/r2e_bench.err | tail -24
