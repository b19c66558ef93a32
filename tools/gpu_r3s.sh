#!/bin/bash
mkdir -p gpurun_out
export DYN_TIMING=1
timeout 1500 python bench.py --config c4 --steps 5 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r3s_c4.json 2> gpurun_out/r3s_c4.err
unset DYN_TIMING
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3s_c4.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("c4 value", round(d["value"],1), "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "faults", r["ribbon_fault_reads"], "fb", r["log2_fallback_reads"], "ok", d["config"]["reads_ok"])
PY
grep "scratch:" gpurun_out/r3s_c4.err | tail -6 | cut -c1-200
grep "ribbon_kernel" gpurun_out/r3s_c4.err | tail -4 | cut -c1-250
