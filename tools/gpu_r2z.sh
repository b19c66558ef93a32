#!/bin/bash
# log2-domain ribbon on the GPU: full -m gpu suite, c4 (4 steps), c2 control
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_pytest.log 2>&1; tail -3 gpurun_out/r2z_pytest.log
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r2z_$tag.json 2> gpurun_out/r2z_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2z_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms",0),1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],1), "faults", r.get("ribbon_fault_reads"), "fb", r.get("log2_fallback_reads"), r.get("lin_retry_reads"), r.get("ribbon_fault_reasons"))
except Exception as e:
    print("$tag FAILED", e)
PY
  grep "dyn timing" gpurun_out/r2z_$tag.err | tail -4 | cut -c1-250
  tail -1 gpurun_out/r2z_$tag.err | cut -c1-250
}
DYN_TIMING=1 run c4 --config c4 --steps 4 --warmup 1 --no-cpu-baseline --no-e2e
unset DYN_TIMING
run c2 --no-cpu-baseline --no-e2e
