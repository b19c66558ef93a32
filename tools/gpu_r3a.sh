#!/bin/bash
# 16-row groups: -m gpu suite, c2, c5, c2v
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r3a_pytest.log 2>&1; tail -3 gpurun_out/r3a_pytest.log
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r3a_$tag.json 2> gpurun_out/r3a_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3a_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms",0),1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],1), "faults", r.get("ribbon_fault_reads"), "fb", r.get("log2_fallback_reads"), r.get("lin_retry_reads"), r.get("ribbon_fault_reasons"))
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -1 gpurun_out/r3a_$tag.err | cut -c1-250
}
run c2 --no-cpu-baseline --no-e2e
run c5 --config c5 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e
run c2v --config c2v --no-cpu-baseline --no-e2e
python tools/gpu_soak.py 8 > gpurun_out/r3a_soak.log 2>&1; tail -12 gpurun_out/r3a_soak.log
