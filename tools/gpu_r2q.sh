#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -3 gpurun_out/r2q_pytest.log
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r2q_$tag.json 2> gpurun_out/r2q_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2q_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],2), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms",0),1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],1), "faults", r.get("ribbon_fault_reads"), d.get("cpu_baseline",{}) and d["cpu_baseline"].get("value"))
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -1 gpurun_out/r2q_$tag.err | cut -c1-200
}
run c2
run c3 --config c3 --steps 2 --warmup 1
