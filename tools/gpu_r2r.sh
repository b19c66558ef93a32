#!/bin/bash
# final bench lines of round 2 (copied to profiles/r2_bench_*.json)
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r2r_$tag.json 2> gpurun_out/r2r_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2r_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],3), "e2e", round(d["e2e"]["value"],2) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms",0),1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],2), "faults", r.get("ribbon_fault_reads"), d.get("train"))
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -1 gpurun_out/r2r_$tag.err | cut -c1-200
}
run c3 --config c3 --steps 2 --warmup 1
run c5 --config c5 --steps 3 --warmup 2 --no-cpu-baseline
run c2v --config c2v --no-cpu-baseline --steps 2 --warmup 2
run c4 --config c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
run ref --impl reference
