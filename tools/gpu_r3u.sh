#!/bin/bash
# last A/B of the session: resident CTAs per SM and batch size with the final 16-row kernel
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 900 python bench.py --no-cpu-baseline --no-e2e "$@" > gpurun_out/r3u_$tag.json 2> gpurun_out/r3u_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3u_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],1), "kernel_ms", round(r.get("kernel_ms") or 0,1), "ms/step", round(d["ms_per_step"],1))
except Exception as e:
    print("$tag FAILED", e)
PY
}
run bps5
run bps6 --opt rib_bps=6
run b33k --batch 33334
run b50k --batch 50000
