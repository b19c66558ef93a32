#!/bin/bash
# host-share experiment: the default bench line with the process confined to 4 host cores (the share one rank has on a
# 32-core box at 8 GPUs), next to the unconfined run
mkdir -p gpurun_out
nproc > gpurun_out/r2s_nproc.txt; free -g >> gpurun_out/r2s_nproc.txt
run() { tag=$1; shift
  timeout 900 "$@" > gpurun_out/r2s_$tag.json 2> gpurun_out/r2s_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms",0),1), "ms/step", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1) if d.get("e2e") else None)
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -2 gpurun_out/r2s_$tag.err | cut -c1-300
}
run c4cores taskset -c 0-3 python bench.py --no-cpu-baseline --steps 3 --warmup 3
run c2cores taskset -c 0-1 python bench.py --no-cpu-baseline --steps 3 --warmup 3
