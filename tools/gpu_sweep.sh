#!/bin/bash
# A/B sweep of the k_align build variants on one B200 (run under gpurun): bench.py on an 8192-read sub-run of c2.
# usage: tools/gpu_sweep.sh TAG "3:fwd_fast=0 3 9 1 4 5 6 7 8"
TAG=$1; shift
mkdir -p gpurun_out
for spec in $1; do
  v=${spec%%:*}; opt=""
  if [[ "$spec" == *:* ]]; then opt="--opt ${spec#*:}"; fi
  python bench.py --reads 8192 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --variant $v $opt \
    > gpurun_out/${TAG}_v${spec//[:=]/_}.json 2> gpurun_out/${TAG}_v${spec//[:=]/_}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_v${spec//[:=]/_}.json").read().strip().splitlines()[-1])
    print("${spec}", d["value"], d.get("roofline",{}).get("frac"), d.get("config",{}).get("log2_fallback_reads"))
except Exception as e:
    print("${spec}", "FAILED", e)
PY
done
