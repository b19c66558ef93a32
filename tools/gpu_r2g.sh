#!/bin/bash
# lanes without memset kernels, BPS 5 default; c5 training; c2v; GPU tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -3 gpurun_out/r2g_pytest.log
run() { # tag args...
  tag=$1; shift
  timeout 1200 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2g_$tag.json 2> gpurun_out/r2g_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_$tag.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1) if d.get("e2e") else None, "ribbon", r.get("ribbon_reads"), r.get("ribbon_fault_reads"), d.get("train"))
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -2 gpurun_out/r2g_$tag.err | cut -c1-300
}
run c2 --steps 3 --warmup 3
run c5 --config c5 --reads 40000 --steps 2 --warmup 2
run c2v --config c2v --reads 20000 --steps 2 --warmup 2 --no-e2e
