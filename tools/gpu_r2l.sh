#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -2 gpurun_out/r2l_pytest.log
run() { tag=$1; shift
  DYN_TIMING=1 timeout 1500 python bench.py --no-cpu-baseline "$@" > gpurun_out/r2l_$tag.json 2> gpurun_out/r2l_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2l_$tag.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],1), "ribbon", r.get("ribbon_reads"), r.get("ribbon_fault_reads"), r.get("ribbon_fault_reasons"))
except Exception as e:
    print("$tag FAILED", e)
PY
  grep "scratch" gpurun_out/r2l_$tag.err | tail -1
}
run c4 --config c4 --reads 1280 --steps 2 --warmup 1 --no-e2e --batch 1280
run c2 --steps 3 --warmup 3
run c1 --config c1 --steps 3 --warmup 3
