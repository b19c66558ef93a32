#!/bin/bash
# three lanes: c4 with three batches in flight (5 steps), c2 control (two in flight), async GPU tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "async or stream or order or multi" > gpurun_out/r3r_pytest.log 2>&1; tail -2 gpurun_out/r3r_pytest.log
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r3r_$tag.json 2> gpurun_out/r3r_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3r_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms") or 0,1), "ms/step", round(d["ms_per_step"],1), "faults", r.get("ribbon_fault_reads"), "fb", r.get("log2_fallback_reads"), "ok", d["config"]["reads_ok"])
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -1 gpurun_out/r3r_$tag.err | cut -c1-250
}
run c4 --config c4 --steps 5 --warmup 1 --no-cpu-baseline --no-e2e
run c2 --no-cpu-baseline
