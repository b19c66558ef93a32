#!/bin/bash
# verification of the last commit of the third session: -m gpu suite, default line (what the driver runs), c3, c5, smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3g_pytest.log 2>&1; tail -3 gpurun_out/r3g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3g_smoke.log 2>&1; tail -1 gpurun_out/r3g_smoke.log | cut -c1-200
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r3g_$tag.json 2> gpurun_out/r3g_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3g_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],3), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms") or 0,1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],2), "faults", r.get("ribbon_fault_reads"), "fb", r.get("log2_fallback_reads"), d.get("train"), (d.get("config") or {}).get("k5", {}).get("reads_per_s"))
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -1 gpurun_out/r3g_$tag.err | cut -c1-200
}
run c2
run c5 --config c5 --steps 3 --warmup 2 --no-cpu-baseline
run c3 --config c3 --steps 2 --warmup 1
