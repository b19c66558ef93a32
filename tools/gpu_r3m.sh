#!/bin/bash
# after the first-row mass check: soak with 48 reads per kind (432 reads), c2 / c5 lines, -m gpu suite
mkdir -p gpurun_out
python tools/gpu_soak.py 48 > gpurun_out/r3m_soak.log 2>&1; tail -11 gpurun_out/r3m_soak.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3m_pytest.log 2>&1; tail -2 gpurun_out/r3m_pytest.log
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r3m_$tag.json 2> gpurun_out/r3m_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3m_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms") or 0,1), "ms/step", round(d["ms_per_step"],1), "faults", r.get("ribbon_fault_reads"), r.get("ribbon_fault_reasons"))
except Exception as e:
    print("$tag FAILED", e)
PY
}
run c2 --no-cpu-baseline
run c2v --config c2v --no-cpu-baseline --no-e2e
run c5 --config c5 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e
