#!/bin/bash
# round 2, first GPU session: parity suite + ribbon A/B on an 8192-read sub-run of c2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
for rb in 2 4 0; do
  timeout 600 python bench.py --reads 8192 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --opt ribbon=$rb > gpurun_out/r2a_rb$rb.json 2> gpurun_out/r2a_rb$rb.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2a_rb$rb.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("ribbon=$rb", round(d["value"],1), "GCUPS kernel_ms", round(r["kernel_ms"],1), "frac", round(r["frac"],3), "ribbon", r.get("ribbon_reads"), r.get("ribbon_fault_reads"), "fb", r["lin_retry_reads"], r["log2_fallback_reads"])
except Exception as e:
    print("ribbon=$rb FAILED", e)
PY
done
