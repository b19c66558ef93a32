#!/bin/bash
# e2e through the async lanes after the allocation fix; BPS A/B of the ribbon kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -2 gpurun_out/r2f_pytest.log
for bps in 6 5 8; do
  timeout 600 python bench.py --reads 20000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --opt rib_bps=$bps > gpurun_out/r2f_bps$bps.json 2> gpurun_out/r2f_bps$bps.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2f_bps$bps.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("bps=$bps", round(d["value"],1), "GCUPS kernel_ms", round(r["kernel_ms"],1), "ribbon", r.get("ribbon_reads"), r.get("ribbon_fault_reads"))
except Exception as e:
    print("bps=$bps FAILED", e)
PY
done
timeout 900 python bench.py --reads 100000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1))
PY
