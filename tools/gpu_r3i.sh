#!/bin/bash
# NTK 5-mer batch of 192 reads with 192 workers (was 64); 2-GPU default line of the final build
mkdir -p gpurun_out
timeout 600 python tools/ntk_timing.py 1000 12.5 192 > gpurun_out/r3i_k5.log 2>&1; grep "^NTK" gpurun_out/r3i_k5.log
timeout 900 python bench.py --config c3 --steps 2 --warmup 1 > gpurun_out/r3i_c3.json 2> gpurun_out/r3i_c3.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r3i_c3.json").read().strip().splitlines()[-1])
    print("c3 k9", d["config"]["k9"]["reads_per_s"], "reads/s", d["config"]["k9"]["ms_per_step"], "ms; k5", d["config"]["k5"]["reads_per_s"], "reads/s", d["config"]["k5"]["ms_per_step"], "ms")
except Exception as e:
    print("c3 FAILED", e)
PY
tail -1 gpurun_out/r3i_c3.err | cut -c1-300
