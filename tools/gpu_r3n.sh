#!/bin/bash
# training soak (default tiers vs full-band log2), the -m gpu suite of HEAD, the default bench line of HEAD
mkdir -p gpurun_out
python tools/gpu_train_soak.py 24 > gpurun_out/r3n_train_soak.log 2>&1; tail -7 gpurun_out/r3n_train_soak.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3n_pytest.log 2>&1; tail -2 gpurun_out/r3n_pytest.log
timeout 1500 python bench.py > gpurun_out/r3n_c2.json 2> gpurun_out/r3n_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3n_c2.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("c2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "cpu", d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
PY
