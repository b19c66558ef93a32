#!/bin/bash
# log2-domain ribbon in training mode: training soak (noisy kinds now stay on the ribbon tiers), -m gpu suite, c5 line
mkdir -p gpurun_out
python tools/gpu_train_soak.py 24 > gpurun_out/r3v_train_soak.log 2>&1; tail -7 gpurun_out/r3v_train_soak.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3v_pytest.log 2>&1; tail -2 gpurun_out/r3v_pytest.log
timeout 900 python bench.py --config c5 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r3v_c5.json 2> gpurun_out/r3v_c5.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3v_c5.json").read().strip().splitlines()[-1])
print("c5 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), d["train"])
PY
