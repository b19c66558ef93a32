#!/bin/bash
# last verification of HEAD: -m gpu suite, smoke, the default bench line and the reference arm exactly as the driver runs them
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r3t_pytest.log 2>&1; tail -2 gpurun_out/r3t_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3t_smoke.log 2>&1; tail -1 gpurun_out/r3t_smoke.log | cut -c1-120
timeout 1500 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/r3t_ref.json 2> gpurun_out/r3t_ref.err
timeout 1500 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r3t_c2.json 2> gpurun_out/r3t_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3t_c2.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("c2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "cpu", round(d["cpu_baseline"]["value"],4), "launches", d["gpu_launches"], "clocks", d["clocks"])
d=json.loads(open("gpurun_out/r3t_ref.json").read().strip().splitlines()[-1])
print("ref value", d["value"], d["cpu_baseline"]["cores"], d["ms_per_step"])
PY
