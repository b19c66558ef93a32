#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -2 gpurun_out/r2i_pytest.log
DYN_TIMING=1 timeout 900 python bench.py --reads 100000 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1))
PY
grep "dyn timing" gpurun_out/r2i_bench.err | grep -v scratch | tail -12
timeout 600 python bench.py --config c5 --reads 40000 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r2i_c5.json 2> gpurun_out/r2i_c5.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2i_c5.json").read().strip().splitlines()[-1])
print("c5 value", round(d["value"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), d["train"])
PY
