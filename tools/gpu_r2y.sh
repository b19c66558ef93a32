#!/bin/bash
# 2 GPUs: the default bench line under torchrun (what the driver's scaling run launches), then c5 on 2 GPUs; c3 on one
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2y_c2_2gpu.json 2> gpurun_out/r2y_c2_2gpu.err
python - <<'PY'
import json
for tag in ("c2_2gpu",):
    try:
        d=json.loads([l for l in open("gpurun_out/r2y_%s.json"%tag).read().strip().splitlines() if l.startswith("{")][-1])
        print(tag, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "ms/step", round(d["ms_per_step"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],1))
    except Exception as e:
        print(tag, "FAILED", e)
PY
tail -2 gpurun_out/r2y_c2_2gpu.err | cut -c1-300
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --config c3 --steps 2 --warmup 1 > gpurun_out/r2y_c3.json 2> gpurun_out/r2y_c3.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2y_c3.json").read().strip().splitlines()[-1])
    print("c3 k9", d["config"]["k9"]["reads_per_s"], "reads/s", d["config"]["k9"]["ms_per_step"], "ms; k5", d["config"]["k5"]["reads_per_s"], "reads/s", d["config"]["k5"]["ms_per_step"], "ms")
except Exception as e:
    print("c3 FAILED", e)
PY
tail -1 gpurun_out/r2y_c3.err | cut -c1-300
