#!/bin/bash
# 2 GPUs, final build: the default line under torchrun (what the driver's scaling run launches) and c5
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r3k_c2_2gpu.json 2> gpurun_out/r3k_c2_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c5 --steps 3 --warmup 2 > gpurun_out/r3k_c5_2gpu.json 2> gpurun_out/r3k_c5_2gpu.err
python - <<'PY'
import json
for tag in ("c2_2gpu", "c5_2gpu"):
    try:
        d=json.loads([l for l in open("gpurun_out/r3k_%s.json"%tag).read().strip().splitlines() if l.startswith("{")][-1])
        print(tag, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "ms/step", round(d["ms_per_step"],1), "kernel_ms", round(d["roofline"]["kernel_ms"],1), d.get("train"))
    except Exception as e:
        print(tag, "FAILED", e)
PY
tail -2 gpurun_out/r3k_c5_2gpu.err | cut -c1-300
