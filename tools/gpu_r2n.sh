#!/bin/bash
# 2-GPU check: c2 (reads shard, no collective) and c5 (NCCL all-reduce of the device-resident statistics)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --reads 40000 --steps 2 --warmup 2 > gpurun_out/r2n_c2_2gpu.json 2> gpurun_out/r2n_c2_2gpu.err
tail -c 600 gpurun_out/r2n_c2_2gpu.json | head -c 600; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c5 --reads 40000 --steps 3 --warmup 2 --no-e2e > gpurun_out/r2n_c5_2gpu.json 2> gpurun_out/r2n_c5_2gpu.err
python - <<PY
import json
for f in ("r2n_c2_2gpu","r2n_c5_2gpu"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],1), d.get("train"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/r2n_c5_2gpu.err | cut -c1-300
