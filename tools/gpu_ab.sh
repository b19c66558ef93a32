#!/bin/bash
# A/B of experiment builds (tools/build_exp.sh) on one GPU box: bench.py on an 8192-read sub-run of c2.
# usage: tools/gpu_ab.sh TAG VARIANT "name1 name2 ..." [extra bench args]
TAG=$1; V=$2; NAMES=$3; shift; shift; shift
mkdir -p gpurun_out
for n in $NAMES; do
  DYNAMONT_B200_LIB=build_exp/libdyn_$n.so python bench.py --reads 8192 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --variant $V $* \
    > gpurun_out/${TAG}_$n.json 2> gpurun_out/${TAG}_$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$n.json").read().strip().splitlines()[-1])
    print("$n", round(d["value"],1), round(d["roofline"]["achieved"],1), d["roofline"]["frac"])
except Exception as e:
    print("$n", "FAILED", e)
PY
done
