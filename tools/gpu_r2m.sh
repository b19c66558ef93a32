#!/bin/bash
# default bench (value through the async lanes) + ncu capture at the bench's launch shape + launch list + train capture
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -2 gpurun_out/r2m_pytest.log
timeout 1500 python bench.py --no-cpu-baseline > gpurun_out/r2m_c2.json 2> gpurun_out/r2m_c2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2m_c2.json").read().strip().splitlines()[-1])
r=d["roofline"]
print("c2 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1), "faults", r.get("ribbon_fault_reads"))
PY
# ncu: one ribbon launch of 20000 reads (the bench's batch), align and train
CMD="python bench.py --config c2 --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/prof_r2m_align_plain.json 2> gpurun_out/prof_r2m_align_plain.err
ncu --set full --clock-control none --import-source on -k regex:k_ribbon -c 1 -f -o gpurun_out/prof_r2m_align $CMD > gpurun_out/prof_r2m_align_ncu.log 2>&1
tail -1 gpurun_out/prof_r2m_align_ncu.log
CMD="python bench.py --config c5 --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/prof_r2m_train_plain.json 2> gpurun_out/prof_r2m_train_plain.err
ncu --set full --clock-control none --import-source on -k regex:k_ribbon -c 1 -f -o gpurun_out/prof_r2m_train $CMD > gpurun_out/prof_r2m_train_ncu.log 2>&1
tail -1 gpurun_out/prof_r2m_train_ncu.log
# launch list of a step
CMD="python bench.py --config c2 --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2m.csv $CMD > gpurun_out/launches_r2m.log 2>&1
tail -1 gpurun_out/launches_r2m.log
