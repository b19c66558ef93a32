#!/bin/bash
# final measurements of round 2 (third session): -m gpu suite, bench lines, ncu captures of the final kernels, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r2f_$tag.json 2> gpurun_out/r2f_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2f_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],3), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms",0),1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],2), "faults", r.get("ribbon_fault_reads"), "fb", r.get("log2_fallback_reads"), r.get("lin_retry_reads"), d.get("train"))
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -1 gpurun_out/r2f_$tag.err | cut -c1-200
}
run c2 
run c2v --config c2v --no-cpu-baseline --steps 3 --warmup 2
run c5 --config c5 --steps 3 --warmup 2 --no-cpu-baseline
run c1 --config c1 --no-cpu-baseline
run c4 --config c4 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e
run c3 --config c3 --steps 2 --warmup 1
run ref --impl reference
# ncu: one ribbon launch of 20000 reads (the bench's batch), align and train
CMD="python bench.py --config c2 --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/prof_r2f_align_plain.json 2> gpurun_out/prof_r2f_align_plain.err
ncu --set full --clock-control none --import-source on -k regex:k_ribbon -c 1 -f -o gpurun_out/prof_r2f_align $CMD > gpurun_out/prof_r2f_align_ncu.log 2>&1
tail -1 gpurun_out/prof_r2f_align_ncu.log
CMD="python bench.py --config c5 --reads 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/prof_r2f_train_plain.json 2> gpurun_out/prof_r2f_train_plain.err
ncu --set full --clock-control none --import-source on -k regex:k_ribbon -c 1 -f -o gpurun_out/prof_r2f_train $CMD > gpurun_out/prof_r2f_train_ncu.log 2>&1
tail -1 gpurun_out/prof_r2f_train_ncu.log
# launch list of a step
CMD="python bench.py --config c2 --reads 20000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2f.csv $CMD > gpurun_out/launches_r2f.log 2>&1
tail -1 gpurun_out/launches_r2f.log
