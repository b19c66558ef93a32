#!/bin/bash
# records-free layout (config 4): parity tests, then the c4 bench line (and c2 as the control)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -m gpu -k "c4 or c2_9mer" > gpurun_out/r2t_pytest.log 2>&1; tail -3 gpurun_out/r2t_pytest.log
run() { tag=$1; shift
  if [ -n "$DT" ]; then export DYN_TIMING=1; else unset DYN_TIMING; fi
  timeout 1500 python bench.py "$@" > gpurun_out/r2t_$tag.json 2> gpurun_out/r2t_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2t_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms",0),1), "ms/step", round(d["ms_per_step"],1), "faults", r.get("ribbon_fault_reads"), r.get("ribbon_fault_reasons"), "fb", r.get("log2_fallback_reads"), r.get("lin_retry_reads"))
except Exception as e:
    print("$tag FAILED", e)
PY
  grep "ribbon scratch" gpurun_out/r2t_$tag.err | tail -1 | cut -c1-250
  tail -1 gpurun_out/r2t_$tag.err | cut -c1-250
}
DT=1 run c4 --config c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
DT= run c4e --config c4 --steps 2 --warmup 1 --no-cpu-baseline
# NTK 9-mer: where does the time of one read go (trace + launch list)
DYN_NTK_TRACE=1 timeout 300 python tools/ntk_k9_one.py 60 4 > gpurun_out/r2t_k9.log 2>&1; tail -6 gpurun_out/r2t_k9.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2t_k9_launches.csv python tools/ntk_k9_one.py 60 1 > gpurun_out/r2t_k9_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2t_k9_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    k = r[ki].split("(")[0][:70]; tot[k][0] += 1; tot[k][1] += v
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]): print("%-72s %6d launches %10.2f ms" % (k, n, ms))
PY
