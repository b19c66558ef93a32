#!/bin/bash
# wide-row NTK masks (tests + timing), c4 with 4 steps incl. e2e, c3 line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ntk_stages.py tests/test_frontend_driver.py -x -q -m gpu > gpurun_out/r2u_pytest.log 2>&1; tail -3 gpurun_out/r2u_pytest.log
DYN_NTK_TRACE=1 timeout 300 python tools/ntk_k9_one.py 60 4 > gpurun_out/r2u_k9.log 2>&1; tail -5 gpurun_out/r2u_k9.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2u_k9_launches.csv python tools/ntk_k9_one.py 60 1 > gpurun_out/r2u_k9_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2u_k9_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    k = r[ki].split("(")[0][:70]; tot[k][0] += 1; tot[k][1] += v
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]): print("%-72s %6d launches %10.2f ms" % (k, n, ms))
PY
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r2u_$tag.json 2> gpurun_out/r2u_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2u_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],3), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms",0),1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],2), "faults", r.get("ribbon_fault_reads"), "fb", r.get("log2_fallback_reads"), r.get("lin_retry_reads"))
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -1 gpurun_out/r2u_$tag.err | cut -c1-250
}
run c3 --config c3 --steps 2 --warmup 1
run c4 --config c4 --steps 4 --warmup 1 --no-cpu-baseline
