#!/bin/bash
# NTK 5-mer batch: per-read trace (where do 192 one-kb reads spend 25 s?)
mkdir -p gpurun_out
DYN_NTK_TRACE=1 timeout 600 python tools/ntk_timing.py 1000 12.5 64 > gpurun_out/r3h_k5.log 2>&1
grep -c "ntk read" gpurun_out/r3h_k5.log; grep "ntk read" gpurun_out/r3h_k5.log | head -3; grep "ntk read" gpurun_out/r3h_k5.log | tail -3; grep "^NTK" gpurun_out/r3h_k5.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3h_k5_launches.csv python tools/ntk_timing.py 1000 12.5 1 > gpurun_out/r3h_k5_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r3h_k5_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    k = r[ki].split("(")[0][:70]; tot[k][0] += 1; tot[k][1] += v
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:10]: print("%-72s %6d launches %10.2f ms" % (k, n, ms))
PY
