#!/bin/bash
# A/B: 6 resident CTAs per SM for the packed ribbon kernel; NTK 9-mer launch list + per-row selection sizes; c4 with 4 steps
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r2x_$tag.json 2> gpurun_out/r2x_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2x_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms",0),1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],1), "faults", r.get("ribbon_fault_reads"), "fb", r.get("log2_fallback_reads"), r.get("lin_retry_reads"), r.get("ribbon_fault_reasons"))
except Exception as e:
    print("$tag FAILED", e)
PY
  tail -1 gpurun_out/r2x_$tag.err | cut -c1-250
}
run c2bps6 --no-cpu-baseline --no-e2e --opt rib_bps=6
run c2bps5 --no-cpu-baseline --no-e2e
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2x_k9_launches.csv python tools/ntk_k9_one.py 60 1 > gpurun_out/r2x_k9_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2x_k9_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    k = r[ki].split("(")[0][:70]; tot[k][0] += 1; tot[k][1] += v
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:8]: print("%-72s %6d launches %10.2f ms" % (k, n, ms))
PY
python - <<'PY'
import os, sys, numpy as np
sys.path.insert(0, ".")
from dynamont_b200 import Aligner
from dynamont_b200.synth import materialize_model, native_model, synth_read
path = materialize_model("synthetic_rna004_9mer", "tests/golden/_models")
nm, ns = native_model(path, "rna004")
s, q, _ = synth_read(np.random.default_rng(77), nm, ns, 9, 60, 12.5)
al = Aligner(path, "rna004", mode="resquiggle")
r = al.ntk_prepass(s.astype(np.float32), q)
tk = r["tk_mask"] if "tk_mask" in r else None
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in r.items()})
if tk is not None:
    cnt = np.array([int(np.unpackbits(row.view(np.uint8)).sum()) for row in tk])
    print("selected kmers per row: min %d median %d mean %.0f max %d; rows > 2048: %d, > 65536: %d of %d" % (cnt.min(), np.median(cnt), cnt.mean(), cnt.max(), (cnt > 2048).sum(), (cnt > 65536).sum(), cnt.size))
PY
run c4 --config c4 --steps 4 --warmup 1 --no-cpu-baseline --no-e2e
