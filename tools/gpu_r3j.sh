#!/bin/bash
# end-to-end: pinned H2D bandwidth of the box, batch size 10000 vs 20000, host phases of the e2e loop
mkdir -p gpurun_out
python - <<'PY'
import torch, time
n = 1_600_000_000
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("pinned H2D %.1f GB/s" % (n * 4 / dt / 1e9))
torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("pinned D2H %.1f GB/s" % (n * 4 / dt / 1e9))
PY
run() { tag=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/r3j_$tag.json 2> gpurun_out/r3j_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r3j_$tag.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("$tag value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, "kernel_ms", round(r.get("kernel_ms") or 0,1), "ms/step", round(d["ms_per_step"],1), "e2e ms", round(d["e2e"]["ms_per_step"],1) if d.get("e2e") else None)
except Exception as e:
    print("$tag FAILED", e)
PY
}
run b20k --no-cpu-baseline
run b10k --no-cpu-baseline --batch 10000
run b20k_s6 --no-cpu-baseline --steps 6
