#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -12 gpurun_out/r2p_pytest.log
DYN_NTK_TRACE=1 python - <<'PY' 2>&1 | tail -4
import sys, time
sys.path.insert(0, "tests")
from test_ntk_stages import load_ntk
from dynamont_b200 import Aligner
c = [x for x in load_ntk() if x.name == "ntk_rna004_9mer_T1000"][0]
al = Aligner(c.model_path, c.pore, mode="resquiggle")
t0 = time.perf_counter(); r = al.align(c.signal, c.sequence, True); t1 = time.perf_counter()
print("9-mer T=%d NTK align: %.2f s, segments %d, Z %.6f vs %.6f" % (c.signal.size + 1, t1 - t0, len(r["states"]), r["Z"], c.align_Z))
PY
timeout 1500 python bench.py --config c4 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2p_c4.json 2> gpurun_out/r2p_c4.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2p_c4.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("c4 value", round(d["value"],1), "kernel_ms", round(r["kernel_ms"],1), "ms/step", round(d["ms_per_step"],1), "reads/s", round(d["reads_per_s"],1), "ribbon", r.get("ribbon_reads"), r.get("ribbon_fault_reads"), r.get("ribbon_fault_reasons"))
except Exception as e:
    print("c4 FAILED", e)
PY
tail -2 gpurun_out/r2p_c4.err | cut -c1-200
