"""Timing of resquiggle (NTK) mode on one read: GPU first path (dyn_ntk_align) vs the repaired reference on one host core.
usage (on the GPU box): python tools/ntk_timing.py [length] [spb]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dynamont_b200 import Aligner  # noqa: E402
from dynamont_b200.synth import materialize_model, native_model, synth_read  # noqa: E402
from oracle import Reference  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 500
spb = float(sys.argv[2]) if len(sys.argv) > 2 else 12.5
models = os.path.join(ROOT, "tests", "golden", "_models")
path = materialize_model("rna004_5mer", models)
nm, ns = native_model(path, "dna_r9")
s, q, _ = synth_read(np.random.default_rng(77), nm, ns, 5, L, spb)
al = Aligner(path, "dna_r9", mode="resquiggle")
al.align(s[:200], q[:20], True)  # warm-up
t0 = time.perf_counter()
r = al.align(s, q, True)
t1 = time.perf_counter()
ref = Reference(path, "dna_r9", mode="resquiggle", ntk_fix=True)
o = ref.align(s, q, True)
t2 = time.perf_counter()
ref_s = t2 - t1
T, N, K = s.size + 1, len(q) - 3, 1024
same = (r["signal_positions"] == o["signal_positions"]).mean() if len(r["states"]) == len(o["states"]) else 0.0
print("NTK k=5 read: L=%d S=%d  dense cells T*N + T*K = %.3g  GPU %.3f s  reference (1 core) %.3f s  speed-up %.1fx  "
      "Z %.6f vs %.6f  segments %d/%d identical borders %.4f polish identical %s" % (
          L, s.size, T * N + T * K, t1 - t0, t2 - t1, (t2 - t1) / (t1 - t0), r["Z"], o["Z"], len(r["states"]), len(o["states"]),
          same, r["polishes"] == o["polishes"]))

# throughput of the batched entry point: n copies of differently seeded reads of the same length
n = int(sys.argv[3]) if len(sys.argv) > 3 else 64
rng = np.random.default_rng(5)
batch = [synth_read(rng, nm, ns, 5, L, spb) for _ in range(n)]
t0 = time.perf_counter()
res = al.align_batch([b[0] for b in batch], [b[1] for b in batch], True)
t1 = time.perf_counter()
ok = sum(isinstance(r, dict) for r in res)
print("NTK batch: %d reads (%d ok) of L=%d in %.2f s = %.1f reads/s  (reference: %.2f reads/s per host core)" % (
    n, ok, L, t1 - t0, n / (t1 - t0), 1.0 / ref_s))
