"""One 9-mer resquiggle-mode read (the c3 bench shape) — for launch lists / DYN_NTK_TRACE.  usage: python tools/ntk_k9_one.py [length] [n]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dynamont_b200 import Aligner  # noqa: E402
from dynamont_b200.synth import materialize_model, native_model, synth_read  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 60
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
models = os.path.join(ROOT, "tests", "golden", "_models")
path = materialize_model("synthetic_rna004_9mer", models)
nm, ns = native_model(path, "rna004")
rng = np.random.default_rng(77)
reads = [synth_read(rng, nm, ns, 9, L, 12.5) for _ in range(n)]
al = Aligner(path, "rna004", mode="resquiggle")
t0 = time.perf_counter()
res = al.align_batch([r[0].astype(np.float32) for r in reads], [r[1] for r in reads], True)
t1 = time.perf_counter()
print("k9: %d reads of L=%d (S=%d) in %.3f s; ok %d" % (n, L, reads[0][0].size, t1 - t0, sum(isinstance(r, dict) for r in res)))
