"""Kernel time of build variants on a NON-uniform (trained) model — bench.py's workloads use uniform-sigma models.
usage: python tools/variant_timing.py [variants...]   (default: 11 13)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dynamont_b200 import Aligner  # noqa: E402
from dynamont_b200.synth import materialize_model, native_model, synth_read  # noqa: E402

variants = [int(v) for v in sys.argv[1:]] or [11, 13]
path = materialize_model("trained_rna002_5mer", os.path.join(ROOT, "tests", "golden", "_models"))
nm, ns = native_model(path, "rna002")
rng = np.random.default_rng(5)
sigs, seqs = [], []
for _ in range(4096):
    s, q, _ = synth_read(rng, nm, ns, 5, int(rng.integers(500, 3000)), 30.0)
    sigs.append(s.astype(np.float32))
    seqs.append(q)
for v in variants:
    al = Aligner(path, "rna002")
    al.set_option("variant", v)
    cells = sum(al.read_cells(len(s), len(q)) for s, q in zip(sigs, seqs))
    al.align_batch(sigs[:256], seqs[:256], True)
    t0 = time.time()
    al.align_batch(sigs, seqs, True)
    tm = al.last_timing()
    print("trained 5-mer model, variant %d: kernel %.1f ms = %.1f GCUPS (kernel time), %s" % (
        v, tm["dp_ms"], cells / tm["dp_ms"] / 1e6, tm), flush=True)
