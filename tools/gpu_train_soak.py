"""Training-statistics soak on the GPU box: pooled Baum-Welch counts of seeded reads of several kinds through the default
tiers (ribbon MODE 2 first) against the full-band log2-domain kernels alone (arith=1, ribbon off) and, for a few reads,
against the CPU oracle.  usage: python tools/gpu_train_soak.py [reads_per_kind]"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dynamont_b200 import Aligner  # noqa: E402
from dynamont_b200.synth import PORE_INFO, materialize_model, native_model, synth_read  # noqa: E402
from oracle import Oracle  # noqa: E402

MODELS = os.path.join(ROOT, "tests", "golden", "_models")
n_per = int(sys.argv[1]) if len(sys.argv) > 1 else 24
KINDS = [("c1", "rna002", "rna002_5mer", (300, 1000), 30, 1.0), ("c2", "rna004", "synthetic_rna004_9mer", (500, 1500), 30, 1.0),
         ("noisy", "rna002", "rna002_5mer", (200, 600), 10, 2.5), ("noisy9", "rna004", "synthetic_rna004_9mer", (300, 700), 9, 2.0),
         ("trained", "rna002", "trained_rna002_5mer", (200, 600), 12, 1.0)]
worst = 0.0
for name, pore, model, (lo, hi), spb, sds in KINDS:
    path = materialize_model(model, MODELS)
    nm, ns = native_model(path, pore)
    k = PORE_INFO[pore][1]
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    sigs, seqs = [], []
    for _ in range(n_per):
        L = int(rng.integers(lo, hi + 1))
        s, q, _ = synth_read(rng, nm, ns, k, L, spb, sd_scale=sds)
        sigs.append(s.astype(np.float32))
        seqs.append(q)
    al = Aligner(path, pore)
    per, pooled = al.train_batch(sigs, seqs)
    tm = al.last_timing()
    ref = Aligner(path, pore)
    ref.set_option("ribbon", 0)
    ref.set_option("arith", 1)
    per2, pooled2 = ref.train_batch(sigs, seqs)
    heavy = pooled2["w"] > 1e-3
    rel = {"w": float(np.max(np.abs(pooled["w"][heavy] - pooled2["w"][heavy]) / pooled2["w"][heavy]))}
    # what the M-step makes of them (NT:519-535): mean = x / w, stdev = sqrt(xx / w - mean^2); the mean is compared with the
    # tests' atol of 1e-5 next to the relative gate (a kmer whose level is ~0 has no relative scale)
    def mstep(p):
        m = p["x"][heavy] / p["w"][heavy]
        return m, np.sqrt(np.maximum(p["xx"][heavy] / p["w"][heavy] - m * m, 0.0))
    (m1_, s1_), (m2_, s2_) = mstep(pooled), mstep(pooled2)
    rel["x"] = float(np.max(np.maximum(np.abs(m1_ - m2_) - 1e-5, 0.0) / np.maximum(np.abs(m2_), 1e-12)))
    # the stdev of a kmer that got its weight from one or two samples is sqrt(xx/w - mean^2) ~ sqrt(0): rounding noise in the
    # reference as well (NT:523-528 floors the variance at 1e-12); compare it where a kmer was actually occupied
    occ = pooled2["w"][heavy] > 5.0
    rel["xx"] = float(np.max(np.abs(s1_[occ] - s2_[occ]) / s2_[occ])) if occ.any() else 0.0
    # a few reads against the oracle's per-read transitions
    orc = Oracle(path, pore)
    dm = 0.0
    for i in range(min(3, n_per)):
        o = orc.train(sigs[i].astype(np.float64), seqs[i])
        r = al.train(sigs[i], seqs[i], as_dicts=False)
        for key in ("m1", "e1", "e2"):
            dm = max(dm, abs(r["transition_params"][key] - o["transition_params"][key]) / o["transition_params"][key])
    worst = max(worst, max(rel.values()), dm)
    print("%-8s %3d reads  ribbon faults %2d  log2 fallback %2d  vs full-band log2: rel weight %.2e  mean (beyond atol 1e-5) %.2e  stdev (weight > 5) %.2e  transitions vs oracle %.2e"
          % (name, n_per, tm["ribbon_faults"], tm["log2_fallback_reads"], rel["w"], rel["x"], rel["xx"], dm))
print("WORST relative difference %.2e (gate 1e-4)" % worst)
