"""Design study: linear-domain FP32 model (lin_model.py) vs the double-precision oracle on stress reads.
usage: python tools/lin_stress.py [n_reads_per_kind]"""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dynamont_b200.synth import (PORE_INFO, encode_kmers, low_complexity_digits, materialize_model, native_model,  # noqa: E402
                                 synth_read)
from lin_model import align_lin  # noqa: E402
from oracle import PORES, Oracle  # noqa: E402

MODELS = os.path.join(ROOT, "tests", "golden", "_models")
n_per = int(sys.argv[1]) if len(sys.argv) > 1 else 2
KINDS = [  # name, pore, model, length, spb, dwell, sd_scale, kind, outliers
    ("c1", "rna002", "rna002_5mer", 600, 30, "geometric", 1.0, "rand", 0.0),
    ("c2", "rna004", "synthetic_rna004_9mer", 700, 30, "geometric", 1.0, "rand", 0.0),
    ("noisy", "rna002", "rna002_5mer", 400, 10, "geometric", 2.5, "rand", 0.0),
    ("spikes", "rna002", "rna002_5mer", 400, 12, "geometric", 1.0, "rand", 0.01),
    ("spikes9", "rna004", "synthetic_rna004_9mer", 400, 12, "gamma", 1.0, "rand", 0.02),
    ("homop", "rna002", "rna002_5mer", 400, 8, "geometric", 1.5, "homopolymer", 0.0),
    ("dinuc", "rna002", "rna002_5mer", 300, 6, "geometric", 2.0, "dinuc", 0.0),
    ("mixed", "rna004", "synthetic_rna004_9mer", 500, 9, "geometric", 1.5, "mixed", 0.005),
    ("wrongmodel", "dna_r9", "rna002_5mer", 400, 10, "geometric", 1.0, "rand", 0.0),
]
code = {c: i for i, c in enumerate("ACGT")}
tot = {}
for name, pore, model, L, spb, dwell, sds, kind, outl in KINDS:
    path = materialize_model(model, MODELS)
    orc = Oracle(path, pore)
    nm, ns = native_model(path, pore)
    rna, k = PORE_INFO[pore]
    gen_nm = nm
    if name == "wrongmodel":  # signal generated from a different table than the one used for alignment
        gen_nm, _ = native_model(materialize_model("rna004_5mer", MODELS), "dna_r9")
    for i in range(n_per):
        rng = np.random.default_rng(7000 + 100 * len(tot) + i)
        digs = None if kind == "rand" else low_complexity_digits(rng, L, kind, k)
        sig, seq, _ = synth_read(rng, gen_nm, ns, k, L, spb, dwell=dwell, sd_scale=sds, seq_digits=digs)
        if outl > 0:
            m = rng.random(sig.size) < outl
            sig[m] += rng.choice([-1, 1], m.sum()) * rng.uniform(2, 6, m.sum())
            sig = sig.astype(np.float32).astype(np.float64)
        try:
            o = orc.align(sig, seq, True)
        except RuntimeError as e:
            print(name, i, "oracle:", e)
            continue
        km = encode_kmers(np.array([code[ch] for ch in seq]), k)
        tr = [math.log(v) for v in PORES[pore][2]]
        t0 = time.time()
        r = align_lin(sig, km, nm, ns, tr, k)
        if "signal_positions" not in r or r["fault"]:
            print("%-10s %d FAULT %s" % (name, i, r["fault"][:3]))
            tot.setdefault(name, []).append(("fault",))
            continue
        same = r["signal_positions"] == o["signal_positions"]
        ok = same.copy()
        ok[:-1] &= same[1:]
        dp = np.abs(r["probabilities"] - o["probabilities"])[ok].max(initial=0.0)
        print("%-10s %d S=%d borders %d/%d max|dp| %.2e dZrel %.2e massdev %.1e lanes/row %.2f  %.1fs" % (
            name, i, sig.size, same.sum(), same.size, dp, abs(r["Z"] - o["Z"]) / abs(o["Z"]), r["mass_dev"],
            r["records_per_row"], time.time() - t0), flush=True)
        tot.setdefault(name, []).append((same.sum(), same.size, dp))
