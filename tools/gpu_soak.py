"""Soak test on the GPU box: many seeded reads of different kinds through the C ABI against the CPU oracle (plain-C
restatement, bit-exact with the reference).  Reports border identity, |dp|, Z error and how many reads took the log2-domain
fallback.  usage: [DYN_SOAK_VARIANT=n] python tools/gpu_soak.py [reads_per_kind]"""
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dynamont_b200 import Aligner  # noqa: E402
from dynamont_b200.synth import PORE_INFO, low_complexity_digits, materialize_model, native_model, synth_read  # noqa: E402
from oracle import Oracle  # noqa: E402

MODELS = os.path.join(ROOT, "tests", "golden", "_models")
n_per = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ONLY = os.environ.get("DYN_SOAK_KINDS", "").split()
KINDS = [  # name, pore, model, (min,max) length, spb, dwell, sd_scale, kind, outlier rate
    ("c1", "rna002", "rna002_5mer", (300, 1000), 30, "geometric", 1.0, "rand", 0.0),
    ("c2", "rna004", "synthetic_rna004_9mer", (500, 1500), 30, "geometric", 1.0, "rand", 0.0),
    ("dense", "dna_r9", "rna004_5mer", (200, 900), 5, "geometric", 1.0, "rand", 0.0),
    ("noisy", "rna002", "rna002_5mer", (200, 600), 10, "geometric", 2.5, "rand", 0.0),
    ("trained", "rna002", "trained_rna002_5mer", (200, 600), 12, "gamma", 1.0, "rand", 0.0),
    ("outliers", "rna002", "rna002_5mer", (200, 600), 12, "geometric", 1.0, "rand", 0.004),
    ("homop", "rna002", "rna002_5mer", (200, 500), 8, "geometric", 1.5, "homopolymer", 0.0),
    ("dinuc", "rna002", "rna002_5mer", (200, 400), 6, "geometric", 2.0, "dinuc", 0.0),
    ("mixed9", "rna004", "synthetic_rna004_9mer", (300, 700), 9, "geometric", 1.5, "mixed", 0.0),
    # long noisy reads (many 16-row groups at 1.5x / 2x the model's noise); the oracle needs ~10 s per read: run them with
    # DYN_SOAK_KINDS="noisy_long noisy2_long" and a small count
    ("noisy_long", "rna004", "synthetic_rna004_9mer", (2000, 4000), 30, "geometric", 1.5, "rand", 0.0),
    ("noisy2_long", "rna002", "rna002_5mer", (1500, 3000), 20, "gamma", 2.0, "rand", 0.0),
]
if not ONLY:
    KINDS = [kd for kd in KINDS if not kd[0].endswith("_long")]
tot_seg = tot_same = tot_fb = tot_reads = 0
worst_dp = worst_z = 0.0
for name, pore, model, (lo, hi), spb, dwell, sds, kind, outl in KINDS:
    if ONLY and name not in ONLY:
        continue
    path = materialize_model(model, MODELS)
    nm, ns = native_model(path, pore)
    k = PORE_INFO[pore][1]
    orc = Oracle(path, pore)
    al = Aligner(path, pore)
    if os.environ.get("DYN_SOAK_VARIANT"):
        al.set_option("variant", int(os.environ["DYN_SOAK_VARIANT"]))  # kernel build variant (csrc/engine.cu)
    for kv in os.environ.get("DYN_SOAK_OPTS", "").split():
        al.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    sigs, seqs = [], []
    for _ in range(n_per):
        L = int(rng.integers(lo, hi + 1))
        digs = None if kind == "rand" else low_complexity_digits(rng, L, kind, k)
        s, q, _ = synth_read(rng, nm, ns, k, L, spb, dwell=dwell, sd_scale=sds, seq_digits=digs)
        if outl > 0:
            m = rng.random(s.size) < outl
            s[m] += rng.choice([-1, 1], m.sum()) * rng.uniform(3, 8, m.sum())
        sigs.append(s.astype(np.float32))
        seqs.append(q)
    t0 = time.time()
    res = al.align_batch(sigs, seqs, True)
    fb = al.last_timing()["log2_fallback_reads"]
    rl = al.last_timing()["lin_retry_reads"]
    seg = same = 0
    dpm = zm = 0.0
    bad = 0
    for s, q, r in zip(sigs, seqs, res):
        try:
            o = orc.align(s.astype(np.float64), q, True)
        except RuntimeError as e:
            assert isinstance(r, RuntimeError) and str(r) == str(e), (name, r, e)
            continue
        assert isinstance(r, dict), (name, r)
        eq = r["signal_positions"] == o["signal_positions"]
        ok = eq.copy()
        ok[:-1] &= eq[1:]
        seg += eq.size
        same += int(eq.sum())
        bad += int(eq.sum() != eq.size)
        dpm = max(dpm, float(np.abs(r["probabilities"] - o["probabilities"])[ok].max(initial=0.0)))
        zm = max(zm, abs(r["Z"] - o["Z"]) / max(1.0, abs(o["Z"])))
    print("%-9s %3d reads  borders %d/%d (%d reads with a moved border)  max|dp| %.2e  max dZ/|Z| %.2e  period-4 retry %d  log2 fallback %d  %.1fs" % (
        name, n_per, same, seg, bad, dpm, zm, rl, fb, time.time() - t0), flush=True)
    tot_seg += seg; tot_same += same; tot_fb += fb; tot_reads += n_per
    worst_dp = max(worst_dp, dpm); worst_z = max(worst_z, zm)
print("TOTAL %d reads: borders identical %.5f %%, max|dp| %.2e, max dZ/|Z| %.2e, fallback reads %d" % (
    tot_reads, 100.0 * tot_same / max(tot_seg, 1), worst_dp, worst_z, tot_fb))
