"""Generate tests/golden/nt_golden.npz from the UNMODIFIED reference C++ (oracle/_ref, needs /root/reference).

Each case stores its inputs (signal as FP32-representable float64, sequence) and the reference's outputs of
Aligner.align(calc_probabilities=True) and Aligner.train (per-read M-step + transitions), plus Zf/Zb and the
raw per-kmer sufficient statistics obtained from the reference's private forward/backward (ref_shim.cpp).
Run:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dynamont_b200.synth import low_complexity_digits, materialize_model, native_model, synth_read  # noqa: E402
from oracle import Reference  # noqa: E402

MODELS = os.path.join(ROOT, "tests", "golden", "_models")
CASES = [  # name, pore, model, length, spb, dwell, sd_scale, kind, seed
    ("rna002_short", "rna002", "rna002_5mer", 60, 10, "geometric", 1.0, "rand", 101),
    ("rna002_300", "rna002", "rna002_5mer", 300, 30, "geometric", 1.0, "rand", 102),
    ("rna002_band", "rna002", "rna002_5mer", 650, 9, "geometric", 1.0, "rand", 103),
    ("rna002_trained", "rna002", "trained_rna002_5mer", 500, 12, "gamma", 1.0, "rand", 104),
    ("dna_r9_500", "dna_r9", "rna004_5mer", 500, 8, "geometric", 1.0, "rand", 105),
    ("rna002_homopolymer", "rna002", "rna002_5mer", 500, 8, "geometric", 1.5, "homopolymer", 106),
    ("rna002_dinuc", "rna002", "rna002_5mer", 300, 8, "geometric", 1.5, "dinuc", 107),
    ("rna002_mixed", "rna002", "rna002_5mer", 600, 10, "geometric", 1.5, "mixed", 108),
    ("rna002_min_dwell", "rna002", "rna002_5mer", 200, 2, "geometric", 1.0, "rand", 109),
    ("rna004_9mer", "rna004", "synthetic_rna004_9mer", 400, 12, "geometric", 1.0, "rand", 110),
    ("dna_r10_9mer", "dna_r10_400bps", "synthetic_rna004_9mer", 520, 12, "gamma", 1.0, "rand", 111),
]

out = {"names": np.array([c[0] for c in CASES])}
for name, pore, model, L, spb, dwell, sds, kind, seed in CASES:
    path = materialize_model(model, MODELS)
    ref = Reference(path, pore)
    nm, ns = native_model(path, pore)
    rng = np.random.default_rng(seed)
    digs = None if kind == "rand" else low_complexity_digits(rng, L, kind, ref.k)
    sig, seq, _ = synth_read(rng, nm, ns, ref.k, L, spb, dwell=dwell, sd_scale=sds, seq_digits=digs)
    a = ref.align(sig, seq, True)
    t = ref.train(sig, seq)
    st = ref.stages(sig, seq, rows=[], stats=True)
    touched = np.nonzero(st["w"] > 0)[0]
    out[name + "/pore"] = np.array(pore)
    out[name + "/model"] = np.array(model)
    out[name + "/signal"] = sig.astype(np.float32)
    out[name + "/sequence"] = np.array(seq)
    out[name + "/Z"] = np.array(a["Z"])
    out[name + "/Zf"] = np.array(st["Zf"])
    out[name + "/sequence_positions"] = a["sequence_positions"]
    out[name + "/signal_positions"] = a["signal_positions"]
    out[name + "/probabilities"] = a["probabilities"]
    out[name + "/train_Z"] = np.array(t["Z"])
    out[name + "/train_trans"] = np.array([t["transition_params"][q] for q in ("m1", "e1", "e2")])
    out[name + "/train_kmers"] = touched.astype(np.int64)
    out[name + "/train_mean"] = t["emission_model"]["mean"][touched]
    out[name + "/train_stdev"] = t["emission_model"]["stdev"][touched]
    out[name + "/stat_w"] = st["w"][touched]
    out[name + "/stat_x"] = st["sx"][touched]
    out[name + "/stat_xx"] = st["sxx"][touched]
    # untouched kmers keep the model (NT:531-534): spot-check value
    un = np.nonzero(st["w"] == 0)[0]
    assert np.array_equal(t["emission_model"]["mean"][un], ref.model()[0][un])
    print(name, "S=%d L=%d Z=%.6f segments=%d touched=%d" % (sig.size, len(seq), a["Z"], a["signal_positions"].size, touched.size))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nt_golden.npz"), **out)
print("written", os.path.getsize(os.path.join(ROOT, "tests", "golden", "nt_golden.npz")), "bytes")
