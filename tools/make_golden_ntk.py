"""Generate tests/golden/ntk_golden.npz from the reference C++ (oracle/_ref, needs /root/reference).

Resquiggle ("NTK") mode.  Two kinds of vectors per case:
  * stage outputs of the code that works in the reference AS SHIPPED (NTK_aligner_api.cpp:197-441): the row masks
    tnMap / tkMap of the dense pre-passes, the sorted sparse-lattice keys, Zf/Zb of both pre-passes, the transition
    table — from the UNMODIFIED reference (oracle/_ref/libdynamont_ref.so);
  * the end-to-end alignment (segments with polish kmers) — from the reference with the two-line repair of
    logF / logB described in SURVEY.md F2 (oracle/build.py: build_reference_ntkfix), because the unmodified
    reference throws for every input in this mode.
Masks are stored as index lists (row pointer + column indices).  Run:  python tools/make_golden_ntk.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from dynamont_b200.synth import materialize_model, native_model, synth_read  # noqa: E402
from oracle import Reference  # noqa: E402

MODELS = os.path.join(ROOT, "tests", "golden", "_models")
CASES = [  # name, pore, model, length, spb, seed, want end-to-end
    ("ntk_rna002_60", "rna002", "rna002_5mer", 60, 10, 201, True),
    ("ntk_rna002_120", "rna002", "rna002_5mer", 120, 8, 202, True),
    ("ntk_dna_r9_80", "dna_r9", "rna004_5mer", 80, 9, 203, True),
    ("ntk_rna004_9mer_14", "rna004", "synthetic_rna004_9mer", 14, 6, 204, False),
    # config 3 is 9-mer: one polyA-prefixed 9-mer read WITH an end-to-end alignment (reference: ~25 s, 1.7 GB for T ~ 150)
    ("ntk_rna004_9mer_polyA", "rna004", "synthetic_rna004_9mer", 24, 9, 205, True),
]
# a 9-mer read at T ~ 1000 (the dense TK lattice alone is 5 x 1000 x 262144 doubles = 10.5 GB in the reference, minutes of
# CPU): written to its own file by `python tools/make_golden_ntk.py big`
BIG = [("ntk_rna004_9mer_T1000", "rna004", "synthetic_rna004_9mer", 90, 12, 206, True)]
OUT_NAME = "ntk_golden.npz"
if len(sys.argv) > 1 and sys.argv[1] == "big":
    CASES, OUT_NAME = BIG, "ntk_golden_9mer_T1000.npz"


def csr(mask):
    ptr = np.concatenate(([0], np.cumsum(mask.sum(1)))).astype(np.int64)
    return ptr, np.nonzero(mask)[1].astype(np.int32)


out = {"names": np.array([c[0] for c in CASES])}
for name, pore, model, L, spb, seed, e2e in CASES:
    path = materialize_model(model, MODELS)
    nm, ns = native_model(path, pore)
    ref = Reference(path, pore, mode="resquiggle")
    rng = np.random.default_rng(seed)
    sig, seq, _ = synth_read(rng, nm, ns, ref.k, L, spb)
    st = ref.ntk_prepass(sig, seq)
    out[name + "/pore"] = np.array(pore)
    out[name + "/model"] = np.array(model)
    out[name + "/signal"] = sig.astype(np.float32)
    out[name + "/sequence"] = np.array(seq)
    out[name + "/tn_ptr"], out[name + "/tn_idx"] = csr(st["tn"])
    out[name + "/tk_ptr"], out[name + "/tk_idx"] = csr(st["tk"])
    out[name + "/keys"] = st["keys"]
    out[name + "/Z"] = st["Z"]
    out[name + "/transitions"] = st["transitions"]
    msg = "S=%d L=%d keys=%d Ztn=%.6f Ztk=%.6f" % (sig.size, len(seq), st["keys"].size, st["Z"][1], st["Z"][3])
    if e2e:
        fix = Reference(path, pore, mode="resquiggle", ntk_fix=True)
        a = fix.align(sig, seq, True)
        out[name + "/align_Z"] = np.array(a["Z"])
        out[name + "/sequence_positions"] = a["sequence_positions"]
        out[name + "/signal_positions"] = a["signal_positions"]
        out[name + "/probabilities"] = a["probabilities"]
        out[name + "/states"] = np.array(a["states"])
        out[name + "/polishes"] = np.array(a["polishes"])
        msg += " align Z=%.6f segments=%d" % (a["Z"], len(a["states"]))
    print(name, msg)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", OUT_NAME), **out)
print("written", os.path.getsize(os.path.join(ROOT, "tests", "golden", OUT_NAME)), "bytes")
