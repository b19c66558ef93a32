"""Front-end stages either side of the DP: N1 signal normalisation + Hampel filter (GPU), N2 CSV
formatting (host).  Known answers: tests/golden/frontend_golden.npz, produced by executing the reference's own Python
source (tools/make_golden_frontend.py), including the vectors of the reference's tests (tests/test_utils.py:7-14,
tests/test_segment.py:179-200)."""
import os
import sys

import numpy as np
import pytest

from conftest import MODELS_DIR, ROOT

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
GOLDEN_FE = os.path.join(ROOT, "tests", "golden", "frontend_golden.npz")


def _aligner(lib=None, pore="rna002"):
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    return Aligner(materialize_model("rna002_5mer", MODELS_DIR), pore, _lib_path=lib)


def _check_preprocess(al):
    with np.load(GOLDEN_FE) as z:
        n = int(z["n_pre"])
        by_params = {}
        for i in range(n):
            shift, scale, window, nsig = z["pre%d/params" % i]
            by_params.setdefault((int(window), float(nsig)), []).append((z["pre%d/raw" % i], shift, scale, z["pre%d/expected" % i]))
        for (window, nsig), items in by_params.items():
            outs = al.preprocess_batch([it[0] for it in items], [it[1] for it in items], [it[2] for it in items], window, nsig)
            for (raw, shift, scale, exp), got in zip(items, outs):
                # float64 arithmetic on both sides, rounded to FP32 at the end: bit-exact
                assert np.array_equal(got, exp.astype(np.float32)), (window, nsig, raw.size)


def _check_format(al_by_rna):
    with np.load(GOLDEN_FE) as z:
        for i in range(int(z["n_fmt"])):
            k, rna = [int(v) for v in z["fmt%d/params" % i]]
            res = {"sequence_positions": z["fmt%d/sequence_positions" % i], "signal_positions": z["fmt%d/signal_positions" % i],
                   "probabilities": z["fmt%d/probabilities" % i], "states": z["fmt%d/states" % i].tolist(),
                   "polishes": z["fmt%d/polishes" % i].tolist()}
            al = al_by_rna[bool(rna)]
            al.kmer_size = k  # the formatter only needs k and the RNA flag
            txt = al.format_segments(res, "read-%d" % i, "sig-%d" % i, 1234, 99999, str(z["fmt%d/read" % i]))
            assert txt == z["fmt%d/expected" % i].tobytes()


def test_emulated_preprocess_and_format():
    import build_emu
    lib = build_emu.build()
    _check_preprocess(_aligner(lib))
    _check_format({True: _aligner(lib, "rna002"), False: _aligner(lib, "dna_r9")})


@pytest.mark.gpu
def test_gpu_preprocess_and_format():
    _check_preprocess(_aligner())
    _check_format({True: _aligner(None, "rna002"), False: _aligner(None, "dna_r9")})
    # fused use: preprocess on the GPU, then align what comes out
    al = _aligner()
    from dynamont_b200.synth import native_model, synth_read
    from conftest import MODELS_DIR as M
    nm, ns = native_model(os.path.join(M, "rna002_5mer.model"), "rna002")
    s, q, _ = synth_read(np.random.default_rng(9), nm, ns, 5, 200, 10)
    raw = (s * 11.25 + 88.5).astype(np.float32)
    raw[[50, 400, 900]] += 150.0  # sensor spikes
    clean = al.preprocess_batch([raw], [88.5], [11.25])[0]
    r = al.align(clean, q, True)
    assert len(r["signal_positions"]) == len(q) - 4 and al.last_timing()["log2_fallback_reads"] == 0
