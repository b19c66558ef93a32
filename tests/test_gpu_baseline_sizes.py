"""GPU parity at the sizes BASELINE.json quotes, against the compiled, unmodified reference run LIVE on the same box
(oracle/_ref/libdynamont_ref.so; the bit-exact C restatement if that library did not travel):

  c1  eight 1 kb rna002 5-mer reads at 30 samples/base (config 1)
  c2  rna004 9-mer reads of 0.5, 2 and 5 kb at 30 samples/base (config 2; the 5 kb read is ~13 s / 3.9 GB of reference)
  c4  one 10 kb Gamma-dwell read at 40 samples/base (the long-read recipe of config 4 at the size the reference can hold)
  an un-rounded float64 signal (the front end's (x - shift) / scale), and training on a c2-sized read.

Gates (north_star): borders >= 99.9 % identical, |dp| <= 1e-4, |dZ| <= 1e-6 |Z|, trained mean AND stdev <= 1e-4 relative."""
import numpy as np
import pytest

from conftest import TRAIN_RTOL, check_alignment

pytestmark = pytest.mark.gpu


def _checker(path, pore):
    import oracle
    return oracle.Reference(path, pore) if oracle.have_reference() else oracle.Oracle(path, pore)


def _reads(path, pore, spec, seed, dwell="geometric", f32=True):
    from dynamont_b200.synth import PORE_INFO, native_model, synth_read
    nm, ns = native_model(path, pore)
    k = PORE_INFO[pore][1]
    rng = np.random.default_rng(seed)
    out = []
    for L, spb in spec:
        s, q, _ = synth_read(rng, nm, ns, k, L, spb, dwell=dwell)
        out.append((s.astype(np.float32) if f32 else s, q))
    return out


def _compare(al, ref, reads, what):
    res = al.align_batch([s for s, _ in reads], [q for _, q in reads], True, raise_errors=True)
    n_seg = n_same = 0
    worst = 0.0
    for (s, q), r in zip(reads, res):
        o = ref.align(np.asarray(s, dtype=np.float64), q, True)
        frac, dp = check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"], what)
        n_seg += o["signal_positions"].size
        n_same += int((r["signal_positions"] == o["signal_positions"]).sum())
        worst = max(worst, dp)
    assert n_same >= 0.999 * n_seg
    return n_same, n_seg, worst


def test_c1_eight_1kb_reads_vs_reference(models_dir):
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    path = materialize_model("rna002_5mer", models_dir)
    al = Aligner(path, "rna002")
    reads = _reads(path, "rna002", [(1000, 30)] * 8, seed=20261000)
    same, seg, dp = _compare(al, _checker(path, "rna002"), reads, "c1")
    tm = al.last_timing()
    assert tm["ribbon_reads"] == 8 and tm["log2_fallback_reads"] == 0
    print("c1: %d/%d borders identical, max |dp| %.2e, ribbon faults %d" % (same, seg, dp, tm["ribbon_faults"]))


@pytest.mark.parametrize("length", [500, 2000, 5000])
def test_c2_9mer_reads_vs_reference(length, models_dir):
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    path = materialize_model("synthetic_rna004_9mer", models_dir)
    al = Aligner(path, "rna004")
    reads = _reads(path, "rna004", [(length, 30)], seed=20262000 + length)
    same, seg, dp = _compare(al, _checker(path, "rna004"), reads, "c2 %d b" % length)
    print("c2 %d b: %d/%d borders identical, max |dp| %.2e, ribbon %s" % (length, same, seg, dp, al.last_timing()["ribbon_reads"]))


def test_c2_full_band_kernels_vs_reference(models_dir):
    """the same 2 kb read through the full-band kernels alone (ribbon tier switched off): the tier the ribbon hands to"""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    path = materialize_model("synthetic_rna004_9mer", models_dir)
    al = Aligner(path, "rna004")
    al.set_option("ribbon", 0)
    reads = _reads(path, "rna004", [(2000, 30)], seed=20262000 + 2000)
    _compare(al, _checker(path, "rna004"), reads, "c2 full band")
    assert al.last_timing()["ribbon_reads"] == 0


def test_c4_10kb_gamma_dwell_read_vs_reference(models_dir):
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    path = materialize_model("synthetic_rna004_9mer", models_dir)
    al = Aligner(path, "rna004")
    reads = _reads(path, "rna004", [(10000, 40)], seed=20264000, dwell="gamma")
    same, seg, dp = _compare(al, _checker(path, "rna004"), reads, "c4 10 kb")
    print("c4 10 kb: %d/%d borders identical, max |dp| %.2e" % (same, seg, dp))


def test_c4_records_free_layout_vs_reference_and_record_layout(models_dir):
    """config 4's scratch layout (chosen by the engine when the resident warps' record buffers would not fit: two-level
    checkpoints, row header = decision words, path posteriors from a second forward sweep after the traceback), forced
    here on a 10 kb Gamma-dwell read and two c2-sized reads: parity against the reference, and identical to the record layout"""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    path = materialize_model("synthetic_rna004_9mer", models_dir)
    reads = _reads(path, "rna004", [(10000, 40)], seed=20264000, dwell="gamma") + \
        _reads(path, "rna004", [(700, 30), (3000, 30)], seed=20264001)
    al = Aligner(path, "rna004")
    al.set_option("rib_gather", 1)
    same, seg, dp = _compare(al, _checker(path, "rna004"), reads, "c4 records-free")
    tm = al.last_timing()
    assert tm["ribbon_reads"] == 3 and tm["ribbon_faults"] == 0 and al.ribbon_fault_reasons()["records_free_layout"]
    got = al.align_batch([s for s, _ in reads], [q for _, q in reads], True, raise_errors=True)
    al2 = Aligner(path, "rna004")
    al2.set_option("rib_gather", 0)
    ref = al2.align_batch([s for s, _ in reads], [q for _, q in reads], True, raise_errors=True)
    assert not al2.ribbon_fault_reasons()["records_free_layout"]
    for a, b in zip(got, ref):
        assert a["Z"] == b["Z"] and np.array_equal(a["signal_positions"], b["signal_positions"])
        kept = b["probabilities"] > 2.0 ** -15   # the record layout drops posteriors below its 2^-16 record threshold
        assert np.array_equal(a["probabilities"][kept], b["probabilities"][kept])
    print("c4 records-free: %d/%d borders identical, max |dp| %.2e" % (same, seg, dp))


def test_unrounded_float64_signal_vs_reference(models_dir):
    """A float64 signal that is NOT FP32-representable (what (x - shift) / scale produces): the C ABI takes float64
    (dyn_align_batch_f64, as aligner_bindings.cpp:111-147 does) and rounds to FP32 on the way to the device; the
    reference computes on the doubles.  The rounding of the samples (2^-24 relative) must stay inside the gates."""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    path = materialize_model("rna002_5mer", models_dir)
    al = Aligner(path, "rna002")
    (s32, q), = _reads(path, "rna002", [(1000, 30)], seed=20261999)
    rng = np.random.default_rng(5)
    raw = np.round(s32.astype(np.float64) * 137.21 + 721.3)          # integer ADC counts
    shift, scale = 721.3 + 0.377, 137.21 * 1.013                     # basecaller's shift / scale
    sig = (raw - shift) / scale                                      # float64, not FP32-representable
    assert np.any(sig != sig.astype(np.float32).astype(np.float64))
    r = al.align(sig, q, True)
    o = _checker(path, "rna002").align(sig, q, True)
    check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"], "float64 signal")


def test_training_c2_sized_read_vs_reference(models_dir):
    """per-read Baum-Welch re-estimates (NT:462-561, 641-725) on a 2 kb 9-mer read: transitions, means AND stdevs at
    1e-4 relative for every kmer above the weight threshold"""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    path = materialize_model("synthetic_rna004_9mer", models_dir)
    al = Aligner(path, "rna004")
    (s, q), = _reads(path, "rna004", [(2000, 30)], seed=20265000)
    per_read, pooled = al.train_batch([s], [q], per_read_model=True)
    assert al.last_timing()["ribbon_reads"] == 1
    o = _checker(path, "rna004").train(s.astype(np.float64), q)
    r = per_read[0]
    assert abs(r["Z"] - o["Z"]) <= 1e-6 * abs(o["Z"])
    for key in ("m1", "e1", "e2"):
        assert abs(r["transition_params"][key] - o["transition_params"][key]) <= TRAIN_RTOL * o["transition_params"][key]
    heavy = pooled["w"] > 1e-3
    assert heavy.sum() > 1500
    np.testing.assert_allclose(r["emission_model"]["mean"][heavy], o["emission_model"]["mean"][heavy], rtol=TRAIN_RTOL, atol=1e-5)
    np.testing.assert_allclose(r["emission_model"]["stdev"][heavy], o["emission_model"]["stdev"][heavy], rtol=TRAIN_RTOL, atol=1e-6)
