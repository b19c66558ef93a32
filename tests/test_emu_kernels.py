"""CPU tests of the CUDA kernels through the SIMT emulator (tests/emu): the very same .cu sources compiled with
g++ against simt_host.h, driven through the same C ABI and Python binding as the GPU library.  They cover the
kernel logic (ring slots, band slides, checkpoints, per-lane offsets, traceback, medians, training statistics)
against the golden vectors of the compiled reference.  The numerical claims that count are the -m gpu tests."""
import os
import sys

import numpy as np
import pytest

from conftest import TRAIN_RTOL, check_alignment, load_golden

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
import build_emu  # noqa: E402

SMALL = ["rna002_short", "rna002_band", "rna002_min_dwell", "rna002_dinuc", "rna004_9mer"]


@pytest.fixture(scope="module")
def emu_lib():
    return build_emu.build()


def _aligner(lib, case, variant=0, band=400, arith=0, ribbon=None):
    """variant >= 0: that build variant of the FULL-BAND kernels with the ribbon tier switched off (so that the tiers behind
    the ribbon keep their coverage); variant -1: the library default (ribbon tier first)."""
    from dynamont_b200 import Aligner
    al = Aligner(case.model_path, case.pore, band=band, _lib_path=lib)
    al.set_option("variant", variant)
    al.set_option("arith", arith)  # 0: linear-domain kernels (+ log2-domain fallback), 1: log2-domain kernels only
    if ribbon is None:
        ribbon = 2 if variant < 0 else 0
    al.set_option("ribbon", ribbon)
    return al


@pytest.mark.parametrize("tier", ["ribbon", "ribbon4", "full_band"])
@pytest.mark.parametrize("case", [c for c in load_golden() if c.name in SMALL], ids=lambda c: c.name)
def test_emulated_align_matches_reference(case, tier, emu_lib):
    al = _aligner(emu_lib, case, ribbon={"ribbon": 2, "ribbon4": 4, "full_band": 0}[tier])
    r = al.align(case.signal, case.sequence, True)
    assert (al.last_timing()["ribbon_reads"] == 1) == (tier != "full_band")
    check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    assert al.align(case.signal, case.sequence, False)["Z"] == r["Z"]


@pytest.mark.parametrize("case", [c for c in load_golden() if c.name in ("rna002_short", "rna002_dinuc")], ids=lambda c: c.name)
def test_emulated_log2_domain_kernels(case, emu_lib):
    """the log2-domain kernels (the fallback of the linear-domain path) on their own"""
    al = _aligner(emu_lib, case, arith=1)
    r = al.align(case.signal, case.sequence, True)
    check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    assert al.last_timing()["log2_fallback_reads"] == 0


def test_emulated_range_fault_falls_back(emu_lib):
    """Outlier samples that no kmer explains underflow FP32 probabilities: the linear-domain kernel must notice
    (ST_LIN_FAULT) and the read must come back from the log2-domain kernels, identical to the oracle."""
    from dynamont_b200.synth import native_model, synth_read
    from oracle import Oracle
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    nm, ns = native_model(case.model_path, case.pore)
    rng = np.random.default_rng(11)
    s, q, _ = synth_read(rng, nm, ns, 5, 120, 9)
    s[200] += 9.0
    s[201] -= 8.0
    s = s.astype(np.float32).astype(np.float64)
    o = Oracle(case.model_path, case.pore).align(s, q, True)
    al = _aligner(emu_lib, case)
    clean = al.align(case.signal, case.sequence, True)
    assert al.last_timing()["log2_fallback_reads"] == 0
    check_alignment(clean, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    res = al.align_batch([case.signal, s], [case.sequence, q], True)
    assert al.last_timing()["log2_fallback_reads"] == 1
    check_alignment(res[0], case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    check_alignment(res[1], o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"], "spikes")
    # training takes the same route and must not double-count the read's statistics
    _, pooled_lin = al.train_batch([s], [q])
    assert al.last_timing()["log2_fallback_reads"] == 1
    _, pooled_log = _aligner(emu_lib, case, arith=1).train_batch([s], [q])
    np.testing.assert_allclose(pooled_lin["w"], pooled_log["w"], rtol=1e-12)


def test_emulated_log2_ribbon_keeps_what_the_linear_ribbon_loses(emu_lib):
    """the log2-domain ribbon (tier 0b): (1) forced on every golden case it must reproduce the reference; (2) a read with
    samples no kmer explains and a read whose band (16 columns) cuts its alignment underflow the FP32 products of the
    linear-domain ribbon — the log2-domain ribbon keeps them on the narrow window (no full-band launch), identical to
    the oracle; (3) with the tier switched off the same reads come back from the full-band log2 kernels, same results"""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import native_model, synth_read
    from oracle import Oracle
    for case in [c for c in load_golden() if c.name in SMALL and c.name != "rna002_dinuc"]:
        al = _aligner(emu_lib, case, -1)
        al.set_option("rib_log", 2)
        r = al.align(case.signal, case.sequence, True)
        assert al.ribbon_fault_reasons()["kept_by_log2_ribbon"] == 1 and al.last_timing()["log2_fallback_reads"] == 0
        check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    nm, ns = native_model(case.model_path, case.pore)
    s, q, _ = synth_read(np.random.default_rng(11), nm, ns, 5, 120, 9)
    s[200] += 9.0
    s[201] -= 8.0
    s = s.astype(np.float32).astype(np.float64)
    o = Oracle(case.model_path, case.pore).align(s, q, True)
    got = {}
    for rib_log in (1, 0):
        al = _aligner(emu_lib, case, -1)
        al.set_option("rib_log", rib_log)
        r = al.align(s, q, True)
        tm, why = al.last_timing(), al.ribbon_fault_reasons()
        assert tm["ribbon_faults"] == 1
        assert (why["kept_by_log2_ribbon"], tm["log2_fallback_reads"]) == ((1, 0) if rib_log else (0, 1)), (why, tm)
        check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"], "spikes rib_log=%d" % rib_log)
        got[rib_log] = r
    assert np.array_equal(got[0]["signal_positions"], got[1]["signal_positions"])
    # a band that cuts the alignment: whatever the reference does (forced alignment or "scores do not match"), so do we
    kept = 0
    for band, seed in ((16, 16), (16, 3), (24, 5), (12, 8)):
        aln = Aligner(case.model_path, case.pore, band=band, _lib_path=emu_lib)
        orn = Oracle(case.model_path, case.pore, band=band)
        s, q, _ = synth_read(np.random.default_rng(seed), nm, ns, 5, 150, 8)
        try:
            o = orn.align(s, q, True)
        except RuntimeError as e:
            with pytest.raises(RuntimeError, match=str(e)[:20]):
                aln.align(s, q, True)
            continue
        check_alignment(aln.align(s, q, True), o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"], "band %d" % band)
        kept += aln.ribbon_fault_reasons()["kept_by_log2_ribbon"]
    print("band-cut reads kept by the log2-domain ribbon:", kept)


def test_emulated_noisy_read_16_row_groups(emu_lib):
    """the read that slipped through the 16-row ribbon before the posterior mass was checked on the FIRST row of a group as
    well (2.5x noise: the backward values lose 2^-128 over one group; one segment came back with a posterior of 2^-10
    instead of 1, borders identical, no fault — tools/gpu_soak.py, kind "noisy", read 8): it must fault in the
    linear-domain ribbon and come back right from the log2-domain ribbon, and from the full-band tiers without it"""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    from conftest import MODELS_DIR, ROOT
    from oracle import Oracle
    path = materialize_model("rna002_5mer", MODELS_DIR)
    s = np.load(os.path.join(ROOT, "tests", "golden", "noisy_read8_signal.npy"))
    q = open(os.path.join(ROOT, "tests", "golden", "noisy_read8_sequence.txt")).read().strip()
    o = Oracle(path, "rna002").align(s.astype(np.float64), q, True)
    for rib_log in (1, 0):
        al = Aligner(path, "rna002", _lib_path=emu_lib)
        al.set_option("rib_log", rib_log)
        r = al.align(s, q, True)
        check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"], "noisy read 8, rib_log=%d" % rib_log)
        assert al.last_timing()["ribbon_faults"] == 1 and al.ribbon_fault_reasons().get(5) == 1
    # training takes the same route: the linear-domain ribbon faults, the log2-domain ribbon (MODE 2) returns the statistics
    ot = Oracle(path, "rna002").train(s.astype(np.float64), q)
    for rib_log in (1, 0):
        al = Aligner(path, "rna002", _lib_path=emu_lib)
        al.set_option("rib_log", rib_log)
        per, pooled = al.train_batch([s], [q], per_read_model=True)
        r = per[0]
        tm, why = al.last_timing(), al.ribbon_fault_reasons()
        assert tm["ribbon_faults"] == 1 and why["kept_by_log2_ribbon"] == (1 if rib_log else 0) and tm["log2_fallback_reads"] == 0
        assert abs(r["Z"] - ot["Z"]) <= 1e-6 * abs(ot["Z"])
        for key in ("m1", "e1", "e2"):
            assert abs(r["transition_params"][key] - ot["transition_params"][key]) <= TRAIN_RTOL * ot["transition_params"][key]
        heavy = pooled["w"] > 1e-3
        np.testing.assert_allclose(r["emission_model"]["mean"][heavy], ot["emission_model"]["mean"][heavy], rtol=TRAIN_RTOL, atol=1e-5)
        np.testing.assert_allclose(r["emission_model"]["stdev"][heavy], ot["emission_model"]["stdev"][heavy], rtol=TRAIN_RTOL, atol=1e-6)


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 6, 8, 9, 10, 11, 12, 13, -1])
def test_emulated_variants(variant, emu_lib):
    """build variants: general kernels (0-3, 9, 11, 13), uniform-sigma kernels (4-8, 10, 12), library default (-1)"""
    case = [c for c in load_golden() if c.name == "rna002_band"][0]
    al = _aligner(emu_lib, case, variant)
    r = al.align(case.signal, case.sequence, True)
    check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    assert al.last_timing()["log2_fallback_reads"] == 0


@pytest.mark.parametrize("variant", [3, 4])
def test_emulated_generic_forward_rows(variant, emu_lib):
    """pass 2 with the branch-free row body switched off must give the same alignment as with it"""
    case = [c for c in load_golden() if c.name == "rna002_dinuc"][0]
    al = _aligner(emu_lib, case, variant)
    al.set_option("fwd_fast", 0)
    r = al.align(case.signal, case.sequence, True)
    check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    al.set_option("fwd_fast", 1)
    r1 = al.align(case.signal, case.sequence, True)
    assert np.array_equal(r["signal_positions"], r1["signal_positions"])
    np.testing.assert_allclose(r["probabilities"], r1["probabilities"], atol=2e-6)


def test_emulated_nonuniform_model_uses_general_kernels(emu_lib):
    """a trained model (per-kmer sigma) asked for a uniform-sigma variant runs the general kernels"""
    case = [c for c in load_golden() if c.name == "rna002_trained"][0]
    al = _aligner(emu_lib, case, 6)
    r = al.align(case.signal, case.sequence, True)
    check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)


def test_emulated_uniform_training_and_narrow_band(emu_lib):
    """uniform-sigma kernels: training statistics and a band narrow enough to slide every other row"""
    from dynamont_b200 import Aligner
    from oracle import Oracle
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    al = _aligner(emu_lib, case, 5)
    per_read, pooled = al.train_batch([case.signal], [case.sequence], per_read_model=True)
    tp = per_read[0]["transition_params"]
    np.testing.assert_allclose([tp["m1"], tp["e1"], tp["e2"]], case.train_trans, rtol=TRAIN_RTOL)
    heavy = case.stat_w > 1e-3
    np.testing.assert_allclose(pooled["w"][case.train_kmers][heavy], case.stat_w[heavy], rtol=TRAIN_RTOL)
    case = [c for c in load_golden() if c.name == "rna002_min_dwell"][0]
    al = Aligner(case.model_path, case.pore, band=40, _lib_path=emu_lib)
    al.set_option("variant", 4)
    o = Oracle(case.model_path, case.pore, band=40).align(case.signal, case.sequence, True)
    check_alignment(al.align(case.signal, case.sequence, True), o["signal_positions"], o["sequence_positions"],
                    o["probabilities"], o["Z"])


def test_emulated_training(emu_lib):
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    al = _aligner(emu_lib, case)
    per_read, pooled = al.train_batch([case.signal], [case.sequence], per_read_model=True)
    r = per_read[0]
    tp = r["transition_params"]
    np.testing.assert_allclose([tp["m1"], tp["e1"], tp["e2"]], case.train_trans, rtol=TRAIN_RTOL)
    heavy = case.stat_w > 1e-3
    km = case.train_kmers
    np.testing.assert_allclose(pooled["w"][km][heavy], case.stat_w[heavy], rtol=TRAIN_RTOL)
    np.testing.assert_allclose(r["emission_model"]["mean"][km][heavy], case.train_mean[heavy], rtol=TRAIN_RTOL, atol=1e-5)


def test_emulated_narrow_band_and_errors(emu_lib):
    """Narrow bands exercise the band-edge gating; error reads must carry the reference's messages."""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model, native_model, synth_read
    from conftest import MODELS_DIR
    from oracle import Oracle
    path = materialize_model("rna002_5mer", MODELS_DIR)
    nm, ns = native_model(path, "rna002")
    rng = np.random.default_rng(3)
    s, q, _ = synth_read(rng, nm, ns, 5, 150, 8)
    for band in (6, 10):
        al = Aligner(path, "rna002", band=band, _lib_path=emu_lib)
        orc = Oracle(path, "rna002", band=band)
        try:
            o = orc.align(s, q, True)
        except RuntimeError as e:
            with pytest.raises(RuntimeError, match=str(e)):
                al.align(s, q, True)
            continue
        check_alignment(al.align(s, q, True), o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"])
        # a band this narrow clips the alignment: forward and backward ridges separate by hundreds of bits, which the
        # linear-domain kernels must detect and hand to a log2-domain tier (the log2-domain ribbon, which holds the whole
        # band in its window here, or the full-band log2-domain kernels)
        tm, why = al.last_timing(), al.ribbon_fault_reasons()
        assert tm["ribbon_faults"] == 1 and why["kept_by_log2_ribbon"] + tm["log2_fallback_reads"] == 1
    al = Aligner(path, "rna002", _lib_path=emu_lib)
    bad = [(np.zeros(0), "ACGTACGT", "Signal is empty"), (s[:50], "ACG", "Sequence shorter than model kmer size"),
           (s[:20], q[:40], "Signal too short compared to sequence"), (s, q[:30] + "N" + q[31:], "Invalid nucleotide: N")]
    res = al.align_batch([b[0] for b in bad] + [s], [b[1] for b in bad] + [q], True)
    for (sig, seq, msg), r in zip(bad, res):
        assert isinstance(r, RuntimeError) and str(r) == msg
    assert isinstance(res[-1], dict)
    with pytest.raises(ValueError, match="Unknown pore type: foo"):
        Aligner(path, "foo", _lib_path=emu_lib)
    with pytest.raises(ValueError, match="Unknown aligner mode: bar"):
        Aligner(path, "rna002", mode="bar", _lib_path=emu_lib)
    with pytest.raises(RuntimeError, match="Could not open model file"):
        Aligner(path + ".missing", "rna002", _lib_path=emu_lib)
    with pytest.raises(RuntimeError, match="Inconsistent kmer size in model"):
        Aligner(path, "rna004", _lib_path=emu_lib)


def test_emulated_training_record_overflow_retry(emu_lib):
    """a read whose sparse posterior records overflow the per-warp buffer is re-run alone with a full buffer; the pooled
    statistics and the per-read model must come out as without the overflow"""
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    al = _aligner(emu_lib, case, -1)
    per0, pooled0 = al.train_batch([case.signal, case.signal], [case.sequence, case.sequence], per_read_model=True)
    al.set_option("recs_per_row", 0.05)
    per1, pooled1 = al.train_batch([case.signal, case.signal], [case.sequence, case.sequence], per_read_model=True)
    for k in ("w", "x", "xx"):
        np.testing.assert_allclose(pooled1[k], pooled0[k], rtol=1e-12, atol=0)
    for a, b in zip(per0, per1):
        assert a["Z"] == b["Z"]
        np.testing.assert_allclose(a["emission_model"]["mean"], b["emission_model"]["mean"], rtol=1e-12)  # atomics: order
    # the alignment path has had this retry since the first kernels
    r0 = al.align(case.signal, case.sequence, True)
    check_alignment(r0, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)



def test_emulated_async_lanes_match_sync_batches(emu_lib):
    """dyn_align_submit / dyn_align_wait (two lanes sharing the root handle's ribbon scratch): same results, same order
    as the synchronous entry point, errors included"""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import native_model, synth_read
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    nm, ns = native_model(case.model_path, case.pore)
    rng = np.random.default_rng(3)
    batches = []
    for b in range(3):
        sigs, seqs = [], []
        for L in (130, 60, 150):
            s, q, _ = synth_read(rng, nm, ns, 5, L, 9)
            sigs.append(s.astype(np.float32))
            seqs.append(q)
        if b == 1:
            sigs.append(sigs[0][:10])
            seqs.append(seqs[0])
        batches.append((sigs, seqs))
    al = Aligner(case.model_path, case.pore, _lib_path=emu_lib)
    ref = [al.align_batch(s, q, True) for s, q in batches]
    got = list(al.align_stream(batches, True, depth=2))
    assert len(got) == len(ref)
    for rb, gb in zip(ref, got):
        assert len(rb) == len(gb)
        for r, g in zip(rb, gb):
            if isinstance(r, Exception):
                assert isinstance(g, Exception) and str(g) == str(r)
                continue
            assert r["Z"] == g["Z"]
            assert np.array_equal(r["signal_positions"], g["signal_positions"])
            assert np.array_equal(r["probabilities"], g["probabilities"])


def test_emulated_ribbon_two_level_checkpoints_identical(emu_lib):
    """ribbon kernels: checkpoints of every 8th group only + replay into the per-warp ring (long reads) must give
    bit-identical alignments and training statistics to a checkpoint per group; both identical to the golden reference"""
    case = [c for c in load_golden() if c.name == "rna002_band"][0]
    out = {}
    for tl in (0, 1):
        al = _aligner(emu_lib, case, -1)
        al.set_option("rib_two_level", tl)
        r = al.align(case.signal, case.sequence, True)
        assert al.last_timing()["ribbon_reads"] == 1 and al.last_timing()["ribbon_faults"] == 0
        check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
        _, pooled = al.train_batch([case.signal], [case.sequence])
        out[tl] = (r, pooled)
    assert out[0][0]["Z"] == out[1][0]["Z"]
    assert np.array_equal(out[0][0]["signal_positions"], out[1][0]["signal_positions"])
    assert np.array_equal(out[0][0]["probabilities"], out[1][0]["probabilities"])
    assert np.array_equal(out[0][1]["w"], out[1][1]["w"]) and np.array_equal(out[0][1]["xx"], out[1][1]["xx"])


def test_emulated_ribbon_records_free_layout_identical(emu_lib):
    """ribbon kernels, records-free scratch (long reads: row header = decision words, no posterior records, the path
    posteriors from a second forward sweep after the traceback): same segments, same Z, and the same posteriors as the
    record layout — bit for bit where the record layout kept the cell (posterior above its 2^-16 record threshold)"""
    for name in ("rna002_band", "rna002_short", "rna004_9mer"):
        cases = [c for c in load_golden() if c.name == name]
        if not cases:
            continue
        case = cases[0]
        out = {}
        for ga in (0, 1):
            al = _aligner(emu_lib, case, -1)
            al.set_option("rib_gather", ga)
            r = al.align(case.signal, case.sequence, True)
            tm = al.last_timing()
            assert tm["ribbon_reads"] == 1 and tm["ribbon_faults"] == 0
            assert al.ribbon_fault_reasons()["records_free_layout"] == bool(ga)
            check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
            out[ga] = r
        assert out[0]["Z"] == out[1]["Z"]
        assert np.array_equal(out[0]["signal_positions"], out[1]["signal_positions"])
        kept = out[0]["probabilities"] > 2.0 ** -15
        assert np.array_equal(out[0]["probabilities"][kept], out[1]["probabilities"][kept])
        np.testing.assert_allclose(out[0]["probabilities"], out[1]["probabilities"], atol=2.0 ** -15)


def test_emulated_ribbon_clips_the_reference_band_exactly(emu_lib):
    """the ribbon kernels clip every row to the reference band (NT:96-106) when the window reaches beyond it: a read whose
    band is narrower than the window, a read whose alignment drifts to the band edge (3 samples per base), and a band so
    narrow that it cuts the alignment (the reference throws; the ribbon faults and the full-band kernels report it)"""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import native_model, synth_read
    from oracle import Oracle
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    al = _aligner(emu_lib, case, -1)
    r = al.align(case.signal, case.sequence, True)
    tm = al.last_timing()
    assert tm["ribbon_reads"] == 1 and tm["ribbon_faults"] == 0
    check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    nm, ns = native_model(case.model_path, case.pore)
    s, q, _ = synth_read(np.random.default_rng(5), nm, ns, 5, 200, 3)
    o = Oracle(case.model_path, case.pore).align(s, q, True)
    r = al.align(s, q, True)
    assert al.last_timing()["ribbon_reads"] == 1
    check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"], "drifting read")
    for band in (40, 16, 6):
        aln = Aligner(case.model_path, case.pore, band=band, _lib_path=emu_lib)
        orn = Oracle(case.model_path, case.pore, band=band)
        rng = np.random.default_rng(band)
        for L in (90, 150):
            s, q, _ = synth_read(rng, nm, ns, 5, L, 8)
            try:
                o = orn.align(s, q, True)
            except RuntimeError as e:
                with pytest.raises(RuntimeError, match=str(e)[:20]):
                    aln.align(s, q, True)
                continue
            check_alignment(aln.align(s, q, True), o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"],
                            "band %d" % band)


def test_emulated_multi_device_front_equals_single(emu_lib):
    """MultiDeviceAligner (reads dealt over a device list, merged in input order): exactly the single-handle results"""
    from dynamont_b200 import MultiDeviceAligner
    from dynamont_b200.synth import native_model, synth_read
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    nm, ns = native_model(case.model_path, case.pore)
    rng = np.random.default_rng(9)
    sigs, seqs = [], []
    for L in (140, 40, 90, 120, 60, 75, 130):
        s, q, _ = synth_read(rng, nm, ns, 5, L, 8)
        sigs.append(s.astype(np.float32))
        seqs.append(q)
    sigs.append(sigs[0][:12])
    seqs.append(seqs[0])
    single = _aligner(emu_lib, case, -1).align_batch(sigs, seqs, True)
    multi = MultiDeviceAligner(case.model_path, case.pore, devices=[0, 0, 0], _lib_path=emu_lib).align_batch(sigs, seqs, True)
    assert len(single) == len(multi)
    for a, b in zip(single, multi):
        if isinstance(a, Exception):
            assert isinstance(b, Exception) and str(a) == str(b)
            continue
        assert a["Z"] == b["Z"] and np.array_equal(a["signal_positions"], b["signal_positions"])
        assert np.array_equal(a["probabilities"], b["probabilities"])
