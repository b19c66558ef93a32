"""GPU parity tests proper: the CUDA path, called through the C ABI (dynamont_b200.Aligner -> ctypes ->
libdynamont_b200.so), against (i) the golden vectors produced by the unmodified reference C++, (ii) the CPU
oracle on seeded reads, (iii) size-independent properties at BASELINE.json's read sizes."""
import numpy as np
import pytest

from conftest import PROB_ATOL, TRAIN_RTOL, check_alignment, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def aligners():
    from dynamont_b200 import Aligner
    cache = {}

    def get(model_path, pore, band=400):
        key = (model_path, pore, band)
        if key not in cache:
            cache[key] = Aligner(model_path, pore, band=band)
        return cache[key]
    return get


@pytest.mark.parametrize("case", load_golden(), ids=lambda c: c.name)
def test_align_matches_reference_golden(case, aligners):
    al = aligners(case.model_path, case.pore)
    r = al.align(case.signal, case.sequence, True)
    check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    assert r["states"] == ["M"] * len(r["signal_positions"]) and r["polishes"] == [""] * len(r["states"])
    z_only = al.align(case.signal, case.sequence, False)
    assert z_only["Z"] == r["Z"] and len(z_only["signal_positions"]) == 0


@pytest.mark.parametrize("case", load_golden(), ids=lambda c: c.name)
def test_train_matches_reference_golden(case, aligners):
    al = aligners(case.model_path, case.pore)
    per_read, pooled = al.train_batch([case.signal], [case.sequence], per_read_model=True)
    r = per_read[0]
    assert abs(r["Z"] - case.train_Z) <= 1e-6 * max(1.0, abs(case.train_Z))
    tp = r["transition_params"]
    np.testing.assert_allclose([tp["m1"], tp["e1"], tp["e2"]], case.train_trans, rtol=TRAIN_RTOL)
    km = case.train_kmers
    # raw sufficient statistics (what a data-parallel trainer all-reduces)
    heavy = case.stat_w > 1e-3  # parity is only meaningful above a weight threshold
    np.testing.assert_allclose(pooled["w"][km][heavy], case.stat_w[heavy], rtol=TRAIN_RTOL)
    np.testing.assert_allclose(pooled["x"][km][heavy], case.stat_x[heavy], rtol=TRAIN_RTOL, atol=1e-6)
    np.testing.assert_allclose(pooled["xx"][km][heavy], case.stat_xx[heavy], rtol=TRAIN_RTOL)
    # per-read M-step (reference semantics)
    np.testing.assert_allclose(r["emission_model"]["mean"][km][heavy], case.train_mean[heavy], rtol=TRAIN_RTOL, atol=1e-5)
    np.testing.assert_allclose(r["emission_model"]["stdev"][km][heavy], case.train_stdev[heavy], rtol=TRAIN_RTOL, atol=1e-6)
    # untouched kmers keep the model (NT:531-534)
    mean0, sd0 = al.model()
    untouched = np.ones(al.num_kmers, bool)
    untouched[km] = False
    idx = np.nonzero(untouched & (pooled["w"] == 0))[0][:1000]
    assert np.array_equal(r["emission_model"]["mean"][idx], mean0[idx])


def _synth_batch(model_path, pore, n, lo, hi, spb, seed, dwell="geometric"):
    from dynamont_b200.synth import PORE_INFO, native_model, synth_read
    nm, ns = native_model(model_path, pore)
    k = PORE_INFO[pore][1]
    rng = np.random.default_rng(seed)
    sigs, seqs = [], []
    for _ in range(n):
        L = int(rng.integers(lo, hi + 1))
        s, q, _ = synth_read(rng, nm, ns, k, L, spb, dwell=dwell)
        sigs.append(s.astype(np.float32))
        seqs.append(q)
    return sigs, seqs


def test_batch_vs_oracle_seeded(aligners, models_dir):
    """A ragged batch (incl. reads the reference throws on) against the CPU oracle, read by read."""
    from dynamont_b200.synth import materialize_model
    from oracle import Oracle
    path = materialize_model("rna002_5mer", models_dir)
    al = aligners(path, "rna002")
    orc = Oracle(path, "rna002")
    sigs, seqs = _synth_batch(path, "rna002", 24, 30, 900, 9, seed=20260001)
    sigs += [np.zeros(0, np.float32), sigs[0][:50], sigs[1]]
    seqs += ["ACGTACGTAC", seqs[0], seqs[1][:20] + "N" + seqs[1][21:]]
    res = al.align_batch(sigs, seqs, True)
    n_seg = n_same = 0
    for s, q, r in zip(sigs, seqs, res):
        try:
            o = orc.align(s.astype(np.float64), q, True)
        except RuntimeError as e:
            assert isinstance(r, RuntimeError) and str(r) == str(e)
            continue
        check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"])
        n_seg += o["signal_positions"].size
        n_same += int((r["signal_positions"] == o["signal_positions"]).sum())
    assert n_same >= 0.999 * n_seg


def test_batch_order_and_idempotence(aligners, models_dir):
    from dynamont_b200.synth import materialize_model
    path = materialize_model("rna002_5mer", models_dir)
    al = aligners(path, "rna002")
    sigs, seqs = _synth_batch(path, "rna002", 40, 100, 1200, 12, seed=7)
    a = al.align_batch(sigs, seqs, True)
    perm = np.random.default_rng(0).permutation(len(sigs))
    b = al.align_batch([sigs[i] for i in perm], [seqs[i] for i in perm], True)
    for j, i in enumerate(perm):
        assert a[i]["Z"] == b[j]["Z"]
        assert np.array_equal(a[i]["signal_positions"], b[j]["signal_positions"])
        assert np.array_equal(a[i]["probabilities"], b[j]["probabilities"])


def test_full_size_properties(aligners, models_dir):
    """BASELINE config sizes (1 kb .. 5 kb reads, 30 samples/base, 9-mer model): properties that need no oracle."""
    from dynamont_b200.synth import materialize_model
    path = materialize_model("synthetic_rna004_9mer", models_dir)
    al = aligners(path, "rna004")
    sigs, seqs = _synth_batch(path, "rna004", 6, 1000, 5000, 30, seed=20262000)
    res = al.align_batch(sigs, seqs, True)
    k = al.kmer_size
    for s, q, r in zip(sigs, seqs, res):
        Kc = len(q) - k + 1
        sp = r["signal_positions"].astype(np.int64)
        assert sp.size == Kc and sp[0] == 0                      # exactly Kc segments, first starts at sample 0
        assert np.all(np.diff(sp) >= 2)                          # every kmer emits >= 2 samples (M then E)
        assert sp[-1] <= s.size - 2
        assert np.array_equal(r["sequence_positions"], np.arange(Kc, dtype=np.uint64) + k // 2)
        assert np.all(r["probabilities"] >= 0) and np.all(r["probabilities"] <= 1 + 1e-6)
        assert np.isfinite(r["Z"])
    # synthetic reads are easy: borders land close to the truth and are confident
    assert np.median(np.concatenate([r["probabilities"] for r in res])) > 0.5


def test_band_cutoff_fails_like_reference(aligners, models_dir):
    """A read whose true path leaves the band: Z is -inf and the reference throws 'Alignment failed'."""
    from dynamont_b200.synth import materialize_model
    from oracle import Oracle
    path = materialize_model("rna002_5mer", models_dir)
    al = aligners(path, "rna002", band=6)
    orc = Oracle(path, "rna002", band=6)
    sigs, seqs = _synth_batch(path, "rna002", 3, 200, 300, 10, seed=3)
    res = al.align_batch(sigs, seqs, True)
    for s, q, r in zip(sigs, seqs, res):
        try:
            o = orc.align(s.astype(np.float64), q, True)
            check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"])
        except RuntimeError as e:
            assert isinstance(r, RuntimeError) and str(r) == str(e)


def test_range_fault_reads_fall_back_to_log2_domain(aligners, models_dir):
    """Reads the FP32 linear-domain kernels cannot represent (outlier samples no kmer explains; a band so narrow that
    it clips the alignment) must be detected on the device and come back from the log2-domain kernels — identical to
    the oracle either way, and clean reads of the same batch must not take the fallback."""
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    from oracle import Oracle
    path = materialize_model("rna002_5mer", models_dir)
    al = aligners(path, "rna002")
    orc = Oracle(path, "rna002")
    sigs, seqs = _synth_batch(path, "rna002", 12, 100, 400, 10, seed=77)
    al.align_batch(sigs, seqs, True)
    assert al.last_timing()["log2_fallback_reads"] == 0
    rng = np.random.default_rng(5)
    dirty = []
    for i in (2, 5, 9):
        s = sigs[i].copy()
        pos = rng.integers(20, s.size - 20, size=3)
        s[pos] += rng.choice([-1.0, 1.0], 3).astype(np.float32) * 9.0
        sigs[i] = s
        dirty.append(i)
    res = al.align_batch(sigs, seqs, True)
    # the linear-domain ribbon loses them (FP32 range); the log2-domain ribbon keeps them on the same narrow window, or
    # hands them to the full-band log2-domain kernels: either way off the linear path, and identical to the oracle
    tm, why = al.last_timing(), al.ribbon_fault_reasons()
    assert tm["ribbon_faults"] == len(dirty)
    assert why["kept_by_log2_ribbon"] + tm["log2_fallback_reads"] >= len(dirty)
    for s, q, r in zip(sigs, seqs, res):
        o = orc.align(s.astype(np.float64), q, True)
        check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"])
    # both arithmetic paths agree on clean reads as well
    al2 = Aligner(path, "rna002")
    al2.set_option("arith", 1)
    res2 = al2.align_batch(sigs, seqs, True)
    for a, b in zip(res, res2):
        assert np.array_equal(a["signal_positions"], b["signal_positions"])
        assert np.abs(a["probabilities"] - b["probabilities"]).max() <= PROB_ATOL
    # narrow band: the alignment hugs the band edge
    aln = aligners(path, "rna002", band=10)
    orn = Oracle(path, "rna002", band=10)
    sigs, seqs = _synth_batch(path, "rna002", 4, 120, 200, 8, seed=3)
    res = aln.align_batch(sigs, seqs, True)
    for s, q, r in zip(sigs, seqs, res):
        try:
            o = orn.align(s.astype(np.float64), q, True)
            check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"])
        except RuntimeError as e:
            assert isinstance(r, RuntimeError) and str(r) == str(e)


def test_async_lanes_match_sync_batches(aligners, models_dir):
    """dyn_align_submit / dyn_align_wait: three batches streamed through the two lanes return exactly what the synchronous
    entry point returns, in input order (the lanes share the root handle's ribbon scratch and stream)."""
    from dynamont_b200.synth import materialize_model
    path = materialize_model("rna002_5mer", models_dir)
    al = aligners(path, "rna002")
    batches = [_synth_batch(path, "rna002", 12, 100, 1500, 12, seed=100 + i) for i in range(3)]
    batches[1][0].append(batches[1][0][0][:10])
    batches[1][1].append(batches[1][1][0])
    ref = [al.align_batch(s, q, True) for s, q in batches]
    got = list(al.align_stream(batches, True, depth=2))
    for rb, gb in zip(ref, got):
        assert len(rb) == len(gb)
        for r, g in zip(rb, gb):
            if isinstance(r, Exception):
                assert isinstance(g, Exception) and str(g) == str(r)
                continue
            assert r["Z"] == g["Z"]
            assert np.array_equal(r["signal_positions"], g["signal_positions"])
            assert np.array_equal(r["probabilities"], g["probabilities"])


def test_pooled_trainer_device_statistics(aligners):
    """dyn_train_accumulate / dyn_train_mstep_device (PooledTrainer.iteration_device, world 1): the statistics stay in a
    CUDA tensor, are the reference's per-read expected counts (golden raw statistics of runTraining), and the device
    M-step equals the reference's per-read M-step.  The all-reduce itself is exercised at world 2 by the gloo test and by
    bench.py --config c5 --gpus N."""
    import torch
    from dynamont_b200 import Aligner
    from dynamont_b200.train import PooledTrainer
    case = [c for c in load_golden() if c.name == "rna002_band"][0]
    al = Aligner(case.model_path, case.pore)
    tr = PooledTrainer(al, 0, 1, device=torch.device("cuda", 0))
    trans, st = tr.iteration_device([case.signal.astype(np.float32)], [case.sequence])
    assert st.is_cuda and st.dtype == torch.float64
    K = al.num_kmers
    h = st.cpu().numpy()
    km = case.train_kmers
    heavy = case.stat_w > 1e-3
    np.testing.assert_allclose(h[:K][km][heavy], case.stat_w[heavy], rtol=TRAIN_RTOL)
    np.testing.assert_allclose(h[K:2 * K][km][heavy], case.stat_x[heavy], rtol=TRAIN_RTOL, atol=1e-6)
    np.testing.assert_allclose(h[2 * K:3 * K][km][heavy], case.stat_xx[heavy], rtol=TRAIN_RTOL)
    assert h[3 * K + 3] == 1.0 and abs(h[3 * K + 2] - case.train_Z) <= 1e-6 * abs(case.train_Z)
    np.testing.assert_allclose([trans["m1"], trans["e1"], trans["e2"]], case.train_trans, rtol=TRAIN_RTOL)
    mean, sd = al.model()
    np.testing.assert_allclose(mean[km][heavy], case.train_mean[heavy], rtol=TRAIN_RTOL, atol=1e-5)
    np.testing.assert_allclose(sd[km][heavy], case.train_stdev[heavy], rtol=TRAIN_RTOL, atol=1e-6)


def test_multi_device_front_equals_single(aligners, models_dir):
    """MultiDeviceAligner over every visible GPU (twice the same GPU when only one is visible): reads dealt by
    shard_indices, results merged in input order — exactly what one handle returns"""
    import torch
    from dynamont_b200 import MultiDeviceAligner
    from dynamont_b200.synth import materialize_model
    path = materialize_model("rna002_5mer", models_dir)
    sigs, seqs = _synth_batch(path, "rna002", 31, 60, 1400, 11, seed=4242)
    single = aligners(path, "rna002").align_batch(sigs, seqs, True)
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    multi = MultiDeviceAligner(path, "rna002", devices=devices).align_batch(sigs, seqs, True)
    for a, b in zip(single, multi):
        assert a["Z"] == b["Z"] and np.array_equal(a["signal_positions"], b["signal_positions"])
        assert np.array_equal(a["probabilities"], b["probabilities"])


@pytest.mark.parametrize("case", load_golden(), ids=lambda c: c.name)
def test_full_band_tiers_match_reference_golden(case):
    """the tiers behind the ribbon (round 1's full-band kernels) on their own: ribbon switched off"""
    from dynamont_b200 import Aligner
    al = Aligner(case.model_path, case.pore)
    al.set_option("ribbon", 0)
    r = al.align(case.signal, case.sequence, True)
    assert al.last_timing()["ribbon_reads"] == 0
    check_alignment(r, case.signal_positions, case.sequence_positions, case.probabilities, case.Z, case.name)
    per_read, pooled = al.train_batch([case.signal], [case.sequence], per_read_model=True)
    km = case.train_kmers
    heavy = case.stat_w > 1e-3
    np.testing.assert_allclose(pooled["w"][km][heavy], case.stat_w[heavy], rtol=TRAIN_RTOL)
    np.testing.assert_allclose(per_read[0]["emission_model"]["stdev"][km][heavy], case.train_stdev[heavy], rtol=TRAIN_RTOL, atol=1e-6)


def test_noisy_read_that_needs_the_first_row_mass_check(models_dir):
    """tools/gpu_soak.py, kind "noisy", read 8 (tests/golden/noisy_read8_*): 2.5x noise loses 2^-128 of backward range over
    one 16-row group; before the posterior mass was checked on the first row of a group as well, the linear-domain ribbon
    returned it with one segment posterior of 2^-10 instead of 1 and no fault"""
    import os
    from conftest import ROOT
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    from oracle import Oracle
    path = materialize_model("rna002_5mer", models_dir)
    s = np.load(os.path.join(ROOT, "tests", "golden", "noisy_read8_signal.npy"))
    q = open(os.path.join(ROOT, "tests", "golden", "noisy_read8_sequence.txt")).read().strip()
    o = Oracle(path, "rna002").align(s.astype(np.float64), q, True)
    for rib_log in (1, 0):
        al = Aligner(path, "rna002")
        al.set_option("rib_log", rib_log)
        r = al.align(s, q, True)
        check_alignment(r, o["signal_positions"], o["sequence_positions"], o["probabilities"], o["Z"])
        assert al.last_timing()["ribbon_faults"] == 1
