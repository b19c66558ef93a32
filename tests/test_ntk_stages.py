"""Resquiggle ("NTK") mode, pre-pass stages.

Golden vectors: tests/golden/ntk_golden.npz (tools/make_golden_ntk.py) — stage outputs of the UNMODIFIED reference C++
for the part of this mode that works as shipped (dense TN / TK pre-passes, row masks, sparse-lattice keys), and the
end-to-end alignment of the reference with the two-line repair of SURVEY.md F2.  CPU tests pin the golden file against the compiled reference where it is available; the GPU test
compares the CUDA stage kernels with the golden vectors through the C ABI (dyn_ntk_prepass)."""
import os

import numpy as np
import pytest

from conftest import MODELS_DIR, ROOT

GOLDEN_NTK = os.path.join(ROOT, "tests", "golden", "ntk_golden.npz")
# one 9-mer read at T = 865 (`python tools/make_golden_ntk.py big`: 8.5 min and 10 GB of the reference per run)
GOLDEN_NTK_BIG = os.path.join(ROOT, "tests", "golden", "ntk_golden_9mer_T1000.npz")


class NtkCase:
    def __init__(self, z, name):
        g = lambda k: z[name + "/" + k]  # noqa: E731
        self.name = name
        self.pore, self.model_name = str(g("pore")), str(g("model"))
        self.signal = g("signal").astype(np.float64)
        self.sequence = str(g("sequence"))
        self.tn_ptr, self.tn_idx, self.tk_ptr, self.tk_idx = g("tn_ptr"), g("tn_idx"), g("tk_ptr"), g("tk_idx")
        self.keys, self.Z, self.transitions = g("keys"), g("Z"), g("transitions")
        self.has_alignment = (name + "/align_Z") in z.files
        if self.has_alignment:
            self.align_Z = float(g("align_Z"))
            self.signal_positions, self.states, self.polishes = g("signal_positions"), g("states"), g("polishes")
            self.sequence_positions, self.probabilities = g("sequence_positions"), g("probabilities")

    @property
    def model_path(self):
        from dynamont_b200.synth import materialize_model
        return materialize_model(self.model_name, MODELS_DIR)

    def mask(self, which, C):
        ptr, idx = (self.tn_ptr, self.tn_idx) if which == "tn" else (self.tk_ptr, self.tk_idx)
        m = np.zeros((ptr.size - 1, C), dtype=bool)
        rows = np.repeat(np.arange(ptr.size - 1), np.diff(ptr))
        m[rows, idx] = True
        return m


def load_ntk():
    out = []
    for path in (GOLDEN_NTK, GOLDEN_NTK_BIG):
        if os.path.exists(path):
            with np.load(path) as z:
                out += [NtkCase(z, str(n)) for n in z["names"]]
    return out


def test_golden_file_shape():
    cases = load_ntk()
    assert len(cases) >= 4
    for c in cases:
        T = c.signal.size + 1
        assert c.tn_ptr.size == T + 1 and c.tk_ptr.size == T + 1
        assert np.all(np.diff(c.keys.astype(np.int64)) > 0)           # sorted, unique (NTK:438-440)
        assert np.all(np.diff(c.tn_ptr) >= 1) and np.all(np.diff(c.tk_ptr) >= 1)  # every row keeps at least one column
        assert c.keys[0] == 0                                          # the seed (0, 0, 0) (NTK:421-425)
        assert abs(c.Z[0] - c.Z[1]) < 1e-6 and abs(c.Z[2] - c.Z[3]) < 1e-6
        if c.has_alignment:
            assert set(c.states.tolist()) <= {"M", "P"} and len(c.polishes) == len(c.states)


@pytest.mark.parametrize("case", load_ntk(), ids=lambda c: c.name)
def test_golden_pinned_against_reference(case):
    """Regenerate the stage outputs with the compiled reference (dev container only) and compare bit for bit."""
    from oracle import Reference, have_reference
    if not have_reference():
        pytest.skip("reference library not available")
    if case.signal.size > 700 or (case.pore in ("rna004", "dna_r10_260bps", "dna_r10_400bps") and case.signal.size > 100):
        pytest.skip("kept short (a 9-mer read costs the reference 25 s per 150 samples): pinned when tools/make_golden_ntk.py wrote it")
    ref = Reference(case.model_path, case.pore, mode="resquiggle")
    st = ref.ntk_prepass(case.signal, case.sequence)
    assert np.array_equal(st["keys"], case.keys)
    assert np.array_equal(st["tn"], case.mask("tn", st["tn"].shape[1]))
    assert np.array_equal(st["tk"], case.mask("tk", st["tk"].shape[1]))
    assert np.array_equal(st["Z"], case.Z)
    # the unmodified reference throws for every input in this mode ...
    with pytest.raises(RuntimeError, match="NTK alignment failed"):
        ref.align(case.signal, case.sequence, True)
    # ... the repaired one reproduces the stored alignment
    if case.has_alignment:
        fix = Reference(case.model_path, case.pore, mode="resquiggle", ntk_fix=True)
        a = fix.align(case.signal, case.sequence, True)
        assert a["Z"] == case.align_Z and np.array_equal(a["signal_positions"], case.signal_positions)
        assert a["polishes"] == case.polishes.tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("case", load_ntk(), ids=lambda c: c.name)
def test_gpu_prepass_matches_reference(case):
    """CUDA TN / TK pre-passes, row masks and keys (through the C ABI) against the reference's stage outputs.
    The CUDA pre-passes run in the FP64 LINEAR domain with per-row scaling (ntk_prepass.cuh) against the reference's FP64
    log space: Z agrees to ~1e-12 relative and a mask decision can only move when a cumulative mass lands on the
    threshold within rounding."""
    from dynamont_b200 import Aligner
    al = Aligner(case.model_path, case.pore, mode="resquiggle")
    np.testing.assert_allclose([al.ntk_transitions()[k] for k in ("a1", "a2", "p1", "p2", "p3", "s1", "s2", "s3", "e1", "e2", "e3",
                                                                  "e4", "i1", "i2", "tn_m", "tn_e", "tk_m", "tk_e")],
                               case.transitions, rtol=0, atol=0)
    r = al.ntk_prepass(case.signal, case.sequence)
    np.testing.assert_allclose(r["Z"], case.Z, rtol=1e-10, atol=1e-9)
    tn, tk = case.mask("tn", r["tn"].shape[1]), case.mask("tk", r["tk"].shape[1])
    rows_tn = (r["tn"] == tn).all(1).mean()
    rows_tk = (r["tk"] == tk).all(1).mean()
    assert rows_tn >= 0.999 and rows_tk >= 0.999, (rows_tn, rows_tk)
    inter = np.intersect1d(r["keys"], case.keys).size
    assert inter >= 0.999 * max(r["keys"].size, case.keys.size)
    if rows_tn == 1.0 and rows_tk == 1.0:
        assert np.array_equal(r["keys"], case.keys)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in load_ntk() if c.has_alignment], ids=lambda c: c.name)
def test_gpu_ntk_alignment_matches_repaired_reference(case):
    """End to end: pre-passes + sparse 5-state forward / backward / MAP / traceback on the GPU (dyn_ntk_align) against
    the reference with the two-line repair of logF / logB: same segments, states and polish kmers;
    Z to 1e-9 relative; probabilities to 1e-4 (north_star tolerance)."""
    from dynamont_b200 import Aligner
    al = Aligner(case.model_path, case.pore, mode="resquiggle")
    r = al.align(case.signal, case.sequence, True)
    assert abs(r["Z"] - case.align_Z) <= 1e-9 * max(1.0, abs(case.align_Z))
    assert r["states"] == case.states.tolist()
    assert r["polishes"] == case.polishes.tolist()
    assert np.array_equal(r["signal_positions"], case.signal_positions)
    assert np.array_equal(r["sequence_positions"], case.sequence_positions)
    assert np.abs(r["probabilities"] - case.probabilities).max() <= 1e-4
    z = al.align(case.signal, case.sequence, False)
    assert z["Z"] == r["Z"] and len(z["states"]) == 0


@pytest.mark.gpu
def test_gpu_prepass_input_errors():
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    al = Aligner(materialize_model("rna002_5mer", MODELS_DIR), "rna002", mode="ntk")
    x = np.zeros(40, dtype=np.float32)
    for sig, seq, msg in [(x[:0], "ACGTACGT", "Signal is empty"), (x, "ACG", "Sequence shorter than model kmer size"),
                          (x[:6], "ACGTACGTACGT", "Signal too short compared to sequence"), (x, "ACGTNACGTA", "Invalid nucleotide")]:
        with pytest.raises(RuntimeError, match=msg):
            al.ntk_prepass(sig, seq)


@pytest.mark.gpu
def test_gpu_ntk_vs_live_repaired_reference_seeded():
    """Seeded reads (not in the golden file) against the repaired reference run live on the same box."""
    import time
    from oracle import Reference, _build
    if _build.build_reference_ntkfix() is None:
        pytest.skip("oracle/_ref/libdynamont_ref_ntkfix.so not available")
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model, native_model, synth_read
    path = materialize_model("rna004_5mer", MODELS_DIR)
    nm, ns = native_model(path, "dna_r9")
    ref = Reference(path, "dna_r9", mode="resquiggle", ntk_fix=True)
    al = Aligner(path, "dna_r9", mode="resquiggle")
    rng = np.random.default_rng(31)
    t_ref = t_gpu = 0.0
    n_seg = n_same = 0
    for L, spb in ((40, 6), (150, 7), (90, 12)):
        s, q, _ = synth_read(rng, nm, ns, 5, L, spb)
        t0 = time.perf_counter()
        o = ref.align(s, q, True)
        t1 = time.perf_counter()
        r = al.align(s, q, True)
        t2 = time.perf_counter()
        t_ref += t1 - t0
        t_gpu += t2 - t1
        assert abs(r["Z"] - o["Z"]) <= 1e-9 * max(1.0, abs(o["Z"]))
        assert r["states"] == o["states"] and r["polishes"] == o["polishes"]
        same = r["signal_positions"] == o["signal_positions"]
        n_seg += same.size
        n_same += int(same.sum())
        assert np.abs(r["probabilities"] - o["probabilities"]).max() <= 1e-4
    assert n_same >= 0.999 * n_seg
    print("NTK 3 reads: reference %.2f s, GPU first path %.2f s" % (t_ref, t_gpu))


@pytest.mark.gpu
def test_gpu_ntk_batch_equals_single_reads():
    """dyn_ntk_align_batch (pool of CUDA streams) returns exactly what the one-read entry point returns, in input order,
    and reports per-read errors without failing the batch."""
    from dynamont_b200 import Aligner
    cases = [c for c in load_ntk() if c.has_alignment and c.pore == "rna002"]
    al = Aligner(cases[0].model_path, "rna002", mode="resquiggle")
    sigs = [c.signal for c in cases] * 3 + [cases[0].signal[:8]]
    seqs = [c.sequence for c in cases] * 3 + [cases[0].sequence]
    res = al.align_batch(sigs, seqs, True)
    assert isinstance(res[-1], RuntimeError) and "Signal too short" in str(res[-1])
    for i, c in enumerate(cases * 3):
        r = res[i]
        assert r["Z"] == al.align(c.signal, c.sequence, False)["Z"]
        assert r["states"] == c.states.tolist() and r["polishes"] == c.polishes.tolist()
        assert np.array_equal(r["signal_positions"], c.signal_positions)
