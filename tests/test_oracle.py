"""The CPU oracle (plain-C restatement, oracle/nt_oracle.c) pinned against golden vectors produced by the UNMODIFIED
reference C++ (tools/make_golden.py) and, where the compiled reference is available, against it live."""
import numpy as np
import pytest

import oracle
from conftest import load_golden


@pytest.mark.parametrize("case", load_golden(), ids=lambda c: c.name)
def test_oracle_align_is_bit_exact_with_reference_golden(case):
    o = oracle.Oracle(case.model_path, case.pore)
    r = o.align(case.signal, case.sequence, True)
    assert r["Z"] == case.Z and r["Zf"] == case.Zf
    assert np.array_equal(r["signal_positions"], case.signal_positions)
    assert np.array_equal(r["sequence_positions"], case.sequence_positions)
    assert np.array_equal(r["probabilities"], case.probabilities)  # same operations in the same order
    assert o.align(case.signal, case.sequence, False)["Z"] == case.Z


@pytest.mark.parametrize("case", [c for c in load_golden() if "9mer" not in c.name], ids=lambda c: c.name)
def test_oracle_train_is_bit_exact_with_reference_golden(case):
    o = oracle.Oracle(case.model_path, case.pore)
    t = o.train(case.signal, case.sequence)
    assert t["Z"] == case.train_Z
    tp = t["transition_params"]
    assert np.array_equal([tp["m1"], tp["e1"], tp["e2"]], case.train_trans)
    km = case.train_kmers
    assert np.array_equal(t["w"][km], case.stat_w) and np.array_equal(t["sx"][km], case.stat_x)
    assert np.array_equal(t["emission_model"]["mean"][km], case.train_mean)
    assert np.array_equal(t["emission_model"]["stdev"][km], case.train_stdev)
    # transition re-estimates are data independent in exact arithmetic: every path has N-1 match transitions and
    # T-1-2(N-1) extension transitions
    T, N = case.signal.size + 1, len(case.sequence) - o.k + 2
    assert abs(tp["m1"] - (N - 1) / (T - N)) < 1e-9


@pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built and /root/reference absent")
def test_oracle_matches_live_reference_on_seeded_reads(models_dir):
    from dynamont_b200.synth import materialize_model, native_model, synth_read
    path = materialize_model("rna002_5mer", models_dir)
    ref, orc = oracle.Reference(path, "rna002"), oracle.Oracle(path, "rna002")
    assert np.array_equal(ref.model()[0], orc.mean) and np.array_equal(ref.model()[1], orc.stdev)
    nm, ns = native_model(path, "rna002")
    rng = np.random.default_rng(20260001)
    for L, spb in [(8, 3), (40, 6), (250, 10), (500, 4)]:
        s, q, _ = synth_read(rng, nm, ns, 5, L, spb)
        a, b = ref.align(s, q, True), orc.align(s, q, True)
        assert a["Z"] == b["Z"] and np.array_equal(a["signal_positions"], b["signal_positions"])
        assert np.array_equal(a["probabilities"], b["probabilities"])
        st = ref.stages(s, q, rows=[0, 1, s.size // 2, s.size])
        rows = orc.align(s, q, False, rows=[0, 1, s.size // 2, s.size])["rows"]
        assert np.array_equal(st["rows"], rows)
        assert orc.cells(s.size, len(q)) > 0


def test_oracle_error_messages_match_reference_strings(models_dir):
    from dynamont_b200.synth import materialize_model
    path = materialize_model("rna002_5mer", models_dir)
    o = oracle.Oracle(path, "rna002")
    cases = [(np.zeros(0), "ACGTACGT", "Signal is empty"), (np.zeros(10), "ACG", "Sequence shorter than model kmer size"),
             (np.zeros(5), "ACGTACGTACGT", "Signal too short compared to sequence"),
             (np.zeros(40), "ACGTNACGTA", "Invalid nucleotide: N")]
    impls = [o] + ([oracle.Reference(path, "rna002")] if oracle.have_reference() else [])
    for impl in impls:
        for sig, seq, msg in cases:
            with pytest.raises(RuntimeError) as e:
                impl.align(sig, seq, True)
            assert str(e.value) == msg
