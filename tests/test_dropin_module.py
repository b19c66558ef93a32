"""The ``dynamont._dynamont`` drop-in (dynamont_b200/compat): driven with the call shapes of the reference front end
(segmentation/utils.py:154-191 train_transition_emission / calcZ, segment.py:34-45,161) — through the emulator build on
CPU and through the CUDA library with -m gpu."""
import os
import sys

import numpy as np
import pytest

from conftest import MODELS_DIR, ROOT, load_golden

COMPAT = os.path.join(ROOT, "dynamont_b200", "compat")


def _front_end_calls(Aligner, PoreType, pore_type, lib_kw):
    from dynamont_b200.synth import read_kmer_model
    case = [c for c in load_golden() if c.name == "rna002_short"][0]
    model, read, signal = case.model_path, case.sequence, case.signal
    params = {"r": "rna002", "t": 4, "band": 400, "m1": 0.02, "e1": 1.0, "e2": 0.98}
    # utils._make_native_aligner
    aligner = Aligner(model, params["r"], mode="basic", threads=int(params["t"]), band=int(params["band"]), **lib_kw)
    # utils.train_transition_emission (script == "basic")
    result = aligner.train(signal, read)
    trained = {key: float(value) for key, value in result["transition_params"].items()}
    model_keys = list(read_kmer_model(model)[0])
    new_models = {kmer: (float(entry["mean"]), float(entry["stdev"])) for kmer, entry in zip(model_keys, result["emission_model"])}
    assert set(trained) == {"m1", "e1", "e2"} and len(new_models) == 4 ** 5
    np.testing.assert_allclose([trained["m1"], trained["e1"], trained["e2"]], case.train_trans, rtol=1e-4)
    assert abs(float(result["Z"]) - case.train_Z) <= 1e-6 * abs(case.train_Z)
    # utils.calcZ
    z = float(aligner.align(signal, read, calc_probabilities=False)["Z"])
    assert abs(z - case.Z) <= 1e-6 * abs(case.Z)
    # segment._segment_one: ALIGNER.align(signal, read, calc_probabilities=True) -> utils.segmentation_to_string's inputs
    seg = aligner.align(signal, read, calc_probabilities=True)
    assert set(seg) >= {"Z", "sequence_positions", "signal_positions", "probabilities", "states", "polishes"}
    assert seg["signal_positions"].dtype == np.uint64 and seg["probabilities"].dtype == np.float64
    assert np.array_equal(seg["signal_positions"], case.signal_positions)
    # enum / error surface (aligner_bindings.cpp:18-51)
    assert pore_type("rna004") == PoreType.RNA004
    Aligner(model, PoreType.RNA002, **lib_kw)
    with pytest.raises(ValueError, match="Unknown pore type: nope"):
        Aligner(model, "nope", **lib_kw)
    with pytest.raises(ValueError, match="Unknown aligner mode: fancy"):
        Aligner(model, "rna002", mode="fancy", **lib_kw)
    with pytest.raises(RuntimeError, match="Signal too short compared to sequence"):
        aligner.align(signal[:20], read, True)
    with pytest.raises(ValueError, match="Signal must be a one-dimensional array"):
        aligner.align(np.zeros((4, 4)), read, False)


def _import_dropin():
    sys.path.insert(0, COMPAT)
    try:
        for name in ("dynamont", "dynamont._dynamont"):
            sys.modules.pop(name, None)
        import dynamont
        from dynamont._dynamont import pore_type
        return dynamont.Aligner, dynamont.PoreType, pore_type
    finally:
        sys.path.remove(COMPAT)


def test_dropin_module_through_emulator():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    A, P, pt = _import_dropin()
    _front_end_calls(A, P, pt, {"_lib_path": build_emu.build()})


@pytest.mark.gpu
def test_dropin_module_on_gpu():
    A, P, pt = _import_dropin()
    _front_end_calls(A, P, pt, {})
