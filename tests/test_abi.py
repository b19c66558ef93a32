"""The C-ABI library builds for sm_100a without a GPU, exports every symbol include/dynamont_b200.h declares, and
refuses to run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, have_cuda


@pytest.fixture(scope="module")
def cuda_lib():
    from dynamont_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dynamont_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dyn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(cuda_lib):
    from dynamont_b200 import _capi
    lib = ctypes.CDLL(cuda_lib)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dynamont_b200.h but not exported"
    assert set(_capi.EXPORTS) <= set(syms)


def test_library_contains_sm100a_code(cuda_lib):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", cuda_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_status_messages_are_the_reference_strings(cuda_lib):
    from dynamont_b200 import _capi
    lib = _capi.load(cuda_lib)
    expect = {1: "Signal is empty", 2: "Sequence shorter than model kmer size", 3: "Signal too short compared to sequence",
              4: "Invalid nucleotide: ", 5: "Alignment failed: alignment scores do not match",
              6: "Training failed: alignment scores do not match"}
    for k, v in expect.items():
        assert lib.dyn_status_message(k).decode() == v


@pytest.mark.skipif(have_cuda(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(cuda_lib, models_dir):
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model
    path = materialize_model("rna002_5mer", models_dir)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        Aligner(path, "rna002")


def test_missing_library_fails_loudly(tmp_path):
    from dynamont_b200 import _capi
    with pytest.raises(ImportError, match="no CPU fallback"):
        _capi.load(str(tmp_path / "libdynamont_b200.so"))
