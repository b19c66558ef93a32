"""``dynamont._dynamont`` — the module name of the reference's pybind11 extension (src/cpp/aligner_bindings.cpp:180-219),
served by the B200 library through ctypes.  Same surface:

    Aligner(model_file, pore, mode="basic", threads=1, band=400)        aligner_bindings.cpp:191-199
        .align(signal, sequence, calc_probabilities=False) -> dict      :53-84, :132-147
        .train(signal, sequence) -> dict                                :86-107, :149-163
    PoreType, pore_type(str)                                            :18-32, :184-189, :218

Exceptions carry the reference's message strings (RuntimeError / ValueError as pybind11 translates them)."""
import os
import sys

_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from dynamont_b200.aligner import Aligner, PoreType, pore_type  # noqa: E402,F401

__all__ = ["Aligner", "PoreType", "pore_type"]
