"""ctypes binding of the C ABI declared in include/dynamont_b200.h.

The product loads exactly one library: ``dynamont_b200/csrc/libdynamont_b200.so`` (built by nvcc for
sm_100a, see build.py).  If it is missing, loading fails loudly — there is no CPU fallback.
(``load(path)`` with an explicit path exists so the test-suite can drive the very same binding against the
SIMT-emulator build under tests/emu/.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "csrc", "libdynamont_b200.so")

u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)


class ReadResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("bad_char", C.c_char), ("pad", C.c_char * 3), ("Z", C.c_double),
                ("seg_offset", C.c_uint64), ("n_segments", C.c_uint64)]


class TrainResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("bad_char", C.c_char), ("pad", C.c_char * 3), ("Z", C.c_double),
                ("m1", C.c_double), ("e1", C.c_double), ("e2", C.c_double)]


EXPORTS = [
    "dyn_create", "dyn_destroy", "dyn_kmer_size", "dyn_num_kmers", "dyn_is_rna", "dyn_model", "dyn_set_model",
    "dyn_transitions", "dyn_count_segments", "dyn_read_cells", "dyn_batch_cells", "dyn_align_batch", "dyn_align_batch_f64",
    "dyn_align_batch_device", "dyn_train_batch", "dyn_status_message", "dyn_last_error", "dyn_last_timing",
    "dyn_last_fallbacks", "dyn_last_lin_retries", "dyn_last_variant", "dyn_last_ribbon", "dyn_ribbon_fault_reasons", "dyn_align_submit", "dyn_align_submit_device", "dyn_align_wait", "dyn_train_accumulate", "dyn_train_mstep_device", "dyn_ntk_transitions", "dyn_ntk_prepass", "dyn_ntk_align", "dyn_ntk_align_batch", "dyn_preprocess_batch", "dyn_format_segments", "dyn_set_option", "dyn_set_stream",
]

_libs: dict = {}


def load(path: str | None = None) -> C.CDLL:
    # DYNAMONT_B200_LIB: kernel experiments only (A/B of two builds of the same sources on one GPU box, tools/gpu_ab.sh)
    path = os.path.abspath(path or os.environ.get("DYNAMONT_B200_LIB") or DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise ImportError(
            f"dynamont_b200: CUDA library {path} is missing. Build it with `python -m dynamont_b200.build` "
            "(needs nvcc; there is no CPU fallback).")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.dyn_create.restype = vp
    lib.dyn_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p,
                               C.c_size_t, C.POINTER(C.c_int)]
    lib.dyn_destroy.argtypes = [vp]
    lib.dyn_kmer_size.argtypes = [vp]
    lib.dyn_num_kmers.argtypes = [vp]
    lib.dyn_num_kmers.restype = C.c_uint64
    lib.dyn_is_rna.argtypes = [vp]
    lib.dyn_model.argtypes = [vp, f64p, f64p]
    lib.dyn_set_model.argtypes = [vp, f64p, f64p]
    lib.dyn_transitions.argtypes = [vp, f64p]
    lib.dyn_count_segments.argtypes = [vp, u64p, C.c_uint32]
    lib.dyn_count_segments.restype = C.c_uint64
    lib.dyn_read_cells.argtypes = [vp, C.c_uint64, C.c_uint64]
    lib.dyn_read_cells.restype = C.c_uint64
    lib.dyn_batch_cells.argtypes = [vp, u64p, u64p, C.c_uint32, u64p]
    lib.dyn_batch_cells.restype = C.c_uint64
    common = [u64p, C.c_void_p, u64p, C.c_uint32, C.c_int, C.POINTER(ReadResult), u64p, u64p, f64p]
    lib.dyn_align_batch.argtypes = [vp, C.c_void_p] + common
    lib.dyn_align_batch_f64.argtypes = [vp, C.c_void_p] + common
    lib.dyn_align_batch_device.argtypes = [vp, C.c_void_p] + common
    lib.dyn_align_submit.argtypes = [vp, C.c_void_p] + common
    lib.dyn_align_submit.restype = C.c_int64
    lib.dyn_align_submit_device.argtypes = [vp, C.c_void_p] + common
    lib.dyn_align_submit_device.restype = C.c_int64
    lib.dyn_align_wait.argtypes = [vp, C.c_int64]
    lib.dyn_train_accumulate.argtypes = [vp, C.c_void_p, u64p, C.c_void_p, u64p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
    lib.dyn_train_mstep_device.argtypes = [vp, C.c_void_p, f64p]
    lib.dyn_train_batch.argtypes = [vp, C.c_void_p, u64p, C.c_void_p, u64p, C.c_uint32, C.POINTER(TrainResult),
                                    f64p, f64p, f64p, f64p, f64p, f64p]
    lib.dyn_status_message.argtypes = [C.c_int]
    lib.dyn_status_message.restype = C.c_char_p
    lib.dyn_last_error.argtypes = [vp]
    lib.dyn_last_error.restype = C.c_char_p
    lib.dyn_last_timing.argtypes = [vp, f64p]
    lib.dyn_last_fallbacks.argtypes = [vp]
    lib.dyn_last_fallbacks.restype = C.c_uint64
    lib.dyn_last_variant.argtypes = [vp]
    lib.dyn_last_ribbon.argtypes = [vp, u64p]
    lib.dyn_ribbon_fault_reasons.argtypes = [vp, u64p]
    lib.dyn_last_lin_retries.argtypes = [vp]
    lib.dyn_last_lin_retries.restype = C.c_uint64
    lib.dyn_ntk_transitions.argtypes = [vp, f64p]
    lib.dyn_ntk_prepass.argtypes = [vp, C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_uint64, u64p, f64p]
    lib.dyn_ntk_align.argtypes = [vp, C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_double), u64p,
                                  C.c_char_p, u64p, u64p, f64p, C.c_void_p, C.c_uint64]
    lib.dyn_ntk_align_batch.argtypes = [vp, C.c_void_p, u64p, C.c_void_p, u64p, C.c_uint32, C.c_int, C.c_void_p, f64p, u64p, u64p,
                                        C.c_void_p, u64p, u64p, f64p, C.c_void_p, C.c_int]
    lib.dyn_preprocess_batch.argtypes = [vp, C.c_void_p, u64p, C.c_uint32, f64p, f64p, C.c_int, C.c_double, C.c_void_p]
    lib.dyn_format_segments.restype = C.c_int64
    lib.dyn_format_segments.argtypes = [C.c_char_p, C.c_char_p, C.c_int64, C.c_int64, C.c_char_p, C.c_int, C.c_int, C.c_uint64,
                                        u64p, u64p, f64p, C.c_char_p, C.POINTER(C.c_char_p), C.c_char_p, C.c_uint64]
    lib.dyn_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    lib.dyn_set_stream.argtypes = [vp, C.c_void_p]
    _libs[path] = lib
    return lib
