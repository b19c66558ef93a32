"""TEST INFRASTRUCTURE — CPU oracles for the Dynamont basic-mode hot path.

``Reference``  wraps ``oracle/_ref/libdynamont_ref.so`` = the unmodified reference C++ (see build.py).
``Oracle``     wraps ``oracle/_build/liboracle_nt.so`` = the plain-C restatement (nt_oracle.c) plus a
               numpy restatement of the host-side helpers (model loading, kmer encoding, validation).

Both expose the reference's operator surface (``aligner_bindings.cpp:53-107``): ``align`` returns
``{"Z", "sequence_positions", "signal_positions", "probabilities", "states", "polishes"}`` and ``train``
returns ``{"Z", "transition_params", "emission_model"}`` (emission model as two arrays, not 4^k dicts),
raising ``RuntimeError``/``ValueError`` with the reference's message strings.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  The product (dynamont_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import build as _build

_c_double_p = C.POINTER(C.c_double)
_c_size_p = C.POINTER(C.c_size_t)
_c_u64_p = C.POINTER(C.c_uint64)
_c_int_p = C.POINTER(C.c_int)

PORES = {  # aligner.cpp:62-86 (rna, k); NT:36-82 (m1, e1, e2 as probabilities)
    "rna002": (True, 5, (0.019889650396799997, 1.0, 0.9801103496029998)),
    "rna004": (True, 9, (0.031111753637096777, 1.0, 0.9688882463622581)),
    "dna_r9": (False, 5, (1.0, 1.0, 1.0)),
    "dna_r10_260bps": (False, 9, (0.031111753637096777, 1.0, 0.9688882463622581)),
    "dna_r10_400bps": (False, 9, (0.031111753637096777, 1.0, 0.9688882463622581)),
}

_BASE = np.full(256, -1, dtype=np.int64)  # aligner.cpp:46-60
for _ch, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3), ("N", 4)):
    _BASE[ord(_ch)] = _v
    _BASE[ord(_ch.lower())] = _v


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


def have_reference() -> bool:
    return _build.build_reference() is not None


class Reference:
    """The unmodified reference C++ (oracle/_ref)."""

    def __init__(self, model_file: str, pore: str, mode: str = "basic", band: int = 400, ntk_fix: bool = False):
        # ntk_fix: the reference with the two-line repair of resquiggle mode (oracle/build.py)
        path = _build.build_reference_ntkfix() if ntk_fix else _build.build_reference()
        if path is None:
            raise FileNotFoundError("oracle/_ref reference library not built and /root/reference absent")
        self._lib = lib = C.CDLL(path)
        lib.ref_create.restype = C.c_void_p
        lib.ref_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        lib.ref_destroy.argtypes = [C.c_void_p]
        lib.ref_kmer_size.argtypes = [C.c_void_p]
        lib.ref_num_kmers.argtypes = [C.c_void_p]
        lib.ref_num_kmers.restype = C.c_long
        lib.ref_is_rna.argtypes = [C.c_void_p]
        lib.ref_model.argtypes = [C.c_void_p, _c_double_p, _c_double_p]
        lib.ref_align.argtypes = [C.c_void_p, _c_double_p, C.c_size_t, C.c_char_p, C.c_int, _c_double_p,
                                  _c_size_p, _c_size_p, _c_size_p, _c_double_p, C.c_char_p, C.c_char_p,
                                  C.c_char_p, C.c_size_t]
        lib.ref_train.argtypes = [C.c_void_p, _c_double_p, C.c_size_t, C.c_char_p, _c_double_p, _c_double_p,
                                  _c_double_p, _c_double_p, C.c_char_p, C.c_size_t]
        lib.ref_nt_stages.argtypes = [C.c_void_p, _c_double_p, C.c_size_t, C.c_char_p, _c_double_p,
                                      _c_double_p, _c_size_p, C.c_size_t, _c_double_p, _c_double_p,
                                      _c_double_p, _c_double_p, C.c_char_p, C.c_size_t]
        err = C.create_string_buffer(512)
        self._h = lib.ref_create(model_file.encode(), pore.encode(), mode.encode(), band, err, 512)
        if not self._h:
            msg = err.value.decode()
            if msg.startswith("Unknown pore type:") or msg.startswith("Unknown aligner mode"):
                raise ValueError(msg)
            raise RuntimeError(msg)
        self.k = lib.ref_kmer_size(self._h)
        self.K = lib.ref_num_kmers(self._h)
        self.rna = bool(lib.ref_is_rna(self._h))
        self.pore = pore
        self.band = band

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_destroy(self._h)
            self._h = None

    def model(self):
        mean = np.empty(self.K)
        sd = np.empty(self.K)
        self._lib.ref_model(self._h, _ptr(mean, _c_double_p), _ptr(sd, _c_double_p))
        return mean, sd

    def align(self, signal, sequence: str, calc_probabilities: bool = False) -> dict:
        sig = np.ascontiguousarray(signal, dtype=np.float64)
        if sig.ndim != 1:
            raise ValueError("Signal must be a one-dimensional array")
        L = max(len(sequence), 1)
        Z = C.c_double()
        n = C.c_size_t()
        seqpos = np.zeros(L, dtype=np.uintp)
        sigpos = np.zeros(L, dtype=np.uintp)
        prob = np.zeros(L)
        state = C.create_string_buffer(L + 1)
        kw = self.k + 1
        polish = C.create_string_buffer(L * kw + 1)
        err = C.create_string_buffer(512)
        rc = self._lib.ref_align(self._h, _ptr(sig, _c_double_p), sig.size, sequence.encode(),
                                 int(calc_probabilities), C.byref(Z), C.byref(n), _ptr(seqpos, _c_size_p),
                                 _ptr(sigpos, _c_size_p), _ptr(prob, _c_double_p), state, polish, err, 512)
        if rc:
            raise RuntimeError(err.value.decode())
        ns = n.value
        raw = polish.raw
        return {
            "Z": Z.value,
            "sequence_positions": seqpos[:ns].astype(np.uint64),
            "signal_positions": sigpos[:ns].astype(np.uint64),
            "probabilities": prob[:ns].copy(),
            "states": [chr(b) for b in state.raw[:ns]],
            "polishes": [raw[i * kw:(i + 1) * kw].split(b"\0")[0].decode() for i in range(ns)],
        }

    def train(self, signal, sequence: str) -> dict:
        sig = np.ascontiguousarray(signal, dtype=np.float64)
        Z = C.c_double()
        trans = np.zeros(3)
        mean = np.zeros(self.K)
        sd = np.zeros(self.K)
        err = C.create_string_buffer(512)
        rc = self._lib.ref_train(self._h, _ptr(sig, _c_double_p), sig.size, sequence.encode(), C.byref(Z),
                                 _ptr(trans, _c_double_p), _ptr(mean, _c_double_p), _ptr(sd, _c_double_p),
                                 err, 512)
        if rc:
            raise RuntimeError(err.value.decode())
        return {"Z": Z.value, "transition_params": {"m1": trans[0], "e1": trans[1], "e2": trans[2]},
                "emission_model": {"mean": mean, "stdev": sd}}

    def stages(self, signal, sequence: str, rows=(), stats: bool = False) -> dict:
        """Private forward/backward of the reference: Zf, Zb, lattice rows [len(rows), 4(fM,fE,bM,bE), N]
        and (optionally) the raw per-kmer sufficient statistics of runTraining."""
        sig = np.ascontiguousarray(signal, dtype=np.float64)
        rows = np.ascontiguousarray(rows, dtype=np.uintp)
        N = len(sequence) - self.k + 2
        out = np.empty((rows.size, 4, N))
        Zf, Zb = C.c_double(), C.c_double()
        w = sx = sxx = None
        if stats:
            w, sx, sxx = np.zeros(self.K), np.zeros(self.K), np.zeros(self.K)
        err = C.create_string_buffer(512)
        rc = self._lib.ref_nt_stages(
            self._h, _ptr(sig, _c_double_p), sig.size, sequence.encode(), C.byref(Zf), C.byref(Zb),
            _ptr(rows, _c_size_p), rows.size, _ptr(out, _c_double_p),
            _ptr(w, _c_double_p) if stats else None, _ptr(sx, _c_double_p) if stats else None,
            _ptr(sxx, _c_double_p) if stats else None, err, 512)
        if rc:
            raise RuntimeError(err.value.decode())
        return {"Zf": Zf.value, "Zb": Zb.value, "rows": out, "w": w, "sx": sx, "sxx": sxx}


def _ntk_prepass(self, signal, sequence: str, want_z: bool = True) -> dict:
    """Reference resquiggle (NTK) mode, the stages that work as shipped (NTK_aligner_api.cpp:197-441): row masks
    of the dense TN / TK pre-passes and the sorted sparse-lattice keys."""
    sig = np.ascontiguousarray(signal, dtype=np.float64)
    T, N, K = sig.size + 1, len(sequence) - self.k + 2, self.K
    tn = np.zeros((T, N), dtype=np.uint8)
    tk = np.zeros((T, K), dtype=np.uint8)
    cap = 64 * T * 64
    keys = np.zeros(cap, dtype=np.uint64)
    nk = C.c_size_t()
    z4 = np.zeros(4)
    tr = np.zeros(18)
    err = C.create_string_buffer(512)
    self._lib.ref_ntk_prepass.argtypes = [C.c_void_p, _c_double_p, C.c_size_t, C.c_char_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), _c_double_p, _c_double_p,
                                          C.c_char_p, C.c_size_t]
    rc = self._lib.ref_ntk_prepass(self._h, _ptr(sig, _c_double_p), sig.size, sequence.encode(), tn.ctypes.data,
                                   tk.ctypes.data, keys.ctypes.data, cap, C.byref(nk),
                                   _ptr(z4, _c_double_p) if want_z else None, _ptr(tr, _c_double_p), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    if nk.value > cap:
        raise RuntimeError("key buffer too small")
    return {"tn": tn.astype(bool), "tk": tk.astype(bool), "keys": keys[:nk.value].copy(), "Z": z4, "transitions": tr}


Reference.ntk_prepass = _ntk_prepass


def load_model_native(path: str, pore: str):
    """numpy restatement of Aligner::loadModel (aligner.cpp:88-143): returns (mean[K], stdev[K]) indexed
    by kmerToInt(rna ? reversed(kmer) : kmer) (aligner.cpp:136-141, 207-220)."""
    if pore not in PORES:
        raise ValueError("Unknown pore type: " + pore)
    rna, k, _ = PORES[pore]
    try:
        fh = open(path)
    except OSError:
        raise RuntimeError("Could not open model file, please prove a valid model path " + path)
    with fh:
        lines = fh.read().split("\n")[1:]
    rows = [ln.split("\t") for ln in lines if ln != ""]
    alphabet = set()
    for r in rows:
        if len(r[0]) != k:
            raise RuntimeError("Inconsistent kmer size in model")
        alphabet.update(r[0])
    A = len(alphabet)
    K = int(round(math.pow(A, k)))
    mean = np.zeros(K)
    sd = np.zeros(K)
    for r in rows:
        kmer = r[0][::-1] if rna else r[0]
        v = 0
        for ch in kmer:
            d = int(_BASE[ord(ch)])
            if d < 0 or d >= A:
                raise RuntimeError("Invalid nucleotide in k-mer: " + kmer)
            v = v * A + d
        mean[v] = float(r[1])
        sd[v] = float(r[2])
    return mean, sd, A


def sequence_to_kmers(sequence: str, k: int, alphabet: int = 4) -> np.ndarray:
    """aligner.cpp:166-205 (validation order and messages included)."""
    if len(sequence) < k:
        raise RuntimeError("Sequence shorter than kmer size")
    d = _BASE[np.frombuffer(sequence.encode("latin-1"), dtype=np.uint8)]
    bad = np.nonzero((d < 0) | (d >= alphabet))[0]
    if bad.size:
        raise RuntimeError("Invalid nucleotide: " + sequence[int(bad[0])])
    Kc = len(sequence) - k + 1
    v = np.zeros(Kc, dtype=np.int64)
    for i in range(k):
        v = v * alphabet + d[i:i + Kc]
    return v.astype(np.int32)


def validate_input(S: int, L: int, k: int):
    """aligner.cpp:145-164."""
    if S < 1:
        raise RuntimeError("Signal is empty")
    if L < k:
        raise RuntimeError("Sequence shorter than model kmer size")
    if S < 2 * (L - k + 1):
        raise RuntimeError("Signal too short compared to sequence")


class Oracle:
    """The plain-C restatement (nt_oracle.c) behind the same operator surface (basic mode only)."""

    def __init__(self, model_file: str, pore: str, mode: str = "basic", band: int = 400):
        if pore not in PORES:
            raise ValueError("Unknown pore type: " + pore)
        if mode not in ("basic", "nt"):
            raise ValueError("Unknown aligner mode: " + mode if mode not in ("resquiggle", "ntk")
                             else "oracle restatement covers basic mode only")
        self.rna, self.k, probs = PORES[pore]
        self.trans = np.log(np.array(probs, dtype=np.float64))  # NT:84-86
        self.mean, self.stdev, self.alphabet = load_model_native(model_file, pore)
        self.K = self.mean.size
        self.band = band
        self.pore = pore
        self._lib = lib = C.CDLL(_build.build_oracle())
        lib.nt_oracle_cells.restype = C.c_uint64
        lib.nt_oracle_cells.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t]
        lib.nt_oracle_align.argtypes = [_c_double_p, C.c_size_t, _c_int_p, C.c_size_t, C.c_int, _c_double_p,
                                        _c_double_p, _c_double_p, C.c_size_t, C.c_int, _c_double_p,
                                        _c_double_p, _c_u64_p, _c_u64_p, _c_double_p, _c_size_p, C.c_size_t,
                                        _c_double_p]
        lib.nt_oracle_train.argtypes = [_c_double_p, C.c_size_t, _c_int_p, C.c_size_t, C.c_size_t,
                                        _c_double_p, _c_double_p, _c_double_p, C.c_size_t, _c_double_p,
                                        _c_double_p, _c_double_p, _c_double_p, _c_double_p, _c_double_p,
                                        _c_double_p, _c_double_p]

    def model(self):
        return self.mean, self.stdev

    def cells(self, S: int, L: int) -> int:
        return int(self._lib.nt_oracle_cells(S, L - self.k + 1, self.band))

    def _prep(self, signal, sequence):
        sig = np.ascontiguousarray(signal, dtype=np.float64)
        if sig.ndim != 1:
            raise ValueError("Signal must be a one-dimensional array")
        validate_input(sig.size, len(sequence), self.k)
        kmers = sequence_to_kmers(sequence, self.k, self.alphabet)
        return sig, kmers

    def align(self, signal, sequence: str, calc_probabilities: bool = False, rows=()) -> dict:
        sig, kmers = self._prep(signal, sequence)
        Kc = kmers.size
        seqpos = np.zeros(Kc, dtype=np.uint64)
        sigpos = np.zeros(Kc, dtype=np.uint64)
        prob = np.zeros(Kc)
        rows = np.ascontiguousarray(rows, dtype=np.uintp)
        rows_out = np.empty((rows.size, 4, Kc + 1))
        Zf, Zb = C.c_double(), C.c_double()
        rc = self._lib.nt_oracle_align(
            _ptr(sig, _c_double_p), sig.size, _ptr(kmers, _c_int_p), Kc, self.k, _ptr(self.mean, _c_double_p),
            _ptr(self.stdev, _c_double_p), _ptr(self.trans, _c_double_p), self.band, int(calc_probabilities),
            C.byref(Zf), C.byref(Zb), _ptr(seqpos, _c_u64_p), _ptr(sigpos, _c_u64_p), _ptr(prob, _c_double_p),
            _ptr(rows, _c_size_p), rows.size, _ptr(rows_out, _c_double_p))
        if rc == 1:
            raise RuntimeError("Alignment failed: alignment scores do not match")
        if rc:
            raise MemoryError("nt_oracle_align rc=%d" % rc)
        ns = Kc if calc_probabilities else 0
        return {"Z": Zb.value, "Zf": Zf.value, "sequence_positions": seqpos[:ns], "signal_positions": sigpos[:ns],
                "probabilities": prob[:ns], "states": ["M"] * ns, "polishes": [""] * ns, "rows": rows_out}

    def train(self, signal, sequence: str) -> dict:
        sig, kmers = self._prep(signal, sequence)
        K = self.K
        Z = C.c_double()
        trans = np.zeros(3)
        xi = np.zeros(2)
        w, sx, sxx, nm, ns = (np.zeros(K) for _ in range(5))
        rc = self._lib.nt_oracle_train(
            _ptr(sig, _c_double_p), sig.size, _ptr(kmers, _c_int_p), kmers.size, K, _ptr(self.mean, _c_double_p),
            _ptr(self.stdev, _c_double_p), _ptr(self.trans, _c_double_p), self.band, C.byref(Z),
            _ptr(trans, _c_double_p), _ptr(w, _c_double_p), _ptr(sx, _c_double_p), _ptr(sxx, _c_double_p),
            _ptr(nm, _c_double_p), _ptr(ns, _c_double_p), _ptr(xi, _c_double_p))
        if rc == 1:
            raise RuntimeError("Training failed: alignment scores do not match")
        if rc:
            raise MemoryError("nt_oracle_train rc=%d" % rc)
        return {"Z": Z.value, "transition_params": {"m1": trans[0], "e1": trans[1], "e2": trans[2]},
                "emission_model": {"mean": nm, "stdev": ns}, "w": w, "sx": sx, "sxx": sxx, "xi": xi}
