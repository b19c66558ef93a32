"""Host-side mirror of the reference's native operator API for the segmentation hot path.

Same names, argument meaning and error behaviour as ``dynamont._dynamont`` (reference
src/cpp/aligner_bindings.cpp:180-219): ``Aligner(model_file, pore, mode="basic", threads=1, band=400)``,
``.align(signal, sequence, calc_probabilities=False) -> dict``, ``.train(signal, sequence) -> dict``,
``PoreType`` and ``pore_type()``.  Added on top (the reference processes one read per call; a B200 needs
thousands in flight): ``.align_batch`` / ``.train_batch``.

Everything numeric happens in the CUDA library behind the C ABI (include/dynamont_b200.h).
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _capi
from ._capi import ReadResult, TrainResult, f64p, u64p


class PoreType(enum.Enum):  # aligner.hpp:26-33 / aligner_bindings.cpp:184-189
    RNA002 = 0
    RNA004 = 1
    DNA_R9 = 2
    DNA_R10_260 = 3
    DNA_R10_400 = 4


_PORE_STR = {
    "rna002": PoreType.RNA002, "rna004": PoreType.RNA004, "dna_r9": PoreType.DNA_R9,
    "dna_r10_260bps": PoreType.DNA_R10_260, "dna_r10_400bps": PoreType.DNA_R10_400,
}
_PORE_NAME = {v: k for k, v in _PORE_STR.items()}


def pore_type(pore: str) -> PoreType:
    """aligner_bindings.cpp:18-32."""
    try:
        return _PORE_STR[pore]
    except KeyError:
        raise ValueError("Unknown pore type: " + str(pore))


def _message(lib, status: int, bad_char: bytes) -> str:
    msg = lib.dyn_status_message(status).decode()
    if status == 4:
        msg += bad_char.decode("latin-1")
    return msg


class Aligner:
    def __init__(self, model_file: str, pore, mode: str = "basic", threads: int = 1, band: int = 400,
                 device: int = -1, _lib_path: str | None = None):
        self._lib = _capi.load(_lib_path)
        if isinstance(pore, PoreType):
            pore = _PORE_NAME[pore]
        err = C.create_string_buffer(1024)
        kind = C.c_int(0)
        self._h = self._lib.dyn_create(str(model_file).encode(), str(pore).encode(), str(mode).encode(),
                                       int(threads), int(band), int(device), err, 1024, C.byref(kind))
        if not self._h:
            msg = err.value.decode()
            raise (ValueError if kind.value == 1 else RuntimeError)(msg)
        self.kmer_size = self._lib.dyn_kmer_size(self._h)
        self._ntk = str(mode) in ("resquiggle", "ntk")
        self.num_kmers = int(self._lib.dyn_num_kmers(self._h))
        self.rna = bool(self._lib.dyn_is_rna(self._h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.dyn_destroy(h)
            self._h = None

    # ------------------------------------------------------------------------------------------ helpers
    def model(self):
        mean = np.empty(self.num_kmers)
        sd = np.empty(self.num_kmers)
        self._lib.dyn_model(self._h, mean.ctypes.data_as(f64p), sd.ctypes.data_as(f64p))
        return mean, sd

    def set_model(self, mean, stdev) -> None:
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        stdev = np.ascontiguousarray(stdev, dtype=np.float64)
        assert mean.size == self.num_kmers and stdev.size == self.num_kmers
        self._lib.dyn_set_model(self._h, mean.ctypes.data_as(f64p), stdev.ctypes.data_as(f64p))

    def set_option(self, key: str, value: float) -> None:
        if self._lib.dyn_set_option(self._h, key.encode(), float(value)) != 0:
            raise KeyError(key)

    def set_stream(self, cuda_stream: int) -> None:
        """Run this handle's work on the given ``cudaStream_t`` (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        self._lib.dyn_set_stream(self._h, C.c_void_p(cuda_stream))

    def batch_cells(self, sig_off, seq_off) -> int:
        sig_off = np.ascontiguousarray(sig_off, dtype=np.uint64)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        return int(self._lib.dyn_batch_cells(self._h, sig_off.ctypes.data_as(u64p), seq_off.ctypes.data_as(u64p),
                                             sig_off.size - 1, None))

    def align_packed(self, signal_ptr: int, sig_off, seq_ptr: int, seq_off, calc_probabilities: bool = True,
                     device: bool = False, f64: bool = False):
        """Zero-copy batched entry used by bench.py / streaming front ends: ``signal_ptr``/``seq_ptr`` are raw
        addresses of the concatenated samples / bases (host, or device memory if ``device``).  Returns
        (results ctypes array, sequence_positions, signal_positions, probabilities)."""
        sig_off = np.ascontiguousarray(sig_off, dtype=np.uint64)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        n = sig_off.size - 1
        res = (ReadResult * max(n, 1))()
        nseg = int(self._lib.dyn_count_segments(self._h, seq_off.ctypes.data_as(u64p), n))
        seqpos = np.empty(max(nseg, 1), dtype=np.uint64)
        sigpos = np.empty(max(nseg, 1), dtype=np.uint64)
        prob = np.empty(max(nseg, 1), dtype=np.float64)
        fn = self._lib.dyn_align_batch_device if device else (
            self._lib.dyn_align_batch_f64 if f64 else self._lib.dyn_align_batch)
        rc = fn(self._h, C.c_void_p(signal_ptr), sig_off.ctypes.data_as(u64p), C.c_void_p(seq_ptr),
                seq_off.ctypes.data_as(u64p), n, int(calc_probabilities), res, seqpos.ctypes.data_as(u64p),
                sigpos.ctypes.data_as(u64p), prob.ctypes.data_as(f64p))
        if rc != 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        return res, seqpos, sigpos, prob

    # ---- asynchronous batches (dyn_align_submit / dyn_align_wait) -------------------------------------------------
    def submit_packed(self, signal_ptr: int, sig_off, seq_ptr: int, seq_off, calc_probabilities: bool = True, keep=None,
                      device: bool = False, out=None):
        """Start a batch of HOST-resident reads (float32 samples at ``signal_ptr``, ASCII bases at ``seq_ptr``) on one of
        the handle's two lanes and return a job; ``wait(job)`` returns what ``align_packed`` returns.  With two jobs in
        flight the copies and the result fan-out of neighbouring batches overlap the kernels.  ``keep``: any object that
        owns the input memory (kept alive until the job is waited for).  ``out``: optional (results, sequence_positions,
        signal_positions, probabilities) buffers of a previous, completed job to write into (a streaming caller rotates a
        few sets instead of allocating — and page-faulting — fresh result arrays for every batch)."""
        sig_off = np.ascontiguousarray(sig_off, dtype=np.uint64)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        n = sig_off.size - 1
        nseg = int(self._lib.dyn_count_segments(self._h, seq_off.ctypes.data_as(u64p), n))
        if out is not None and len(out[0]) >= max(n, 1) and out[1].size >= max(nseg, 1):
            res, seqpos, sigpos, prob = out
        else:
            res = (ReadResult * max(n, 1))()
            seqpos = np.empty(max(nseg, 1), dtype=np.uint64)
            sigpos = np.empty(max(nseg, 1), dtype=np.uint64)
            prob = np.empty(max(nseg, 1), dtype=np.float64)
        fn = self._lib.dyn_align_submit_device if device else self._lib.dyn_align_submit
        ticket = fn(self._h, C.c_void_p(signal_ptr), sig_off.ctypes.data_as(u64p), C.c_void_p(seq_ptr),
                                            seq_off.ctypes.data_as(u64p), n, int(calc_probabilities), res,
                                            seqpos.ctypes.data_as(u64p), sigpos.ctypes.data_as(u64p), prob.ctypes.data_as(f64p))
        if ticket < 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        return {"ticket": ticket, "out": (res, seqpos, sigpos, prob), "keep": (keep, sig_off, seq_off)}

    def wait(self, job):
        rc = self._lib.dyn_align_wait(self._h, job["ticket"])
        if rc != 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        return job["out"]

    def align_stream(self, batches, calc_probabilities: bool = True, depth: int = 2):
        """Pipelined alignment of an iterable of (signals, sequences) batches: yields, in input order, the list that
        ``align_batch`` would return for each batch while up to ``depth`` batches are in flight on the device."""
        from collections import deque
        inflight = deque()

        def finish(item):
            job, meta = item
            res, seqpos, sigpos, prob = self.wait(job)
            return self._unpack(meta, res, seqpos, sigpos, prob, calc_probabilities)
        for signals, sequences in batches:
            sig, sig_off, seq, seq_off = self._pack(signals, sequences, np.float32)
            seqbuf = np.frombuffer(seq, dtype=np.uint8)
            job = self.submit_packed(sig.ctypes.data, sig_off, seqbuf.ctypes.data, seq_off, calc_probabilities, keep=(sig, seq, seqbuf))
            inflight.append((job, len(signals)))
            if len(inflight) >= depth:
                yield finish(inflight.popleft())
        while inflight:
            yield finish(inflight.popleft())

    # ---- front-end stages either side of the DP ------------------------------------------
    def preprocess_batch(self, raws, shifts, scales, window: int = 3, n_sigmas: float = 3.0):
        """(raw - shift) / scale + Hampel filter (utils.py:16-43, segment.py:151-153) on the GPU; returns float32 arrays."""
        sig, sig_off, _, _ = self._pack(raws, [""] * len(raws), np.float32)
        sh = np.ascontiguousarray(shifts, dtype=np.float64)
        sc = np.ascontiguousarray(scales, dtype=np.float64)
        out = np.empty_like(sig)
        rc = self._lib.dyn_preprocess_batch(self._h, C.c_void_p(sig.ctypes.data), sig_off.ctypes.data_as(u64p), len(raws),
                                            sh.ctypes.data_as(f64p), sc.ctypes.data_as(f64p), int(window), float(n_sigmas),
                                            C.c_void_p(out.ctypes.data))
        if rc != 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        return [out[int(sig_off[i]):int(sig_off[i + 1])] for i in range(len(raws))]

    def format_segments(self, result: dict, readid: str, signalid: str, sig_offset: int, last_index: int, read: str) -> bytes:
        """utils.segmentation_to_string (utils.py:193-232) for one alignment result of this aligner."""
        n = len(result["sequence_positions"])
        seqpos = np.ascontiguousarray(result["sequence_positions"], dtype=np.uint64)
        sigpos = np.ascontiguousarray(result["signal_positions"], dtype=np.uint64)
        prob = np.ascontiguousarray(result["probabilities"], dtype=np.float64)
        states = "".join(result["states"]).encode("ascii")
        pol = result.get("polishes")
        arr = None
        if pol is not None:
            arr = (C.c_char_p * max(n, 1))(*[p.encode("ascii") for p in pol])
        args = [readid.encode(), signalid.encode(), int(sig_offset), int(last_index), read.encode("latin-1"), self.kmer_size,
                int(self.rna), n, seqpos.ctypes.data_as(u64p), sigpos.ctypes.data_as(u64p), prob.ctypes.data_as(f64p), states, arr]
        need = self._lib.dyn_format_segments(*args, None, 0)
        buf = C.create_string_buffer(max(int(need), 1))
        self._lib.dyn_format_segments(*args, buf, need)
        return buf.raw[:need]

    # ---- resquiggle (NTK) mode: pre-pass stages (reference NTK_aligner_api.cpp:120-441) ---------------------
    def ntk_transitions(self) -> dict:
        t = np.zeros(18)
        self._lib.dyn_ntk_transitions(self._h, t.ctypes.data_as(f64p))
        names = ["a1", "a2", "p1", "p2", "p3", "s1", "s2", "s3", "e1", "e2", "e3", "e4", "i1", "i2",
                 "tn_m", "tn_e", "tk_m", "tk_e"]
        return dict(zip(names, t.tolist()))

    def ntk_prepass(self, signal, sequence: str) -> dict:
        """Dense TN / TK pre-passes of resquiggle mode on the GPU: boolean row masks ``tn`` [T, N] and ``tk`` [T, K]
        (the reference's tnMap / tkMap), the sorted sparse-lattice ``keys`` and ``Z`` = (Zf, Zb) of both passes."""
        sig = np.ascontiguousarray(signal, dtype=np.float32)
        seq = sequence.encode("latin-1")
        S, L = sig.size, len(seq)
        T, N, K = S + 1, max(L - self.kmer_size + 2, 1), self.num_kmers
        wn, wk = (N + 31) // 32, (K + 31) // 32
        tn = np.zeros((T, wn), dtype=np.uint32)
        tk = np.zeros((T, wk), dtype=np.uint32)
        z4 = np.zeros(4)
        nk = C.c_uint64(0)
        cap = 1 << 16
        while True:
            keys = np.zeros(cap, dtype=np.uint64)
            rc = self._lib.dyn_ntk_prepass(self._h, C.c_void_p(sig.ctypes.data if S else 0), S, seq, L, C.c_void_p(tn.ctypes.data),
                                           C.c_void_p(tk.ctypes.data), C.c_void_p(keys.ctypes.data), cap, C.byref(nk),
                                           z4.ctypes.data_as(f64p))
            if rc < 0:
                raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
            if rc > 0:
                raise RuntimeError(self._lib.dyn_status_message(rc).decode())
            if nk.value <= cap:
                break
            cap = int(nk.value)
        unpack = lambda m, C_: np.unpackbits(m.view(np.uint8), axis=1, bitorder="little")[:, :C_].astype(bool)  # noqa: E731
        return {"tn": unpack(tn, N), "tk": unpack(tk, K), "keys": keys[:nk.value].copy(), "Z": z4}

    def read_cells(self, S: int, L: int) -> int:
        return int(self._lib.dyn_read_cells(self._h, int(S), int(L)))

    def ribbon_fault_reasons(self):
        """cumulative {reason code: reads} the ribbon kernels handed to the full-band kernels (csrc/dp_ribbon.cuh)"""
        r = np.zeros(16, dtype=np.uint64)
        self._lib.dyn_ribbon_fault_reasons(self._h, r.ctypes.data_as(u64p))
        out = {int(i): int(v) for i, v in enumerate(r[:13]) if v}
        out["records_per_row"] = float(r[13]) / 1000.0
        out["two_level_checkpoints"] = bool(r[14])
        out["records_free_layout"] = int(r[14]) == 2  # long reads: path posteriors from a second forward sweep
        out["kept_by_log2_ribbon"] = int(r[15])         # of the reads counted above: re-run by the log2-domain ribbon, not handed on
        return out

    def last_timing(self):
        t = np.zeros(3)
        self._lib.dyn_last_timing(self._h, t.ctypes.data_as(f64p))
        rb = np.zeros(2, dtype=np.uint64)
        self._lib.dyn_last_ribbon(self._h, rb.ctypes.data_as(u64p))
        return {"encode_ms": t[0], "dp_ms": t[1], "launches": int(t[2]),
                "ribbon_reads": int(rb[0]), "ribbon_faults": int(rb[1]),
                "log2_fallback_reads": int(self._lib.dyn_last_fallbacks(self._h)),
                "lin_retry_reads": int(self._lib.dyn_last_lin_retries(self._h)),
                "variant": int(self._lib.dyn_last_variant(self._h))}

    @staticmethod
    def _pack(signals, sequences, dtype):
        n = len(signals)
        sig_off = np.zeros(n + 1, dtype=np.uint64)
        seq_off = np.zeros(n + 1, dtype=np.uint64)
        arrs = []
        for i, s in enumerate(signals):
            a = np.ascontiguousarray(s, dtype=dtype)
            if a.ndim != 1:
                raise ValueError("Signal must be a one-dimensional array")  # aligner_bindings.cpp:138
            arrs.append(a)
            sig_off[i + 1] = sig_off[i] + a.size
        sig = np.concatenate(arrs) if arrs else np.zeros(0, dtype=dtype)
        if sig.size == 0:
            sig = np.zeros(1, dtype=dtype)  # keep a valid pointer
        seqb = [q.encode("latin-1") if isinstance(q, str) else bytes(q) for q in sequences]
        for i, q in enumerate(seqb):
            seq_off[i + 1] = seq_off[i] + len(q)
        seq = b"".join(seqb) or b"\0"
        return sig, sig_off, seq, seq_off

    # --------------------------------------------------------------------------------------------- align
    def align_batch(self, signals, sequences, calc_probabilities: bool = False, raise_errors: bool = False):
        """Batched ``align``: returns one dict per read (same keys as ``align``) or an Exception instance for
        reads the reference would have thrown on (``raise_errors`` re-raises the first)."""
        n = len(signals)
        assert len(sequences) == n
        if self._ntk:
            return self._ntk_align_batch(signals, sequences, calc_probabilities, raise_errors)
        dtype = np.float64 if any(np.asarray(s).dtype == np.float64 for s in signals) else np.float32
        sig, sig_off, seq, seq_off = self._pack(signals, sequences, dtype)
        res = (ReadResult * max(n, 1))()
        nseg = int(self._lib.dyn_count_segments(self._h, seq_off.ctypes.data_as(u64p), n))
        seqpos = np.zeros(max(nseg, 1), dtype=np.uint64)
        sigpos = np.zeros(max(nseg, 1), dtype=np.uint64)
        prob = np.zeros(max(nseg, 1), dtype=np.float64)
        fn = self._lib.dyn_align_batch_f64 if dtype == np.float64 else self._lib.dyn_align_batch
        rc = fn(self._h, sig.ctypes.data, sig_off.ctypes.data_as(u64p), C.cast(C.c_char_p(seq), C.c_void_p),
                seq_off.ctypes.data_as(u64p), n, int(calc_probabilities), res, seqpos.ctypes.data_as(u64p),
                sigpos.ctypes.data_as(u64p), prob.ctypes.data_as(f64p))
        if rc != 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        return self._unpack(n, res, seqpos, sigpos, prob, calc_probabilities, raise_errors)

    def _unpack(self, n, res, seqpos, sigpos, prob, calc_probabilities=True, raise_errors=False):
        out = []
        for i in range(n):
            r = res[i]
            if r.status != 0:
                e = RuntimeError(_message(self._lib, r.status, r.bad_char))
                if raise_errors:
                    raise e
                out.append(e)
                continue
            a, b = int(r.seg_offset), int(r.seg_offset + r.n_segments)
            ns = b - a
            out.append({
                "Z": r.Z,
                "sequence_positions": seqpos[a:b].copy(),
                "signal_positions": sigpos[a:b].copy(),
                "probabilities": prob[a:b].copy(),
                "states": ["M"] * ns,       # basic mode only emits match segments (NT:424-430)
                "polishes": [""] * ns,      # Segment.polish default (aligner.hpp:41)
            })
        return out

    def align(self, signal, sequence: str, calc_probabilities: bool = False) -> dict:
        """aligner_bindings.cpp:132-147 -> NTAligner::align (NT_aligner_api.cpp:230-312)."""
        sig = np.asarray(signal)
        if sig.ndim != 1:
            raise ValueError("Signal must be a one-dimensional array")
        if self._ntk:
            return self._ntk_align(sig, sequence, calc_probabilities)
        return self.align_batch([np.ascontiguousarray(sig, dtype=np.float64)], [sequence], calc_probabilities,
                                raise_errors=True)[0]

    def _ntk_align_batch(self, signals, sequences, calc_probabilities: bool, raise_errors: bool, concurrency: int = 0):
        """dyn_ntk_align_batch: independent reads on a pool of CUDA streams."""
        n = len(signals)
        sig, sig_off, seq, seq_off = self._pack(signals, sequences, np.float32)
        caps = (np.diff(sig_off.astype(np.int64)) + np.diff(seq_off.astype(np.int64)) + 16).astype(np.uint64)
        out_off = np.concatenate(([0], np.cumsum(caps))).astype(np.uint64)
        tot = int(out_off[-1])
        status = np.zeros(max(n, 1), dtype=np.int32)
        Z = np.zeros(max(n, 1))
        ns = np.zeros(max(n, 1), dtype=np.uint64)
        states = np.zeros(max(tot, 1), dtype=np.uint8)
        seqpos = np.zeros(max(tot, 1), dtype=np.uint64)
        sigpos = np.zeros(max(tot, 1), dtype=np.uint64)
        prob = np.zeros(max(tot, 1))
        pk = np.zeros(max(tot, 1), dtype=np.uint32)
        rc = self._lib.dyn_ntk_align_batch(
            self._h, C.c_void_p(sig.ctypes.data), sig_off.ctypes.data_as(u64p), C.cast(C.c_char_p(seq), C.c_void_p),
            seq_off.ctypes.data_as(u64p), n, int(calc_probabilities), C.c_void_p(status.ctypes.data), Z.ctypes.data_as(f64p),
            ns.ctypes.data_as(u64p), out_off.ctypes.data_as(u64p), C.c_void_p(states.ctypes.data), seqpos.ctypes.data_as(u64p),
            sigpos.ctypes.data_as(u64p), prob.ctypes.data_as(f64p), C.c_void_p(pk.ctypes.data), int(concurrency))
        if rc != 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        out = []
        for i in range(n):
            if status[i] != 0:
                msg = self._lib.dyn_status_message(int(status[i])).decode()
                if status[i] == 4:
                    msg += next((ch for ch in sequences[i] if ch not in "ACGTUacgtu"), "?")
                e = RuntimeError(msg)
                if raise_errors:
                    raise e
                out.append(e)
                continue
            a, b = int(out_off[i]), int(out_off[i]) + int(ns[i])
            out.append({"Z": float(Z[i]), "sequence_positions": seqpos[a:b].copy(), "signal_positions": sigpos[a:b].copy(),
                        "probabilities": prob[a:b].copy(), "states": [chr(c) for c in states[a:b]],
                        "polishes": [self._int_to_kmer(int(q)) for q in pk[a:b]]})
        return out

    def _int_to_kmer(self, q: int) -> str:
        """Aligner::intToKmer (aligner.cpp:222-239): base-4 digits, most significant first; reversed for RNA pores."""
        k = self.kmer_size
        s = "".join("ACGT"[(q >> (2 * (k - 1 - i))) & 3] for i in range(k))
        return s[::-1] if self.rna else s

    def _ntk_align(self, sig, sequence: str, calc_probabilities: bool) -> dict:
        """NTKAligner::align (NTK_aligner_api.cpp:881-927) through dyn_ntk_align: segments carry 'M' / 'P' states and
        the polished kmer."""
        x = np.ascontiguousarray(sig, dtype=np.float32)
        seq = sequence.encode("latin-1")
        cap = x.size + len(seq) + 16
        Z = C.c_double(0.0)
        ns = C.c_uint64(0)
        states = C.create_string_buffer(cap)
        seqpos = np.zeros(cap, dtype=np.uint64)
        sigpos = np.zeros(cap, dtype=np.uint64)
        prob = np.zeros(cap)
        pk = np.zeros(cap, dtype=np.uint32)
        rc = self._lib.dyn_ntk_align(self._h, C.c_void_p(x.ctypes.data if x.size else 0), x.size, seq, len(seq),
                                     int(calc_probabilities), C.byref(Z), C.byref(ns), states, seqpos.ctypes.data_as(u64p),
                                     sigpos.ctypes.data_as(u64p), prob.ctypes.data_as(f64p), C.c_void_p(pk.ctypes.data), cap)
        if rc < 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        if rc > 0:
            msg = self._lib.dyn_status_message(rc).decode()
            if rc == 4:  # DYN_INVALID_NT: the reference names the offending character (aligner.cpp:182)
                bad = next((ch for ch in sequence if ch not in "ACGTUacgtu"), "?")
                msg += bad
            raise RuntimeError(msg)
        n = int(ns.value)
        return {"Z": Z.value, "sequence_positions": seqpos[:n].copy(), "signal_positions": sigpos[:n].copy(),
                "probabilities": prob[:n].copy(), "states": [chr(c) for c in states.raw[:n]],
                "polishes": [self._int_to_kmer(int(q)) for q in pk[:n]]}

    # --------------------------------------------------------------------------------------------- train
    def train_batch(self, signals, sequences, per_read_model: bool = False):
        """Batched ``train``.  Returns (per-read list, pooled) where pooled = dict(w, x, xx, xi) holds the
        sufficient statistics summed over the successful reads of the batch (native kmer order)."""
        n = len(signals)
        sig, sig_off, seq, seq_off = self._pack(signals, sequences, np.float32)
        K = self.num_kmers
        res = (TrainResult * max(n, 1))()
        w, x, xx = np.zeros(K), np.zeros(K), np.zeros(K)
        xi = np.zeros(2)
        pm = ps = None
        if per_read_model:
            pm, ps = np.zeros((n, K)), np.zeros((n, K))
        rc = self._lib.dyn_train_batch(
            self._h, sig.ctypes.data, sig_off.ctypes.data_as(u64p), C.cast(C.c_char_p(seq), C.c_void_p),
            seq_off.ctypes.data_as(u64p), n, res, w.ctypes.data_as(f64p), x.ctypes.data_as(f64p),
            xx.ctypes.data_as(f64p), xi.ctypes.data_as(f64p),
            pm.ctypes.data_as(f64p) if per_read_model else None, ps.ctypes.data_as(f64p) if per_read_model else None)
        if rc != 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        out = []
        for i in range(n):
            r = res[i]
            if r.status != 0:
                out.append(RuntimeError(_message(self._lib, r.status, r.bad_char)))
                continue
            d = {"Z": r.Z, "transition_params": {"m1": r.m1, "e1": r.e1, "e2": r.e2}}
            if per_read_model:
                d["emission_model"] = {"mean": pm[i], "stdev": ps[i]}
            out.append(d)
        return out, {"w": w, "x": x, "xx": xx, "xi": xi}

    # ---- pooled training with device-resident statistics (dyn_train_accumulate / dyn_train_mstep_device) ----------
    def train_accumulate_packed(self, signal_ptr: int, sig_off, seq_ptr: int, seq_off, stats_ptr: int, device: bool = False):
        """Add the sufficient statistics of a batch to the DEVICE buffer at ``stats_ptr`` (3K + 4 doubles: w, x, xx,
        xi_m, xi_e, sum Z, reads ok).  Returns the per-read status codes (numpy int32)."""
        sig_off = np.ascontiguousarray(sig_off, dtype=np.uint64)
        seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
        n = sig_off.size - 1
        status = np.zeros(max(n, 1), dtype=np.int32)
        rc = self._lib.dyn_train_accumulate(self._h, C.c_void_p(signal_ptr), sig_off.ctypes.data_as(u64p), C.c_void_p(seq_ptr),
                                            seq_off.ctypes.data_as(u64p), n, int(device), C.c_void_p(stats_ptr),
                                            C.c_void_p(status.ctypes.data))
        if rc != 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        return status[:n]

    def train_accumulate(self, signals, sequences, stats_ptr: int):
        sig, sig_off, seq, seq_off = self._pack(signals, sequences, np.float32)
        seqbuf = np.frombuffer(seq, dtype=np.uint8)
        return self.train_accumulate_packed(sig.ctypes.data, sig_off, seqbuf.ctypes.data, seq_off, stats_ptr, device=False)

    def mstep_device(self, stats_ptr: int) -> dict:
        """M-step (NT:519-535, 703-722) on the pooled statistics at ``stats_ptr`` (device); the result becomes this
        handle's model.  Returns the re-estimated transitions."""
        t = np.zeros(3)
        rc = self._lib.dyn_train_mstep_device(self._h, C.c_void_p(stats_ptr), t.ctypes.data_as(f64p))
        if rc != 0:
            raise RuntimeError(self._lib.dyn_last_error(self._h).decode())
        return {"m1": float(t[0]), "e1": float(t[1]), "e2": float(t[2])}

    def train(self, signal, sequence: str, as_dicts: bool = True) -> dict:
        """aligner_bindings.cpp:149-163 -> NTAligner::train (NT_aligner_api.cpp:567-639).  ``emission_model`` is the
        reference's list of 4^k ``{"mean","stdev"}`` dicts (aligner_bindings.cpp:93-100) unless as_dicts=False."""
        sig = np.asarray(signal)
        if sig.ndim != 1:
            raise ValueError("Signal must be a one-dimensional array")
        out, _ = self.train_batch([sig], [sequence], per_read_model=True)
        r = out[0]
        if isinstance(r, Exception):
            raise r
        if as_dicts:
            m, s = r["emission_model"]["mean"], r["emission_model"]["stdev"]
            r["emission_model"] = [{"mean": float(a), "stdev": float(b)} for a, b in zip(m, s)]
        return r
