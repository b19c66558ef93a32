"""Batched front end around the hot path: what `dynamont-resquiggle` does per worker process and in its listener
(reference src/dynamont/segmentation/segment.py:69-176, 182-256), with BATCHES of reads on the GPU in place of one read per
worker process:

    jobs (raw file, shift, scale, start, end, sequence, read id, signal id)            segment.py:182-256  generate_jobs
      -> raw signal slice                                                              segment.py:144-148  get_signal
      -> (x - shift) / scale, Hampel filter           GPU: Aligner.preprocess_batch    segment.py:149-153
      -> RNA: reversed read, 9 x 'A' prefix                                            segment.py:155-158
      -> align(calc_probabilities=True)               GPU: Aligner.align_batch         segment.py:160-161
      -> CSV rows                                     C ABI: dyn_format_segments       segment.py:162-171, utils.py:193-232
      -> zstd-compressed CSV + .errors file                                            segment.py:69-107   listener

The ingest libraries of the reference (pysam for BAM, pod5, zstandard) are imported lazily: `jobs_from_bam`,
`Pod5SignalSource` and compressed output raise a clear ImportError where they are missing; everything else (any iterator of
job tuples, any callable that returns a raw signal, plain CSV output) has no dependency beyond numpy.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Callable, Iterable, Iterator, Optional, Tuple

import numpy as np

HEADER = b"readid,signalid,start,end,basepos,base,motif,state,posterior_probability,polish\n"  # segment.py:80
POLYA = "AAAAAAAAA"                                                                            # segment.py:157-158
Job = Tuple[str, float, float, int, int, str, str, str]  # rawFile, shift, scale, start, end, sequence, readid, signalid


class SegmentWriter:
    """The listener of segment.py:69-107: header + rows into `outfile` (zstd level 3 like the reference when `zstandard` is
    importable), error lines into `<outfile without its last two extensions>.errors`."""

    def __init__(self, outfile: str, compress="auto"):
        if os.path.isdir(outfile):
            outfile = os.path.join(outfile, "dynamont.csv")  # the reference's CLI help: a directory means dynamont.csv in it
        self.outfile = outfile
        self.errfile = os.path.splitext(os.path.splitext(outfile)[0])[0] + ".errors"  # segment.py:73
        self.n_rows = 0
        self.n_errors = 0
        zstd = None
        if compress in ("auto", True):
            try:
                import zstandard as zstd  # noqa: F811
            except ImportError:
                if compress is True:
                    raise ImportError("SegmentWriter(compress=True) needs the 'zstandard' package (the reference's output format); "
                                      "pass compress=False for plain CSV")
        self.compressed = zstd is not None
        self._raw = open(outfile, "wb")
        self._out = zstd.ZstdCompressor(level=3).stream_writer(self._raw) if zstd is not None else self._raw
        self._out.write(HEADER)

    def write(self, rows: bytes) -> None:
        self._out.write(rows)
        self.n_rows += 1

    def error(self, message: str) -> None:
        with open(self.errfile, "a") as err:
            err.write(message + "\n")
        self.n_errors += 1

    def close(self) -> None:
        if self._out is not self._raw:
            self._out.close()
        if not self._raw.closed:
            self._raw.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def jobs_from_bam(data_path: str, basecalls: str, min_qual: float = 0.0) -> Iterator[Job]:
    """generate_jobs (segment.py:182-256): one job per basecalled read of a BAM/SAM file, with the basecaller's tags
    (qs, pi, ns, ts, sp, fn|f5, sm, sd).  Needs pysam."""
    try:
        import pysam
    except ImportError as e:
        raise ImportError("jobs_from_bam needs 'pysam' (not in this image); any iterator of job tuples works in its place") from e
    with pysam.AlignmentFile(basecalls, "rb", check_sq=False) as sam:
        for rec in sam.fetch(until_eof=True):
            if min_qual and rec.get_tag("qs") < min_qual:
                continue
            readid = rec.query_name
            signalid = rec.get_tag("pi") if rec.has_tag("pi") else readid
            ns, ts = rec.get_tag("ns"), rec.get_tag("ts")
            sp = rec.get_tag("sp") if rec.has_tag("sp") else 0
            raw = os.path.join(data_path, rec.get_tag("fn") if rec.has_tag("fn") else rec.get_tag("f5"))
            yield (raw, rec.get_tag("sm"), rec.get_tag("sd"), sp + ts, sp + ns, rec.query_sequence, readid, signalid)


class Pod5SignalSource:
    """get_raw + get_signal of the reference (segment.py:109-139, pod5_io.py): an LRU of open POD5 readers (3, the files are
    more or less ordered).  Needs pod5.  Call it as source(raw_file, signalid, calibrated)."""

    def __init__(self, cache_size: int = 3):
        try:
            import pod5  # noqa: F401
        except ImportError as e:
            raise ImportError("Pod5SignalSource needs 'pod5' (not in this image); any callable (raw_file, signalid, calibrated) "
                              "-> 1-D array works in its place") from e
        self._pod5 = pod5
        self._cache: "OrderedDict[str, object]" = OrderedDict()
        self._size = cache_size

    def _reader(self, path: str):
        if path in self._cache:
            self._cache.move_to_end(path)
            return self._cache[path]
        if len(self._cache) >= self._size:
            _, old = self._cache.popitem(last=False)
            try:
                old.close()
            except Exception:
                pass
        self._cache[path] = self._pod5.Reader(path)
        return self._cache[path]

    def __call__(self, raw_file: str, signalid: str, calibrated: bool):
        rec = next(self._reader(raw_file).reads(selection=[signalid], missing_ok=False, preload={"samples"}))
        return rec.signal_pa if calibrated else rec.signal

    def close(self) -> None:
        while self._cache:
            _, r = self._cache.popitem(last=False)
            try:
                r.close()
            except Exception:
                pass


def prepare_read(read: str, is_rna: bool) -> str:
    """RNA reads are aligned 3'->5' with a polyA prefix (segment.py:155-158)."""
    if is_rna:
        read = read[::-1]
        if not read.startswith(POLYA):
            read = POLYA + read
    return read


def segment_jobs(aligner, jobs: Iterable[Job], get_raw_signal: Callable[[str, str, bool], np.ndarray], writer: SegmentWriter,
                 batch_reads: int = 4096, batch_samples: int = 1 << 28) -> dict:
    """The worker loop of segment.py:141-176 over batches: returns {"reads", "segmented", "errors"}.

    Per batch: raw slices [start:end] -> normalisation + Hampel on the GPU (dyn_preprocess_batch; float64 arithmetic like the
    numpy code, FP32 on the way into the DP) -> align_batch -> rows.  A read that fails leaves the reference's error line
    ("error: native, <message>\\tT: ..\\tN: ..\\tRid: ..\\tSid: ..") and never fails the batch."""
    counts = {"reads": 0, "segmented": 0, "errors": 0}
    is_rna = bool(aligner.rna)
    batch, n_samples = [], 0

    def flush():
        nonlocal batch, n_samples
        if not batch:
            return
        sigs = aligner.preprocess_batch([b[0] for b in batch], [b[1] for b in batch], [b[2] for b in batch])
        reads = [b[3] for b in batch]
        results = aligner.align_batch(sigs, reads, True)
        for (raw, _, _, read, readid, signalid, start), sig, res in zip(batch, sigs, results):
            counts["reads"] += 1
            if isinstance(res, Exception):
                writer.error(f"error: native, {res}\tT: {len(sig)}\tN: {len(read)}\tRid: {readid}\tSid: {signalid}")
                counts["errors"] += 1
                continue
            writer.write(aligner.format_segments(res, readid, signalid, start, len(sig) + start, read))
            counts["segmented"] += 1
        batch, n_samples = [], 0

    for job in jobs:
        raw_file, shift, scale, start, end, read, readid, signalid = job
        try:
            raw = np.asarray(get_raw_signal(raw_file, signalid, shift <= 400)[start:end])  # calibrated (pA) iff shift <= 400, segment.py:146
        except Exception as e:  # the reference's "error: worker" line (segment.py:178-187)
            writer.error(f"error: worker, {e}\tN: {len(read)}\tRid: {readid}\tSid: {signalid}")
            counts["reads"] += 1
            counts["errors"] += 1
            continue
        batch.append((raw, float(shift), float(scale), prepare_read(read, is_rna), readid, signalid, int(start)))
        n_samples += raw.size
        if len(batch) >= batch_reads or n_samples >= batch_samples:
            flush()
    flush()
    return counts


def segment(data_path: str, basecalls: str, outfile: str, model_path: str, pore: str, mode: str = "basic", minq: float = 0.0,
            device: int = 0, compress="auto", signal_source: Optional[Callable] = None, jobs: Optional[Iterable[Job]] = None) -> dict:
    """segment() of segment.py:259-340 on one GPU: BAM + POD5 in, (zstd) CSV + .errors out."""
    from .aligner import Aligner
    al = Aligner(model_path, pore, mode=mode, band=400, device=device)
    src = signal_source or Pod5SignalSource()
    try:
        with SegmentWriter(outfile, compress) as w:
            return segment_jobs(al, jobs if jobs is not None else jobs_from_bam(data_path, basecalls, minq), src, w)
    finally:
        if hasattr(src, "close"):
            src.close()
