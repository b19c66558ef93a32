"""The batched front end (dynamont_b200/frontend.py): the worker + listener of the reference's dynamont-resquiggle
(segment.py:69-176) over batches.  Runs on the CPU through the emulator build of the kernels; BAM / POD5 ingest needs
pysam / pod5 (not in this image) and is replaced by an in-memory job list and signal source."""
import os
import sys

import numpy as np
import pytest

from conftest import MODELS_DIR

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))


def _setup(lib, tmp_path, pore="rna002"):
    from dynamont_b200 import Aligner
    from dynamont_b200.synth import materialize_model, native_model, synth_read
    path = materialize_model("rna002_5mer", MODELS_DIR)
    al = Aligner(path, pore, _lib_path=lib)
    nm, ns = native_model(path, pore)
    rng = np.random.default_rng(31)
    store, jobs = {}, []
    for i, L in enumerate((60, 90, 45)):
        s, q, _ = synth_read(rng, nm, ns, 5, L, 9)
        # what the basecaller saw: the read 5'->3' without the polyA prefix the front end adds (segment.py:155-158)
        read = q[::-1] if al.rna else q
        shift, scale = 88.5 + i, 11.25
        raw = (s * scale + shift).astype(np.float32)
        raw[[20, 100]] += 150.0  # sensor spikes for the Hampel filter
        pad = 7 * (i + 1)
        store["sig-%d" % i] = np.concatenate([np.zeros(pad, np.float32), raw, np.zeros(5, np.float32)])
        jobs.append(("file.pod5", shift, scale, pad, pad + raw.size, read, "read-%d" % i, "sig-%d" % i))
    jobs.append(("file.pod5", 88.5, 11.25, 0, 8, "ACGTACGTACGTACGT", "read-short", "sig-0"))      # signal too short
    jobs.append(("file.pod5", 88.5, 11.25, 0, 100, "ACGTACGTACGT", "read-missing", "sig-none"))   # no such signal
    return al, store, jobs


def _run(lib, tmp_path, pore):
    from dynamont_b200 import frontend
    al, store, jobs = _setup(lib, tmp_path, pore)
    calls = []

    def source(raw_file, signalid, calibrated):
        calls.append((raw_file, signalid, calibrated))
        return store[signalid]

    out = str(tmp_path / ("seg_%s.csv.zst" % pore))
    with frontend.SegmentWriter(out, compress=False) as w:
        counts = frontend.segment_jobs(al, jobs, source, w, batch_reads=2)
    assert counts == {"reads": 5, "segmented": 3, "errors": 2}
    assert all(c[2] is True for c in calls)  # shift <= 400: calibrated (pA) signal, segment.py:146
    data = open(out, "rb").read()
    assert data.startswith(frontend.HEADER)
    # the same three reads one by one through the per-read API
    want = frontend.HEADER
    for raw_file, shift, scale, start, end, read, rid, sid in jobs[:3]:
        sig = al.preprocess_batch([store[sid][start:end]], [shift], [scale])[0]
        rd = frontend.prepare_read(read, al.rna)
        res = al.align(sig, rd, True)
        want += al.format_segments(res, rid, sid, start, len(sig) + start, rd)
    assert data == want
    errs = open(str(tmp_path / ("seg_%s.errors" % pore))).read().splitlines()
    assert len(errs) == 2
    assert errs[0].startswith("error: native, Signal too short compared to sequence") and errs[0].endswith("\tRid: read-short\tSid: sig-0")
    assert "\tT: 8\tN: " in errs[0]
    assert errs[1].startswith("error: worker, ") and errs[1].endswith("\tRid: read-missing\tSid: sig-none")
    return data


def test_frontend_driver_rna_and_dna(tmp_path):
    import build_emu
    from dynamont_b200 import frontend
    lib = build_emu.build()
    data = _run(lib, tmp_path, "rna002")
    # RNA: reversed + polyA-prefixed; every row carries the read id and 10 columns
    rows = data.decode().splitlines()[1:]
    assert rows and all(len(r.split(",")) == 10 for r in rows)
    assert frontend.prepare_read("CCGT", True) == "AAAAAAAAATGCC" and frontend.prepare_read("CCGT", False) == "CCGT"
    assert frontend.prepare_read("AAAAAAAAAC"[::-1], True) == "AAAAAAAAAC"
    _run(lib, tmp_path, "dna_r9")


def test_frontend_ingest_needs_its_libraries(tmp_path):
    """BAM / POD5 ingest and zstd output are the reference's own libraries: a clear ImportError where they are missing"""
    from dynamont_b200 import frontend
    for mod, call in (("pysam", lambda: next(frontend.jobs_from_bam(".", "x.bam"))), ("pod5", lambda: frontend.Pod5SignalSource()),
                      ("zstandard", lambda: frontend.SegmentWriter(str(tmp_path / "o.csv.zst"), compress=True))):
        try:
            __import__(mod)
        except ImportError:
            with pytest.raises(ImportError, match=mod):
                call()
    w = frontend.SegmentWriter(str(tmp_path), compress=False)
    w.close()
    assert w.outfile.endswith("dynamont.csv") and open(w.outfile, "rb").read() == frontend.HEADER


@pytest.mark.gpu
def test_frontend_driver_gpu(tmp_path):
    """the same driver on the CUDA library: batches through dyn_preprocess_batch + dyn_align_batch equal the per-read calls"""
    _run(None, tmp_path, "rna002")
    _run(None, tmp_path, "dna_r9")
