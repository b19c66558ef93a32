"""Deterministic synthetic reads and pore models (SURVEY.md §8d recipe).

Used by tests and bench.py; numpy only.  Signals are rounded to FP32 and widened, so the reference
(double) and the CUDA path (float) see identical sample values.
"""
from __future__ import annotations

import os

import numpy as np

BASES = "ACGT"
PORE_INFO = {  # pore -> (rna, k)   (reference aligner.cpp:62-86)
    "rna002": (True, 5),
    "rna004": (True, 9),
    "dna_r9": (False, 5),
    "dna_r10_260bps": (False, 9),
    "dna_r10_400bps": (False, 9),
}


def read_kmer_model(path: str):
    """TSV ``kmer\\tlevel_mean\\tlevel_stdv`` -> (list of kmers in file order, mean[], stdev[])."""
    kmers, mean, sd = [], [], []
    with open(path) as fh:
        next(fh)
        for line in fh:
            line = line.rstrip("\n")
            if not line:
                continue
            f = line.split("\t")
            kmers.append(f[0])
            mean.append(float(f[1]))
            sd.append(float(f[2]))
    return kmers, np.array(mean), np.array(sd)


def write_kmer_model(path: str, kmers, mean, sd) -> None:
    """Same text format as the reference's utils.write_kmer_model (utils.py:136-152): str(float)."""
    with open(path, "w") as fh:
        fh.write("kmer\tlevel_mean\tlevel_stdv\n")
        for q, m, s in zip(kmers, mean, sd):
            fh.write(f"{q}\t{float(m)}\t{float(s)}\n")


def kmer_strings(k: int):
    """All 4^k kmers in lexicographic ACGT order."""
    idx = np.arange(4 ** k)
    chars = np.empty((4 ** k, k), dtype="U1")
    for i in range(k):
        chars[:, k - 1 - i] = np.array(list(BASES))[(idx >> (2 * i)) & 3]
    return ["".join(r) for r in chars]


def make_synthetic_9mer_model(path_5mer: str, out_path: str, seed: int = 42, stdv: float = 0.15, var_sd: bool = False) -> str:
    """level_mean(kmer9) = model5[kmer9[2:7]] + 0.05*N(0,1); level_stdv = 0.15 (inverse of models/9merTo5mer.py).
    var_sd: per-kmer level_stdv drawn uniformly from [0.10, 0.30] (the spread of the shipped trained 5-mer model) — what
    a trained 9-mer model looks like to the kernels."""
    if os.path.exists(out_path):
        return out_path
    k5, m5, _ = read_kmer_model(path_5mer)
    lut = np.zeros(4 ** 5)
    code = {c: i for i, c in enumerate(BASES)}
    for q, m in zip(k5, m5):
        v = 0
        for ch in q:
            v = v * 4 + code[ch]
        lut[v] = m
    idx9 = np.arange(4 ** 9)
    central = (idx9 >> 4) & (4 ** 5 - 1)  # digits 2..6 of a 9-digit base-4 number
    rng = np.random.default_rng(seed)
    mean9 = lut[central] + 0.05 * rng.standard_normal(4 ** 9)
    sd9 = rng.uniform(0.10, 0.30, 4 ** 9) if var_sd else np.full(4 ** 9, stdv)
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    tmp = out_path + ".tmp%d" % os.getpid()
    kmers = kmer_strings(9)
    with open(tmp, "w") as fh:
        fh.write("kmer\tlevel_mean\tlevel_stdv\n")
        fh.write("".join(f"{q}\t{float(m)}\t{float(v)}\n" for q, m, v in zip(kmers, mean9, sd9)))
    os.replace(tmp, out_path)
    return out_path


def native_model(path: str, pore: str):
    """(mean[K], stdev[K]) in the aligner's native index order: index = base-4 value of the kmer as it
    appears in the (signal-oriented) sequence, i.e. of the REVERSED file kmer for RNA pores."""
    rna, k = PORE_INFO[pore]
    kmers, mean, sd = read_kmer_model(path)
    code = np.full(256, -1)
    for i, c in enumerate(BASES):
        code[ord(c)] = i
    code[ord("U")] = 3
    arr = np.frombuffer("".join(kmers).encode(), dtype=np.uint8).reshape(len(kmers), k)
    digits = code[arr]
    if rna:
        digits = digits[:, ::-1]
    idx = np.zeros(len(kmers), dtype=np.int64)
    for i in range(k):
        idx = idx * 4 + digits[:, i]
    nm = np.zeros(4 ** k)
    ns = np.ones(4 ** k)
    nm[idx] = mean
    ns[idx] = sd
    return nm, ns


def encode_kmers(seq_digits: np.ndarray, k: int) -> np.ndarray:
    Kc = seq_digits.size - k + 1
    v = np.zeros(Kc, dtype=np.int64)
    for i in range(k):
        v = v * 4 + seq_digits[i:i + Kc]
    return v


def synth_read(rng, nmean, nstdev, k: int, length: int, spb: float, dwell: str = "geometric",
               prefix_a: bool = True, sd_scale: float = 1.0, seq_digits=None):
    """One synthetic read: returns (signal float64 holding FP32-representable values, sequence str,
    true border array).  ``sequence`` is in signal orientation (what Aligner.align receives)."""
    if seq_digits is None:
        seq_digits = rng.integers(0, 4, size=length)
        if prefix_a:
            seq_digits[:k] = 0
    seq_digits = np.asarray(seq_digits, dtype=np.int64)
    kmers = encode_kmers(seq_digits, k)
    Kc = kmers.size
    if dwell == "geometric":
        d = np.maximum(2, rng.geometric(1.0 / spb, size=Kc))
    elif dwell == "gamma":
        d = np.maximum(2, np.rint(rng.gamma(4.0, spb / 4.0, size=Kc))).astype(np.int64)
    else:
        raise ValueError(dwell)
    per_sample = np.repeat(kmers, d)
    sig = nmean[per_sample] + sd_scale * nstdev[per_sample] * rng.standard_normal(per_sample.size)
    sig = sig.astype(np.float32).astype(np.float64)
    seq = "".join(BASES[i] for i in seq_digits)
    borders = np.concatenate(([0], np.cumsum(d)[:-1]))
    return sig, seq, borders


def low_complexity_digits(rng, length: int, kind: str, k: int):
    """Homopolymer / dinucleotide-repeat stress sequences."""
    if kind == "homopolymer":
        runs = []
        while sum(len(r) for r in runs) < length:
            runs.append([int(rng.integers(0, 4))] * int(rng.integers(4, 20)))
        d = np.array([b for r in runs for b in r][:length])
    elif kind == "dinuc":
        a, b = rng.choice(4, size=2, replace=False)
        d = np.tile([a, b], length // 2 + 1)[:length]
    elif kind == "mixed":
        d = rng.integers(0, 4, size=length)
        i = 0
        while i < length:
            run = int(rng.integers(5, 30))
            if rng.random() < 0.5:
                d[i:i + run] = rng.integers(0, 4)
            i += run + int(rng.integers(0, 20))
    else:
        raise ValueError(kind)
    d = np.asarray(d, dtype=np.int64)
    d[:k] = 0
    return d


_TABLES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "models", "pore_tables.npz")


def materialize_model(name: str, outdir: str) -> str:
    """Write ``<outdir>/<name>.model`` for one of the packed 5-mer tables (rna002_5mer,
    trained_rna002_5mer, rna004_5mer) or the synthetic 9-mer model ``synthetic_rna004_9mer``
    (the shipped tree has no 9-mer model).  Returns the path."""
    os.makedirs(outdir, exist_ok=True)
    path = os.path.join(outdir, name + ".model")
    if os.path.exists(path):
        return path
    if name == "synthetic_rna004_9mer":
        return make_synthetic_9mer_model(materialize_model("rna004_5mer", outdir), path)
    if name == "synthetic_rna004_9mer_varsd":
        return make_synthetic_9mer_model(materialize_model("rna004_5mer", outdir), path, var_sd=True)
    with np.load(_TABLES) as z:
        mean, sd = z[name + "_mean"], z[name + "_stdv"]
    tmp = path + ".tmp%d" % os.getpid()
    write_kmer_model(tmp, kmer_strings(5), mean, sd)
    os.replace(tmp, path)
    return path
