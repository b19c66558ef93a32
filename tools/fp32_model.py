"""Design-validation model (numpy, FP32) of the CUDA algorithm in dynamont_b200/csrc — NOT an oracle and not
part of the product.  It mirrors, row by row, the arithmetic the kernels perform (log2 domain, FP32 state,
renormalisation every R rows with the offsets accumulated in double, forward pass normalised by the backward
pass's offsets, posterior-Viterbi with periodic shift, sparse posterior records with per-row mass correction)
so that numerical design choices can be evaluated against the double-precision oracle without a GPU.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32
NEG = f32(-1.0e30)
LOG2E = 1.4426950408889634


PRECISE = True


def pos_consts(kmers, mean, stdev):
    """Per-position emission constants (log2 domain).
    fast:    s(x) = c - (x*a - b)^2                    with b = mu*a
    precise: s(x) = c - ((x - mu_hi)*a - b)^2         with b = (mu - mu_hi)*a"""
    sd = stdev[kmers]
    mu = mean[kmers]
    a = np.sqrt(0.5 * LOG2E) / sd
    c = -np.log2(sd) - 0.5 * math.log2(2.0 * math.pi)
    if PRECISE:
        mu_hi = mu.astype(f32)
        b = (mu - mu_hi.astype(np.float64)) * a
        return a.astype(f32), b.astype(f32), c.astype(f32), mu_hi
    b = mu * a
    return a.astype(f32), b.astype(f32), c.astype(f32), None


def emis(x, a, b, c, mu_hi=None):
    if mu_hi is not None:
        d = (f32(x) - mu_hi).astype(f32)
        z = (d.astype(np.float64) * a.astype(np.float64) - b.astype(np.float64)).astype(f32)  # fma
    else:
        z = (np.float64(x) * a.astype(np.float64) - b.astype(np.float64)).astype(f32)  # fma
    return (c.astype(np.float64) - z.astype(np.float64) ** 2).astype(f32)  # fma


def logplus2(u, v):
    mx = np.maximum(u, v)
    d = -np.abs(u - v)
    with np.errstate(over="ignore"):
        e = np.exp2(d.astype(f32)).astype(f32)
    return (mx + np.log2((f32(1.0) + e).astype(f32)).astype(f32)).astype(f32)


def bounds(T, N, band):
    bw = min(band // 2, N // 2)
    ratio = float(N) / float(T)
    mid = (np.arange(T, dtype=np.float64) * ratio).astype(np.int64)
    lo = np.maximum(mid - bw, 0)
    hi = np.minimum(mid + bw + 1, N)
    return lo, hi, bw


def align_fp32(x, kmers, mean, stdev, trans, k, band=400, R=1, RV=8, thr2=-22.0, exact_eq=True, mass_corr=True):
    x = np.asarray(x, dtype=f32)
    S = x.size
    Kc = kmers.size
    T, N = S + 1, Kc + 1
    lo, hi, bw = bounds(T, N, band)
    a, b, c, mu_hi = pos_consts(kmers, mean, stdev)  # index j-1 for cell n=j
    m1 = f32(trans[0] * LOG2E)
    e2 = f32(trans[2] * LOG2E)
    n_idx = np.arange(N)

    def mask(row, t):
        out = np.full(N, NEG, dtype=f32)
        out[lo[t]:hi[t]] = row[lo[t]:hi[t]]
        return out

    # ---- P1: backward, all rows kept here only because this is a model
    bM = np.full((T, N), NEG, dtype=f32)
    bE = np.full((T, N), NEG, dtype=f32)
    inc = np.zeros(T, dtype=f32)  # increment subtracted from row t's values
    bE[T - 1, N - 1] = 0.0
    pending = f32(0.0)
    OB = 0.0
    OBrow = np.zeros(T)
    for t in range(T - 2, -1, -1):
        s = np.full(N, NEG, dtype=f32)
        s[1:] = emis(x[t], a, b, c, mu_hi)
        A = (bM[t + 1] + (s + m1)).astype(f32)
        newM = (bE[t + 1] + s).astype(f32)
        newM[0] = NEG
        ext2 = (newM + e2).astype(f32)
        ext1 = np.full(N, NEG, dtype=f32)
        ext1[:-1] = A[1:]
        newE = logplus2(ext1, ext2)
        newM = mask(newM, t)
        newE = mask(newE, t)
        if pending != 0.0:
            newM = (newM - pending).astype(f32)
            newE = (newE - pending).astype(f32)
            inc[t] = pending
            OB += float(pending)
            pending = f32(0.0)
        newM = np.maximum(newM, NEG)
        newE = np.maximum(newE, NEG)
        bM[t], bE[t] = newM, newE
        OBrow[t] = OB
        if t % R == 0 and t > 0:
            pending = f32(newE[lo[t]:hi[t]].max())
    Zb2 = float(bE[0, 0]) + OB
    Zb = Zb2 / LOG2E

    # ---- P2: forward normalised by the backward offsets, LP, posterior-Viterbi, decisions, sparse records
    fM = np.full(N, NEG, dtype=f32)
    fE = np.full(N, NEG, dtype=f32)
    fE[0] = -bE[0, 0]
    VM = np.full(N, NEG, dtype=f32)
    VE = np.full(N, NEG, dtype=f32)
    VE[0] = 0.0
    bits = np.zeros((T, N), dtype=bool)
    rec_n, rec_M, rec_E = [None] * T, [None] * T, [None] * T
    vshift = f32(0.0)
    for t in range(1, T):
        s = np.full(N, NEG, dtype=f32)
        s[1:] = emis(x[t - 1], a, b, c, mu_hi)
        nfM = np.full(N, NEG, dtype=f32)
        nfM[1:] = (fE[:-1] + (s[1:] + m1)).astype(f32)
        nfE = (logplus2(fM, (fE + e2).astype(f32)) + s).astype(f32)
        nfE[0] = NEG
        if inc[t - 1] != 0.0:
            nfM = (nfM - inc[t - 1]).astype(f32)
            nfE = (nfE - inc[t - 1]).astype(f32)
        fM = np.maximum(mask(nfM, t), NEG)
        fE = np.maximum(mask(nfE, t), NEG)
        LPM = (fM + bM[t]).astype(f32)
        LPE = (fE + bE[t]).astype(f32)
        nVM = np.full(N, NEG, dtype=f32)
        nVM[1:] = (VE[:-1] + LPM[1:]).astype(f32)
        if exact_eq:
            av = (VM + LPE).astype(f32)
            bv = (VE + LPE).astype(f32)
            nVE = np.maximum(av, bv)
            bits[t] = av >= bv
        else:
            nVE = (np.maximum(VM, VE) + LPE).astype(f32)
            bits[t] = VM >= VE
        nVE[0] = NEG
        VM = np.maximum(mask(nVM, t), NEG)
        VE = np.maximum(mask(nVE, t), NEG)
        if vshift != 0.0:
            VM = np.maximum((VM - vshift).astype(f32), NEG)
            VE = np.maximum((VE - vshift).astype(f32), NEG)
            vshift = f32(0.0)
        if t % RV == 0:
            vshift = f32(max(VE[lo[t]:hi[t]].max(), VM[lo[t]:hi[t]].max()))
        sel = np.nonzero(np.maximum(LPM, LPE) > f32(thr2))[0]
        rec_n[t], rec_M[t], rec_E[t] = sel, LPM[sel], LPE[sel]
    Zf_minus_Zb2 = float(fE[N - 1]) + float(bE[T - 1, N - 1])

    # ---- P3: traceback + per-segment median of corrected posteriors
    def lookup(t, n, isM):
        r = np.nonzero(rec_n[t] == n)[0]
        if r.size == 0:
            return 0.0
        if mass_corr:
            mass = np.exp2(rec_M[t].astype(np.float64)).sum() + np.exp2(rec_E[t].astype(np.float64)).sum()
            cm = math.log2(mass)
        else:
            cm = 0.0
        lp = float(rec_M[t][r[0]] if isM else rec_E[t][r[0]])
        return 2.0 ** (lp - cm)

    t, n = T - 1, N - 1
    inM = False
    buf = []
    seqpos, sigpos, prob = [], [], []
    while t and n:
        if inM:
            buf.append(lookup(t, n, True))
            seqpos.append(n - 1 + k // 2)
            sigpos.append(t - 1)
            prob.append(float(np.median(buf)))
            buf = []
            t -= 1
            n -= 1
            inM = False
        else:
            buf.append(lookup(t, n, False))
            inM = bool(bits[t, n])
            t -= 1
    nrec = sum(r.size for r in rec_n[1:])
    return {"Z": Zb, "dZ2": Zf_minus_Zb2, "sequence_positions": np.array(seqpos[::-1], dtype=np.uint64),
            "signal_positions": np.array(sigpos[::-1], dtype=np.uint64), "probabilities": np.array(prob[::-1]),
            "records_per_row": nrec / max(T - 1, 1)}
