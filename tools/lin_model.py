"""Design-validation model (numpy, FP32) of the LINEAR-domain variant of the CUDA algorithm — NOT an oracle and not
part of the product.  Forward/backward values are plain FP32 probabilities scaled by one power of two per lane
(block floating point, integer exponents), so the recurrences need no log-sum-exp: one MUFU (the emission
2^s) per cell-update instead of two.  The model mirrors the kernel's arithmetic row by row (flush-to-zero,
renormalisation every R rows by exact powers of two, offset coupling to the source-side neighbour lane, forward
values expressed relative to Z - OB so that f*b IS the posterior, linear posterior-Viterbi with its own per-lane
exponents, sparse posterior records, row-mass fault detection) so that the numerical design can be evaluated
against the double-precision oracle without a GPU.
"""
from __future__ import annotations

import math

import numpy as np

from fp32_model import LOG2E, bounds, emis, pos_consts

f32 = np.float32
TINY = f32(2.0 ** -126)
KAPMAX = int(__import__("os").environ.get("LIN_KAPMAX", "100"))


def ftz(a):
    a = np.asarray(a, dtype=f32)
    a[np.abs(a) < TINY] = 0.0
    return a


def scale2(a, e):
    """a * 2^e in FP32 with flush-to-zero; e integer array or scalar (any magnitude)."""
    with np.errstate(over="ignore", under="ignore"):
        return ftz(np.ldexp(a.astype(f32), np.clip(e, -400, 400).astype(np.int32)).astype(f32))


def align_lin(x, kmers, mean, stdev, trans, k, band=400, R=4, RV=8, thr=2.0 ** -22, CPL=13, D=100, E0V=20,
              mass_tol=1e-3):
    x = np.asarray(x, dtype=f32)
    S = x.size
    Kc = kmers.size
    T, N = S + 1, Kc + 1
    lo, hi, bw = bounds(T, N, band)
    a, b, c, mu_hi = pos_consts(kmers, mean, stdev)
    m1 = f32(math.exp(trans[0]))
    e2 = f32(math.exp(trans[2]))
    lane = (np.arange(N) % (32 * CPL)) // CPL
    lb = np.nonzero(lane[1:] != lane[:-1])[0] + 1  # cells n whose left neighbour n-1 is in a different lane
    faults = []

    def pvec(t):
        s = emis(x[t], a, b, c, mu_hi)
        with np.errstate(under="ignore"):
            p = np.exp2(s.astype(f32)).astype(f32)
        out = np.zeros(N, dtype=f32)
        out[1:] = ftz(p)
        return out

    def mask(row, t):
        out = np.zeros(N, dtype=f32)
        out[lo[t]:hi[t]] = row[lo[t]:hi[t]]
        return out

    def lane_max(row, t):
        out = np.zeros(32, dtype=f32)
        np.maximum.at(out, lane[lo[t]:hi[t]], row[lo[t]:hi[t]])
        return out

    def renorm_exp(lm, O, src_shift, e0):
        """per-lane exponent after renormalisation: own maximum -> [2^e0, 2^(e0+1)); never more than D below the
        source-side neighbour's candidate; a lane without a non-zero cell sits D below the nearest live lane
        (ring distance, source side first)"""
        alive = lm > 0
        if not alive.any():
            return O
        ex = np.zeros(32, dtype=np.int64)
        ex[alive] = np.frexp(lm[alive])[1] - 1 - e0
        cand = np.where(alive, O + ex, np.iinfo(np.int64).min // 4)
        cs = np.roll(cand, src_shift)  # candidate of the source-side neighbour
        newO = np.where(alive, np.maximum(cand, cs - D), 0)
        step = -src_shift  # +1: source is the next lane (backward), -1: the previous lane (forward / Viterbi)
        for i in np.nonzero(~alive)[0]:
            for dist in range(1, 33):
                j = (i + step * dist) % 32
                if alive[j]:
                    break
                j = (i - step * dist) % 32
                if alive[j]:
                    break
            newO[i] = cand[j] - D
        return newO

    # ---------------- P1 backward (all rows kept only because this is a model)
    bM = np.zeros((T, N), dtype=f32)
    bE = np.zeros((T, N), dtype=f32)
    OBat = np.zeros((T, 32), dtype=np.int64)
    OB = np.full(32, -D, dtype=np.int64)
    OB[lane[N - 1]] = 0
    bE[T - 1, N - 1] = 1.0
    OBat[T - 1] = OB
    for t in range(T - 2, -1, -1):
        p = pvec(t)
        pm = ftz(p * m1)
        A = ftz(bM[t + 1] * pm)
        ext1 = np.zeros(N, dtype=f32)
        ext1[:-1] = A[1:]
        ext1[lb - 1] = scale2(A[lb], OB[lane[lb]] - OB[lane[lb - 1]])
        nm = ftz(bE[t + 1] * p)
        nm[0] = 0.0
        with np.errstate(over="ignore"):
            nE = ftz((nm.astype(np.float64) * np.float64(e2) + ext1.astype(np.float64)).astype(f32))
        nm = mask(nm, t)
        nE = mask(nE, t)
        if t % R == 0:
            lm = lane_max(nE, t)
            newO = renorm_exp(lm, OB, -1, 0)
            sh = (OB - newO)
            nm = scale2(nm, sh[lane])
            nE = scale2(nE, sh[lane])
            OB = newO
        bM[t], bE[t] = nm, nE
        OBat[t] = OB
    if not np.isfinite(bE[0, 0]) or bE[0, 0] <= 0:
        return {"fault": ["Zb"], "Z": float("nan")}
    Z2 = math.log2(float(bE[0, 0])) + float(OB[lane[0]])
    Zb = Z2 / LOG2E

    # ---------------- P2 forward with its OWN per-lane exponents OF (renormalised every R rows like the backward
    # pass); posterior = sf * sb * 2^(OF + OB - Z2).  Keeping the two directions independent is what makes range
    # losses visible: a cell one direction had to flush still carries weight in the other, so the row mass moves
    # away from 1.
    Z2i = math.floor(Z2)
    c0 = f32(2.0 ** -(Z2 - Z2i))
    fM = np.zeros(N, dtype=f32)
    fE = np.zeros(N, dtype=f32)
    fE[0] = 1.0
    OF = np.full(32, -D, dtype=np.int64)
    OF[lane[0]] = 0
    VM = np.zeros(N, dtype=f32)
    VE = np.zeros(N, dtype=f32)
    VE[0] = 1.0
    OV = np.full(32, -D, dtype=np.int64)
    OV[lane[0]] = 0
    bits = np.zeros((T, N), dtype=bool)
    rec_n, rec_M, rec_E = [None] * T, [None] * T, [None] * T
    mass_dev = 0.0
    guard = -1e9
    for t in range(1, T):
        p = pvec(t - 1)
        pm = ftz(p * m1)
        left = np.zeros(N, dtype=f32)
        left[1:] = fE[:-1]
        left[lb] = scale2(fE[lb - 1], OF[lane[lb - 1]] - OF[lane[lb]])
        with np.errstate(over="ignore", invalid="ignore"):
            nfM = ftz(left * pm)
            nfM[0] = 0.0
            nfE = ftz(ftz((fE.astype(np.float64) * np.float64(e2) + fM.astype(np.float64)).astype(f32)) * p)
        nfE[0] = 0.0
        fM = mask(nfM, t)
        fE = mask(nfE, t)
        if t % R == 0:
            lm = np.maximum(lane_max(fM, t), lane_max(fE, t))
            newO = renorm_exp(lm, OF, 1, 0)
            shf = OF - newO
            fM = scale2(fM, shf[lane])
            fE = scale2(fE, shf[lane])
            OF = newO
        kap = (c0 * np.ldexp(f32(1.0), np.clip(OF + OBat[t] - Z2i, -127, KAPMAX).astype(np.int32))).astype(f32)
        if t % R == 0:
            # guard: F_lane * B_lane / Z bounds the posterior any flushed cell of that lane could have had (times 2^-126)
            fl_ = np.maximum(lane_max(fM, t), lane_max(fE, t)); bl_ = lane_max(bE[t], t)
            okl = (fl_ > 0) & (bl_ > 0)
            if okl.any():
                g = (np.log2(fl_[okl].astype(np.float64)) + np.log2(bl_[okl].astype(np.float64)) + (OF + OBat[t])[okl] - Z2).max()
                guard = max(guard, g)
        kap[(OF + OBat[t] - Z2i) < -126] = 0.0
        with np.errstate(over="ignore", invalid="ignore"):
            PM = ftz(fM * ftz(bM[t] * kap[lane]))
            PE = ftz(fE * ftz(bE[t] * kap[lane]))
        # posterior-Viterbi (NT:357-362), linear
        vleft = np.zeros(N, dtype=f32)
        vleft[1:] = VE[:-1]
        vleft[lb] = scale2(VE[lb - 1], OV[lane[lb - 1]] - OV[lane[lb]])
        with np.errstate(over="ignore", invalid="ignore"):
            nVM = ftz(vleft * PM)
            nVM[0] = 0.0
            bits[t] = VM >= VE
            nVE = ftz(np.maximum(VM, VE) * PE)
        nVE[0] = 0.0
        VM = mask(nVM, t)
        VE = mask(nVE, t)
        if t % RV == 0:
            lm = np.maximum(lane_max(VM, t), lane_max(VE, t))
            newO = renorm_exp(lm, OV, 1, E0V)
            shv = OV - newO
            VM = scale2(VM, shv[lane])
            VE = scale2(VE, shv[lane])
            OV = newO
        # records: lanes holding a posterior above thr dump all their cells
        with np.errstate(invalid="ignore"):
            cellmax = np.maximum(PM, PE)
        lm = np.zeros(32, dtype=f32)
        np.maximum.at(lm, lane[lo[t]:hi[t]], np.nan_to_num(cellmax[lo[t]:hi[t]], nan=np.inf))
        hot = lm > f32(thr)
        sel = np.nonzero(hot[lane] & (np.arange(N) >= lo[t]) & (np.arange(N) < hi[t]))[0]
        rec_n[t], rec_M[t], rec_E[t] = sel, PM[sel], PE[sel]
        mass = float(PM[sel].astype(np.float64).sum() + PE[sel].astype(np.float64).sum())
        if not (abs(mass - 1.0) <= mass_tol):
            faults.append(("mass", t, mass))
        else:
            mass_dev = max(mass_dev, abs(mass - 1.0))
    Zf2 = (math.log2(float(fE[N - 1])) + float(OF[lane[N - 1]])) if fE[N - 1] > 0 and np.isfinite(fE[N - 1]) else float("nan")
    dz = 2.0 ** (Zf2 - Z2) if Zf2 == Zf2 else float("nan")
    if not (abs(dz - 1.0) < 2e-3):
        faults.append(("Zf", dz))

    def lookup(t, n, isM):
        r = np.nonzero(rec_n[t] == n)[0]
        if r.size == 0:
            return 0.0
        mass = rec_M[t].astype(np.float64).sum() + rec_E[t].astype(np.float64).sum()
        v = float(rec_M[t][r[0]] if isM else rec_E[t][r[0]])
        return v / mass if mass > 0 else 0.0

    t, n = T - 1, N - 1
    inM = False
    buf, seqpos, sigpos, prob = [], [], [], []
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
    if n != 0 or inM:
        faults.append(("traceback", t, n))
    nrec = sum(r.size for r in rec_n[1:])
    return {"Z": Zb, "dz": dz, "sequence_positions": np.array(seqpos[::-1], dtype=np.uint64),
            "signal_positions": np.array(sigpos[::-1], dtype=np.uint64), "probabilities": np.array(prob[::-1]),
            "records_per_row": nrec / max(T - 1, 1) / CPL, "fault": faults, "mass_dev": mass_dev, "guard": guard}


if __name__ == "__main__":
    import os
    import sys
    import time

    ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden  # noqa: E402
    from dynamont_b200.synth import PORE_INFO, encode_kmers, native_model  # noqa: E402
    from oracle import PORES  # noqa: E402

    code = {c: i for i, c in enumerate("ACGT")}
    code["U"] = 3
    for g in load_golden():
        nm, ns = native_model(g.model_path, g.pore)
        rna, k = PORE_INFO[g.pore]
        digs = np.array([code[ch] for ch in g.sequence.upper()])
        km = encode_kmers(digs, k)
        tr = [math.log(v) for v in PORES[g.pore][2]]
        t0 = time.time()
        r = align_lin(g.signal, km, nm, ns, tr, k)
        if "signal_positions" not in r:
            print(g.name, "FAULT", r["fault"])
            continue
        same = r["signal_positions"] == g.signal_positions if r["signal_positions"].size == g.signal_positions.size else np.zeros(1, bool)
        ok = same.copy()
        if ok.size > 1:
            ok[:-1] &= same[1:]
        dp = np.abs(r["probabilities"] - g.probabilities)[ok].max(initial=0.0) if same.size == g.probabilities.size else -1
        print("%-20s borders %d/%d  max|dp| %.2e  dZrel %.2e  dz %.2e  massdev %.1e guard %.0f rec/row %.2f faults %d  %.1fs" % (
            g.name, same.sum(), same.size, dp, abs(r["Z"] - g.Z) / abs(g.Z), r["dz"] - 1, r["mass_dev"], r["guard"],
            r["records_per_row"], len(r["fault"]), time.time() - t0), r["fault"][:3])
