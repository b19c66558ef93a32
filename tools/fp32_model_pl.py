"""Design-validation model, variant with PER-LANE (block floating point) normalisation: cell n lives in
ring slot q = n mod (32*CPL), lane = q // CPL; each lane carries its own double offset.  See fp32_model.py."""
from __future__ import annotations

import math

import numpy as np

from fp32_model import LOG2E, bounds, emis, logplus2, pos_consts

f32 = np.float32
NEG = f32(-1.0e30)
NEGT = f32(-1.0e29)


def align_fp32_pl(x, kmers, mean, stdev, trans, k, band=400, R=4, RV=8, thr2=-22.0, CPL=13, mass_corr=True,
                  return_internal=False):
    x = np.asarray(x, dtype=f32)
    S = x.size
    Kc = kmers.size
    T, N = S + 1, Kc + 1
    lo, hi, bw = bounds(T, N, band)
    a, b, c, mu_hi = pos_consts(kmers, mean, stdev)
    m1 = f32(trans[0] * LOG2E)
    e2 = f32(trans[2] * LOG2E)
    lane = (np.arange(N) % (32 * CPL)) // CPL  # lane of cell n
    # boundary cells: left neighbour (n-1) lives in another lane
    lb = np.nonzero(lane[1:] != lane[:-1])[0] + 1  # cells n whose left neighbour is in a different lane

    def mask(row, t):
        out = np.full(N, NEG, dtype=f32)
        out[lo[t]:hi[t]] = row[lo[t]:hi[t]]
        return out

    def lane_max(row, t):
        """per-lane max over in-band cells -> array[32] (NEG where the lane has no finite cell)"""
        out = np.full(32, NEG, dtype=f32)
        seg = row[lo[t]:hi[t]]
        np.maximum.at(out, lane[lo[t]:hi[t]], seg)
        return out

    # ---------------- P1 backward
    bM = np.full((T, N), NEG, dtype=f32)
    bE = np.full((T, N), NEG, dtype=f32)
    incB = np.zeros((T, 32), dtype=f32)
    OBrow = np.zeros((T, 32))
    OB = np.zeros(32)
    bE[T - 1, N - 1] = 0.0
    pending = np.zeros(32, dtype=f32)
    for t in range(T - 2, -1, -1):
        s = np.full(N, NEG, dtype=f32)
        s[1:] = emis(x[t], a, b, c, mu_hi)
        A = (bM[t + 1] + (s + m1)).astype(f32)  # in lane(n) units
        # transfer A[n] -> cell n-1: across lane boundary add (OB[lane(n)] - OB[lane(n-1)])
        ext1 = np.full(N, NEG, dtype=f32)
        ext1[:-1] = A[1:]
        d = (OB[lane[lb]] - OB[lane[lb - 1]]).astype(f32)
        ext1[lb - 1] = (A[lb] + d).astype(f32)
        newM = (bE[t + 1] + s).astype(f32)
        newM[0] = NEG
        ext2 = (newM + e2).astype(f32)
        newE = logplus2(ext1, ext2)
        newM = mask(newM, t)
        newE = mask(newE, t)
        if pending.any():
            newM = (newM - pending[lane]).astype(f32)
            newE = (newE - pending[lane]).astype(f32)
            incB[t] = pending
            OB = OB + pending.astype(np.float64)
            pending = np.zeros(32, dtype=f32)
        newM = np.maximum(newM, NEG)
        newE = np.maximum(newE, NEG)
        bM[t], bE[t] = newM, newE
        if t % R == 0 and t > 0:
            lm = lane_max(newE, t)
            dead = lm < NEGT
            pending = np.where(dead, f32(0), lm).astype(f32)
            # dead lanes adopt the offset of their right neighbour (source side for backward) after the update
            OBn = OB + pending.astype(np.float64)
            adopt = np.roll(OBn, -1)
            # express adoption as an extra pending increment (values are all NEG there, so it is harmless)
            OB = np.where(dead, adopt - 0.0, OB)
        OBrow[t] = OB + pending.astype(np.float64)  # offset in force for row t AFTER pending is applied at t-1? (kept for debug)
    l0 = lane[0]
    Zb2 = float(bE[0, 0]) + OB[l0]
    Zb = Zb2 / LOG2E
    # offsets in force at each row (needed by forward): recompute exactly
    # OBat[t] = offset vector that row t's stored values are relative to
    OBat = np.zeros((T, 32))
    OBat[T - 1] = 0.0
    # replay: simpler to recompute by rerunning the bookkeeping
    # (done below by a second light pass)
    OB = np.zeros(32)
    pending = np.zeros(32, dtype=f32)
    OBat[T - 1] = OB
    for t in range(T - 2, -1, -1):
        if pending.any():
            OB = OB + pending.astype(np.float64)
            pending = np.zeros(32, dtype=f32)
        OBat[t] = OB
        if t % R == 0 and t > 0:
            lm = lane_max(bE[t], t)
            dead = lm < NEGT
            pending = np.where(dead, f32(0), lm).astype(f32)
            OBn = OB + pending.astype(np.float64)
            OB = np.where(dead, np.roll(OBn, -1), OB)
            OBat[t] = np.where(dead, OB, OBat[t])  # dead lanes: values are NEG, any offset is fine

    # ---------------- P2 forward normalised by OF = Zb2 - OB
    fM = np.full(N, NEG, dtype=f32)
    fE = np.full(N, NEG, dtype=f32)
    fE[0] = -bE[0, 0]
    VM = np.full(N, NEG, dtype=f32)
    VE = np.full(N, NEG, dtype=f32)
    VE[0] = 0.0
    OV = np.zeros(32)
    vpend = np.zeros(32, dtype=f32)
    bits = np.zeros((T, N), dtype=bool)
    rec_n, rec_M, rec_E = [None] * T, [None] * T, [None] * T
    for t in range(1, T):
        s = np.full(N, NEG, dtype=f32)
        s[1:] = emis(x[t - 1], a, b, c, mu_hi)
        # forward values at row t-1 are relative to OF(t-1) = Z - OBat[t-1]; row t must be relative to Z - OBat[t]
        # step: value_new_true = rec(value_prev_true); repr_new = true - OF_t[lane]
        src = fE.copy()
        left = np.full(N, NEG, dtype=f32)
        left[1:] = src[:-1]
        dl = (OBat[t - 1][lane[lb]] - OBat[t - 1][lane[lb - 1]]).astype(f32)  # OF_left - OF_me = OB_me - OB_left
        left[lb] = (src[lb - 1] + dl).astype(f32)
        nfM = (left + (s + m1)).astype(f32)
        nfM[0] = NEG
        nfE = (logplus2(fM, (fE + e2).astype(f32)) + s).astype(f32)
        nfE[0] = NEG
        shift = (OBat[t] - OBat[t - 1]).astype(f32)  # OF_{t-1} - OF_t  = OB_t - OB_{t-1}
        if shift.any():
            nfM = (nfM + shift[lane]).astype(f32)
            nfE = (nfE + shift[lane]).astype(f32)
        fM = np.maximum(mask(nfM, t), NEG)
        fE = np.maximum(mask(nfE, t), NEG)
        LPM = (fM + bM[t]).astype(f32)
        LPE = (fE + bE[t]).astype(f32)
        # posterior-Viterbi, per-lane offsets OV
        vleft = np.full(N, NEG, dtype=f32)
        vleft[1:] = VE[:-1]
        dv = (OV[lane[lb - 1]] - OV[lane[lb]]).astype(f32)
        vleft[lb] = (VE[lb - 1] + dv).astype(f32)
        nVM = (vleft + LPM).astype(f32)
        nVM[0] = NEG
        av = (VM + LPE).astype(f32)
        bv = (VE + LPE).astype(f32)
        nVE = np.maximum(av, bv)
        bits[t] = av >= bv
        nVE[0] = NEG
        VM = np.maximum(mask(nVM, t), NEG)
        VE = np.maximum(mask(nVE, t), NEG)
        if vpend.any():
            VM = np.maximum((VM - vpend[lane]).astype(f32), NEG)
            VE = np.maximum((VE - vpend[lane]).astype(f32), NEG)
            OV = OV + vpend.astype(np.float64)
            vpend = np.zeros(32, dtype=f32)
        if t % RV == 0:
            lm = np.maximum(lane_max(VE, t), lane_max(VM, t))
            dead = lm < NEGT
            vpend = np.where(dead, f32(0), lm).astype(f32)
            OV = np.where(dead, np.roll(OV + vpend.astype(np.float64), 1), OV)  # adopt left neighbour
        sel = np.nonzero(np.maximum(LPM, LPE) > f32(thr2))[0]
        rec_n[t], rec_M[t], rec_E[t] = sel, LPM[sel], LPE[sel]
    dZ2 = float(fE[N - 1]) + float(bE[T - 1, N - 1])

    def lookup(t, n, isM):
        r = np.nonzero(rec_n[t] == n)[0]
        if r.size == 0:
            return 0.0
        cm = 0.0
        if mass_corr:
            mass = np.exp2(rec_M[t].astype(np.float64)).sum() + np.exp2(rec_E[t].astype(np.float64)).sum()
            cm = math.log2(mass)
        lp = float(rec_M[t][r[0]] if isM else rec_E[t][r[0]])
        return 2.0 ** (lp - cm)

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
    nrec = sum(r.size for r in rec_n[1:])
    return {"Z": Zb, "dZ2": dZ2, "sequence_positions": np.array(seqpos[::-1], dtype=np.uint64),
            "signal_positions": np.array(sigpos[::-1], dtype=np.uint64), "probabilities": np.array(prob[::-1]),
            "records_per_row": nrec / max(T - 1, 1)}
