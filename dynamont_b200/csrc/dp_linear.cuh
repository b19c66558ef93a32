// Linear-domain variant of the warp-per-read banded forward/backward/posterior-decoding kernels (dp_kernels.cuh).
//
// Same lattice, same ring-slot layout, same three passes and the same scratch buffers as the log2-domain kernels,
// but the forward / backward / posterior-Viterbi values are plain FP32 probabilities scaled by ONE POWER OF TWO PER
// LANE (block floating point with integer exponents).  The recurrences of NT_aligner_api.cpp:110-207 then need no
// log-sum-exp:
//     backward   bM[t][n] = bE[t+1][n] * p(t,n)            bE[t][n] = bM[t+1][n+1] * p(t,n+1) * m1 + bM[t][n] * e2
//     forward    fM[t+1][n] = fE[t][n-1] * p(t,n) * m1     fE[t+1][n] = (fM[t][n] + fE[t][n] * e2) * p(t,n)
// with p(t,n) = N(x[t]; kmer[n-1]) = 2^(c - z^2): one MUFU (ex2) per cell-update instead of two, ~7 FP32 instructions
// instead of ~13, and better accuracy (sums of positive terms, no cancellation; renormalisation by exact powers of
// two).  Posterior-Viterbi (NT:338-363) is a max-PRODUCT recurrence on the posteriors themselves.
//
// Range management:
//   * every RN rows a lane renormalises its cells so that its largest value lies in [1, 2); the exponent goes into
//     the lane's integer offset OB.  A lane's offset is never more than DCPL below its source-side neighbour's, and a
//     lane without a non-zero cell sits DCPL below the nearest live lane, so a value handed to a neighbour lane can
//     never overflow.  Cells more than ~2^-126 below their lane's maximum flush to zero; they carry posteriors far
//     below anything the outputs can resolve.
//   * the forward pass renormalises the same way with its OWN offsets OF (source side = the left lane); the
//     posterior of a cell is sf * sb * 2^(OF + OB - Z2), the per-lane factor being folded into the backward operand.
//     Keeping the two directions independent makes range losses visible: a cell one direction had to flush still
//     carries weight in the other, so the recorded row mass moves away from 1.
//   * what FP32 cannot represent is detected, not approximated:
//       - every row's recorded posterior mass must be 1 within LIN_MASS_TOL, Zf must equal Zb, nothing may be NaN/inf;
//       - guard: F_lane(t) * B_lane(t) / Z (largest forward value times largest backward value of a lane) bounds, times
//         2^-126, the posterior any flushed cell of that lane can have had; it must stay below 2^LIN_GUARD_BITS
//         (sane reads: < 2^45; band-clipped alignments, where the forward and backward ridges separate: > 2^80).
//     A read that fails a check gets ST_LIN_FAULT and the host re-runs it through the log2-domain kernels (same GPU,
//     same scratch), whose range is unlimited.
#pragma once

#include "dp_kernels.cuh"

#ifndef DYN_RCP_UNROLL
#define DYN_RCP_UNROLL 1
#endif
#ifndef DYN_FWD_FAST
#define DYN_FWD_FAST 1
#endif

namespace dyn
{
namespace lin
{

// CTA-wide barriers inside pass 2 as well (recomputation / forward rows of every block in step): measured slower
// (1251 vs 1363 G MUFU/s on c2 x 8192: the barriers cost more than the instruction-cache locality gains), off
#ifndef DYN_BLOCK_SYNC
#define DYN_BLOCK_SYNC 0
#endif
template <int WPC>
DYN_DEV void cta_sync();

constexpr int DCPL = 100;              // largest offset deficit of a lane against its source-side neighbour
constexpr int E0V = 20;                // exponent of the lane maximum after a posterior-Viterbi renormalisation
constexpr float LIN_MASS_TOL = 2e-4f;  // |recorded posterior mass of a row - 1| above this is a range fault (observed on sane reads: < 5e-5;
                                       // the path posteriors are divided by the recorded mass, so what this bounds is the unrecorded rest)
constexpr double LIN_Z_TOL = 3e-3;     // |log2 Zf - log2 Zb| above this is a range fault
constexpr int LIN_GUARD_BITS = 70;     // see the header comment
constexpr int KAPPA_MAX_EXP = 118;     // the posterior factor 2^(OF + OB - Z2) is clamped here (sb * kappa must stay finite)

// component c (compile-time after unrolling) of a float4 held in registers
DYN_DEV float f4_get(const float4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }

// 2^e as a float; 0 for e < -126, 2^127 for e > 127
DYN_DEV float pow2i(int e)
{
	e = max(e, -127);
	e = min(e, 127);
	return __int_as_float((e + 127) << 23);
}

// two factors whose product is 2^d (exact for |d| <= 252)
DYN_DEV void pow2_split(int d, float& f1, float& f2)
{
	d = max(d, -300);
	d = min(d, 300);
	const int h = d / 2;
	f1 = pow2i(h);
	f2 = pow2i(d - h);
}

DYN_DEV float max3f(float a, float b, float c)
{
#ifndef DYN_HOST_EMU
	float r;
	asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
	return r;
#else
	return fmaxf(fmaxf(a, b), c);
#endif
}

// emission probability of the lane's CPL cells for sample x (inactive slots: c = CNEG -> 0)
template <class CFG>
DYN_DEV void emis_lin(const Warp<CFG>& w, float x, float (&p)[CFG::CPL])
{
#pragma unroll
	for (int j = 0; j < CFG::CPL; ++j) p[j] = ex2(w.emis(j, x));
}

// New integer offset of this lane after a renormalisation.
//   lm    largest value of the lane (0: the lane holds no non-zero cell)
//   O     current offset
//   DIR   +1: the source-side neighbour is lane+1 (backward), -1: lane-1 (posterior-Viterbi)
// Returns the new offset; all 32 lanes must call.
template <int DIR>
DYN_DEV int renorm_offset(float lm, int O, int e0, int lane)
{
	const bool alive = lm > 0.0f && lm < 3.0e38f;
	const unsigned m = __ballot_sync(FULL, alive);
	const int ex = ((__float_as_int(lm) >> 23) & 0xff) - 127 - e0;
	const int cand = O + ex;
	const int src = (lane + DIR) & 31;
	const int cand_src = __shfl_sync(FULL, cand, src);
	// nearest live lane in ring distance, source side first
	const unsigned r = __funnelshift_r(m, m, lane);  // bit k <=> lane (lane + k) & 31
	const int up = __ffs(r) - 1;                     // distance towards higher lanes (own bit is clear for a dead lane)
	const int dn = __clz(r) + 1;                     // distance towards lower lanes
	int pick;
	if (DIR > 0) pick = (up <= dn) ? lane + up : lane - dn;
	else pick = (dn <= up) ? lane - dn : lane + up;
	const int cand_near = __shfl_sync(FULL, cand, pick & 31);
	if (m == 0u) return O;  // nothing alive anywhere: the read is lost (detected by the caller through Z)
	if (alive) return ((m >> src) & 1u) ? max(cand, cand_src - DCPL) : cand;
	return cand_near - DCPL;
}

// ------------------------------------------------------------------------------------------------------
// backward recurrence (NT_aligner_api.cpp:158-207), linear domain
// ------------------------------------------------------------------------------------------------------
template <int CPL>
struct BwdL
{
	float bM[CPL], bE[CPL];
	int OB;          // true value = stored * 2^OB
	float sR1, sR2;  // sR1 * sR2 = 2^(OB(right lane) - OB(this lane))
};

template <class CFG>
DYN_DEV void bwd_row(Warp<CFG>& w, BwdL<CFG::CPL>& b, float x, float m1, float e2)
{
	constexpr int CPL = CFG::CPL;
	float p[CPL], A[CPL];
	// A[n] = bM[t+1][n] * p(t,n) * m1 is consumed by column n-1.  Slot 0 goes first: A[0] is what the left lane needs,
	// and the shuffle that carries it then has the rest of the row to complete.
	p[0] = ex2(w.emis(0, x));
	A[0] = b.bM[0] * (p[0] * m1);
	const float Araw = __shfl_sync(FULL, A[0], (w.lane + 1) & 31);
#pragma unroll
	for (int j = 1; j < CPL; ++j) p[j] = ex2(w.emis(j, x));
#pragma unroll
	for (int j = 1; j < CPL; ++j) A[j] = b.bM[j] * (p[j] * m1);
#pragma unroll
	for (int j = 0; j + 1 < CPL; ++j)
	{
		const float nm = b.bE[j] * p[j];   // bM[t][n] = bE[t+1][n] * p            (NT:200)
		b.bE[j] = fmaf(nm, e2, A[j + 1]);  //                                      (NT:194,201,204)
		b.bM[j] = nm;
	}
	const float nml = b.bE[CPL - 1] * p[CPL - 1];
	b.bM[CPL - 1] = nml;
	b.bE[CPL - 1] = fmaf(nml, e2, (Araw * b.sR1) * b.sR2);
}

// lane-local renormalisation by an exact power of two; returns the exponent increment (new OB - old OB)
template <class CFG>
DYN_DEV int bwd_renorm(Warp<CFG>& w, BwdL<CFG::CPL>& b)
{
	constexpr int CPL = CFG::CPL;
	float lm = b.bE[0];
#pragma unroll
	for (int j = 1; j + 1 < CPL; j += 2) lm = max3f(lm, b.bE[j], b.bE[j + 1]);
	if ((CPL & 1) == 0) lm = fmaxf(lm, b.bE[CPL - 1]);
	const int nO = renorm_offset<+1>(lm, b.OB, 0, w.lane);
	const int inc = nO - b.OB;
	const float sc = pow2i(-inc);
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		b.bM[j] *= sc;
		b.bE[j] *= sc;
	}
	b.OB = nO;
	const int obr = __shfl_sync(FULL, nO, (w.lane + 1) & 31);
	pow2_split(obr - nO, b.sR1, b.sR2);
	return inc;
}

template <class CFG>
DYN_DEV void bwd_step(Warp<CFG>& w, BwdL<CFG::CPL>& b, float x, bool slide, int& mid, float m1, float e2)
{
	constexpr int CPL = CFG::CPL;
	if (slide)
	{
		// see dp_kernels.cuh bwd_step: the column entering the band carried the ungated M-transition term
		const int nb = mid - 1 - w.bw;
		if (nb >= 0)
		{
			const PosConst v = w.pc[nb];
			with_slot<CPL>(w.lane, pmod(nb, CFG::SLOTS), SetEmisState<CPL>{w.em, v.a, w.bsel(v), v.c, b.bM, b.bE, 0.0f});
		}
	}
	bwd_row<CFG>(w, b, x, m1, e2);
	if (slide)
	{
		const int ntop = mid + w.bw;
		if (ntop < (int)w.N)
			with_slot<CPL>(w.lane, pmod(ntop, CFG::SLOTS), SetEmisState<CPL>{w.em, 0.0f, w.b_off(), CNEG, b.bM, b.bE, 0.0f});
		--mid;
	}
}

template <class CFG>
DYN_DEV void bwd_init_terminal(Warp<CFG>& w, BwdL<CFG::CPL>& b)
{
	constexpr int CPL = CFG::CPL;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		b.bM[j] = 0.0f;
		b.bE[j] = 0.0f;
	}
	const int q = pmod((int)w.N - 1, CFG::SLOTS);
	with_slot<CPL>(w.lane, q, SetOne<CPL>{b.bE, 1.0f});  // bE[T-1][N-1] = 1 (NT:170)
	b.OB = (w.lane == q / CPL) ? 0 : -DCPL;
	const int obr = __shfl_sync(FULL, b.OB, (w.lane + 1) & 31);
	pow2_split(obr - b.OB, b.sR1, b.sR2);
}

// checkpoints share the log-domain layout: [2*CPL][32] floats + one 8-byte word per lane (here: the int offset)
template <class CFG>
DYN_DEV void ckpt_store(const SlotScratch& sc, uint32_t idx, int lane, const BwdL<CFG::CPL>& b)
{
	constexpr int CPL = CFG::CPL;
	float* f = sc.ckpt + (size_t)idx * CFG::CKF;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		f[j * 32 + lane] = b.bM[j];
		f[(CPL + j) * 32 + lane] = b.bE[j];
	}
	reinterpret_cast<int*>(sc.ckpt_ob)[(size_t)idx * 64 + lane] = b.OB;
}

template <class CFG>
DYN_DEV void ckpt_load(const SlotScratch& sc, uint32_t idx, int lane, BwdL<CFG::CPL>& b)
{
	constexpr int CPL = CFG::CPL;
	const float* f = sc.ckpt + (size_t)idx * CFG::CKF;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		b.bM[j] = f[j * 32 + lane];
		b.bE[j] = f[(CPL + j) * 32 + lane];
	}
	b.OB = reinterpret_cast<const int*>(sc.ckpt_ob)[(size_t)idx * 64 + lane];
	const int obr = __shfl_sync(FULL, b.OB, (lane + 1) & 31);
	pow2_split(obr - b.OB, b.sR1, b.sR2);
}

// pull checkpoint idx towards L2 ahead of its use: a lane's values are CPL*2 rows of 128 bytes, one sector each
template <class CFG>
DYN_DEV void ckpt_prefetch(const SlotScratch& sc, uint32_t idx, int lane)
{
#ifndef DYN_HOST_EMU
	const float* f = sc.ckpt + (size_t)idx * CFG::CKF;
	// 2*CPL rows x 128 B = 104 sectors of 32 B for CPL = 13; lane l touches sectors l, l+32, l+64, l+96
#pragma unroll
	for (int q = 0; q < (2 * CFG::CPL * 4 + 31) / 32; ++q)
	{
		const int sct = q * 32 + lane;
		if (sct < 2 * CFG::CPL * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(f + sct * 8));
	}
	if (lane < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(sc.ckpt_ob) + (size_t)idx * 256 + lane * 32));
#else
	(void)sc; (void)idx; (void)lane;
#endif
}

// pass 1: backward over the whole read.  Returns log2 Zb (double) in every lane (NaN/-inf on a range fault).
// UNR: unroll factor of the slide-free 4-row groups.  1 (a loop over the bare row) for single-warp CTAs, whose warps
// are in different passes and must share the instruction cache with the loops of pass 2; 4 for the phase-synchronised
// CTAs, where every warp of the SM runs this pass at the same time (+2.7 %: the emissions of the next row overlap the tail
// of the previous one)
template <class CFG, bool STORE, int UNR = 1>
DYN_DEV double backward_pass(Warp<CFG>& w, const SlotScratch& sc, float m1, float e2)
{
	constexpr int CPL = CFG::CPL;
	BwdL<CPL> b;
	int mid = (int)band_mid(w.T - 1, w.ratio);
	w.load_window(mid);
	bwd_init_terminal<CFG>(w, b);
	if (STORE && ((w.T - 1) & (CFG::CK - 1)) == 0) ckpt_store<CFG>(sc, (w.T - 1) / CFG::CK, w.lane, b);

	int t = (int)w.T - 2;
	Chunk nxt = chunk_load<CFG>(w, (uint32_t)t & ~31u);
	while (t >= 0)
	{
		const uint32_t base = (uint32_t)t & ~31u;
		const Chunk cur = nxt;
		if (base >= 32) nxt = chunk_load<CFG>(w, base - 32);
		int i = t - (int)base;
#pragma unroll 1
		while (i >= 0)
		{
			if (CFG::RN % 4 == 0 && (i & 3) == 3 && ((cur.smask >> (i - 3)) & 0xfu) == 0u)
			{
				// an aligned group of four rows without a band slide (see UNR above)
#pragma unroll UNR
				for (int q = 0; q < 4; ++q) bwd_row<CFG>(w, b, __shfl_sync(FULL, cur.xv, i - q), m1, e2);
				const uint32_t tt = base + i - 3;
				if ((tt & (CFG::RN - 1)) == 0)
				{
					bwd_renorm<CFG>(w, b);
					if (STORE && (tt & (CFG::CK - 1)) == 0) ckpt_store<CFG>(sc, tt / CFG::CK, w.lane, b);
				}
				i -= 4;
				continue;
			}
			const uint32_t tt = base + i;
			const float x = __shfl_sync(FULL, cur.xv, i);
			bwd_step<CFG>(w, b, x, (cur.smask >> i) & 1u, mid, m1, e2);
			if ((tt & (CFG::RN - 1)) == 0)
			{
				bwd_renorm<CFG>(w, b);
				if (STORE && (tt & (CFG::CK - 1)) == 0) ckpt_store<CFG>(sc, tt / CFG::CK, w.lane, b);
			}
			--i;
		}
		t = (int)base - 1;
	}
	// Zb = bE[0][0] (NT:286): column 0 is ring slot 0 = lane 0, j 0
	const double z = log2((double)b.bE[0]) + (double)b.OB;
	return shfl_f64(z, 0);
}

// ------------------------------------------------------------------------------------------------------
// pass 2 state
// ------------------------------------------------------------------------------------------------------
template <int CPL>
struct FwdL
{
	float fM[CPL], fE[CPL];  // (true value) * 2^-OF
	float VM[CPL], VE[CPL];  // posterior-Viterbi products, (true value) * 2^-OV
	int OF;
	int OV;
	float sL1, sL2;          // product = 2^(OF(left lane) - OF(this lane)): brings the left lane's fE to this lane's scale
	float sV1, sV2;          // product = 2^(OV(left lane) - OV(this lane))
	float kap;               // 2^(OF + OB - Z2) for the backward offsets OB of the current row
	bool fault;              // sticky: the range guard tripped in this lane
};

// posterior factor of a lane: 2^(OF + OB - Z2) = c0 * 2^(OF + OB - floor(Z2)), exponent clamped to KAPPA_MAX_EXP
DYN_DEV float kappa(int OF, int OB, int Z2i, float c0)
{
	return c0 * pow2i(min(OF + OB - Z2i, KAPPA_MAX_EXP));
}

template <class CFG>
struct SmemL
{
	float* bE;  // [(CK+1)][CPL][32]
	int* OB;    // [NRN][32]  backward lane offsets in force for the rows up to and including t_lo + i*RN
	DYN_DEV explicit SmemL(unsigned char* p)
	{
		OB = reinterpret_cast<int*>(p);
		bE = reinterpret_cast<float*>(p + (size_t)CFG::NRN * 32 * 8);
	}
};

// lane-local renormalisation of the forward values (own maximum -> [1, 2), coupled to the left lane) and the range
// guard.  brow: this lane's backward values of (about) the same row; OB: their offset.
template <class CFG>
DYN_DEV void fwd_renorm(Warp<CFG>& w, FwdL<CFG::CPL>& f, const float (&brow)[CFG::CPL], int OB, int Z2i, float c0)
{
	constexpr int CPL = CFG::CPL;
	float lm = 0.0f, bm = 0.0f;
#pragma unroll
	for (int j = 0; j < CPL; ++j) lm = max3f(lm, f.fM[j], f.fE[j]);
#pragma unroll
	for (int j = 0; j + 1 < CPL; j += 2) bm = max3f(bm, brow[j], brow[j + 1]);
	if (CPL & 1) bm = fmaxf(bm, brow[CPL - 1]);
	// guard (only where both directions hold something): log2(F_lane * B_lane / Z) <= LIN_GUARD_BITS
	if (lm > 0.0f && bm > 0.0f)
	{
		const int g = ((__float_as_int(lm) >> 23) & 0xff) + ((__float_as_int(bm) >> 23) & 0xff) - 254 + f.OF + OB - Z2i;
		if (g > LIN_GUARD_BITS) f.fault = true;
	}
	const int nO = renorm_offset<-1>(lm, f.OF, 0, w.lane);
	const float sc = pow2i(f.OF - nO);
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		f.fM[j] *= sc;
		f.fE[j] *= sc;
	}
	f.OF = nO;
	const int ofl = __shfl_sync(FULL, nO, (w.lane + 31) & 31);
	pow2_split(ofl - nO, f.sL1, f.sL2);
	f.kap = kappa(nO, OB, Z2i, c0);
}

template <class CFG>
DYN_DEV void vit_renorm(Warp<CFG>& w, FwdL<CFG::CPL>& f)
{
	constexpr int CPL = CFG::CPL;
	float lm = 0.0f;
#pragma unroll
	for (int j = 0; j < CPL; ++j) lm = max3f(lm, f.VM[j], f.VE[j]);
	const int nO = renorm_offset<-1>(lm, f.OV, E0V, w.lane);
	const int inc = nO - f.OV;
	const float sc = pow2i(-inc);
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		f.VM[j] *= sc;
		f.VE[j] *= sc;
	}
	f.OV = nO;
	const int ovl = __shfl_sync(FULL, nO, (w.lane + 31) & 31);
	pow2_split(ovl - nO, f.sV1, f.sV2);
}

// One row of pass 2 (see dp_kernels.cuh fwd_row for the contract).  kapE / kapM: posterior factors of the extend and
// the match state of row t (they differ on a renormalisation row: bM[t] = bE[t+1] * p is formed in the offsets of
// row t+1).
template <class CFG, bool DO_V, bool DO_STEP>
DYN_DEV void fwd_row(Warp<CFG>& w, FwdL<CFG::CPL>& f, const SlotScratch& sc, RecSink& rs, float thr, uint32_t t, float x,
	bool slide, int& mid_f, float (&bc)[CFG::CPL], float (&bn)[CFG::CPL], const float* pf, bool has_pf, float kapE, float kapM,
	float m1, float e2, bool maybe_vit = true)
{
	constexpr int CPL = CFG::CPL;
	const int lane = w.lane;
	// the values the right lane needs are those of the previous row: send them first, consume them last
	const float vlraw = DO_V ? __shfl_sync(FULL, f.VE[CPL - 1], (lane + 31) & 31) : 0.0f;
	const float flraw = DO_STEP ? __shfl_sync(FULL, f.fE[CPL - 1], (lane + 31) & 31) : 0.0f;
	if (DO_STEP && slide) w.activate(mid_f + 1 + w.bw);  // column entering band(t+1)

	float p[CPL], PM[CPL], PE[CPL];
	if (DO_STEP) emis_lin<CFG>(w, x, p);
	if (DO_V)
	{
#pragma unroll
		for (int j = 0; j < CPL; ++j)
		{
			PE[j] = f.fE[j] * (bc[j] * kapE);
			// bM[t][n] = bE[t+1][n] * p(t,n) (NT:200); the last row has no match state
			PM[j] = DO_STEP ? f.fM[j] * ((bn[j] * p[j]) * kapM) : 0.0f;
		}
		if (DO_STEP && slide && mid_f - w.bw >= 0)
			with_slot<CPL>(lane, pmod(mid_f - w.bw, CFG::SLOTS), SetOne<CPL>{PM, 0.0f});  // see dp_kernels.cuh fwd_row
		// posterior-Viterbi fill (NT:357-362) as a max-product, in place from the highest slot down;
		// decision bit = sign(VM - VE): set <=> the E state of this cell is entered from E (test of NT:448)
		const float vl = (vlraw * f.sV1) * f.sV2;
		unsigned acc = 0;
		float lmax = 0.0f;
#pragma unroll
		for (int j = CPL - 1; j >= 0; --j)
		{
			const float vmx = fmaxf(f.VM[j], f.VE[j]);
			acc = __funnelshift_l(__float_as_uint(f.VM[j] - f.VE[j]), acc, 1);
			const float left = (j > 0) ? f.VE[j - 1] : vl;
			f.VM[j] = left * PM[j];
			f.VE[j] = vmx * PE[j];
			lmax = max3f(lmax, PM[j], PE[j]);
		}
		if (maybe_vit && (t & (CFG::RV - 1)) == 0) vit_renorm<CFG>(w, f);
		sc.bits[(size_t)t * 32 + lane] = (uint16_t)acc;

		// sparse posterior records (linear posteriors); a NaN/inf lane is recorded too so that the mass check sees it
		if (lane == 0) sc.rowptr[t] = rs.n;
		const bool hot = !(lmax <= thr);
		const unsigned hm = __ballot_sync(FULL, hot);
		if (hm)
		{
			const uint32_t pos = rs.n + __popc(hm & ((1u << lane) - 1u));
			if (hot && pos < rs.cap)
			{
				typedef LaneRec<CPL> Rec;
				float4* dst = reinterpret_cast<float4*>(static_cast<Rec*>(rs.recs) + pos);
				float tmp[Rec::NF];
#pragma unroll
				for (int j = 0; j < CPL; ++j)
				{
					tmp[j] = PM[j];
					tmp[CPL + j] = PE[j];
				}
				tmp[2 * CPL] = __int_as_float(lane);
#pragma unroll
				for (int q = 2 * CPL + 1; q < Rec::NF; ++q) tmp[q] = 0.0f;
#pragma unroll
				for (int q = 0; q < Rec::NF / 4; ++q) dst[q] = make_float4(tmp[4 * q], tmp[4 * q + 1], tmp[4 * q + 2], tmp[4 * q + 3]);
			}
			rs.n += __popc(hm);
			if (rs.n > rs.cap)
			{
				rs.overflow = true;
				rs.n = (uint32_t)rs.cap;
			}
		}
	}
	if (DO_STEP)
	{
		const float fl = (flraw * f.sL1) * f.sL2;
		if (has_pf)
		{
#pragma unroll
			for (int j = 0; j < CPL; ++j)
			{
				bc[j] = pf[j * 32 + lane];
				bn[j] = pf[CFG::ROWF + j * 32 + lane];
			}
		}
#pragma unroll
		for (int j = CPL - 1; j >= 0; --j)
		{
			const float left = (j > 0) ? f.fE[j - 1] : fl;
			const float ne = fmaf(f.fE[j], e2, f.fM[j]) * p[j];  // (fM + fE*e2) * p    (NT:146-150, e1 = 1)
			f.fM[j] = left * (p[j] * m1);                        // fE[t][n-1] * p * m1  (NT:143)
			f.fE[j] = ne;
		}
		if (slide)
		{
			const int nold = mid_f - w.bw;
			if (nold >= 0)
				with_slot<CPL>(lane, pmod(nold, CFG::SLOTS), SetEmisState<CPL>{w.em, 0.0f, w.b_off(), CNEG, f.fM, f.fE, 0.0f});
			++mid_f;
		}
	}
}

// 16-byte global store under a predicate: keeps the sparse-record write of the fast row free of divergent regions
DYN_DEV void st_v4_if(bool on, float* dst, float a, float b, float c, float d)
{
#ifndef DYN_HOST_EMU
	asm volatile(
		"{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %0, 0;\n\t@p st.global.v4.f32 [%1], {%2, %3, %4, %5};\n\t}" ::"r"((int)on),
		"l"(__cvta_generic_to_global(dst)), "f"(a), "f"(b), "f"(c), "f"(d)
		: "memory");
#else
	if (on)
	{
		dst[0] = a;
		dst[1] = b;
		dst[2] = c;
		dst[3] = d;
	}
#endif
}

// fwd_row<CFG, true, true> for a row that (a) does not slide the band, (b) has both neighbouring backward rows in shared
// memory and (c) is not a renormalisation row of the posterior-Viterbi scores: ONE basic block.  The generic row spends
// about a quarter of its time at the boundaries of its nine conditional regions (nothing is scheduled across them);
// groups of RN rows without a slide (87 % at 30 samples per base) run through this body instead.
template <class CFG>
DYN_DEV void fwd_row_fast(Warp<CFG>& w, FwdL<CFG::CPL>& f, const SlotScratch& sc, RecSink& rs, float thr, uint32_t t, float x,
	float (&bc)[CFG::CPL], float (&bn)[CFG::CPL], const float* pf, float kapE, float kapM, float m1, float e2)
{
	constexpr int CPL = CFG::CPL;
	typedef LaneRec<CPL> Rec;
	const int lane = w.lane;
	const float vlraw = __shfl_sync(FULL, f.VE[CPL - 1], (lane + 31) & 31);
	const float flraw = __shfl_sync(FULL, f.fE[CPL - 1], (lane + 31) & 31);
	float p[CPL], PM[CPL], PE[CPL];
	emis_lin<CFG>(w, x, p);
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		PE[j] = f.fE[j] * (bc[j] * kapE);
		PM[j] = f.fM[j] * ((bn[j] * p[j]) * kapM);  // bM[t][n] = bE[t+1][n] * p(t,n) (NT:200)
	}
	const float vl = (vlraw * f.sV1) * f.sV2;
	unsigned acc = 0;
	float lmax = 0.0f;
#pragma unroll
	for (int j = CPL - 1; j >= 0; --j)
	{
		const float vmx = fmaxf(f.VM[j], f.VE[j]);
		acc = __funnelshift_l(__float_as_uint(f.VM[j] - f.VE[j]), acc, 1);
		const float left = (j > 0) ? f.VE[j - 1] : vl;
		f.VM[j] = left * PM[j];
		f.VE[j] = vmx * PE[j];
		lmax = max3f(lmax, PM[j], PE[j]);
	}
	sc.bits[(size_t)t * 32 + lane] = (uint16_t)acc;
	if (lane == 0) sc.rowptr[t] = rs.n;
	{
		const bool hot = !(lmax <= thr);
		const unsigned hm = __ballot_sync(FULL, hot);
		const uint32_t pos = rs.n + __popc(hm & ((1u << lane) - 1u));
		const bool wr = hot && pos < rs.cap;
		float* dst = static_cast<Rec*>(rs.recs)[wr ? pos : 0u].v;
		float tmp[Rec::NF];
#pragma unroll
		for (int j = 0; j < CPL; ++j)
		{
			tmp[j] = PM[j];
			tmp[CPL + j] = PE[j];
		}
		tmp[2 * CPL] = __int_as_float(lane);
#pragma unroll
		for (int q = 2 * CPL + 1; q < Rec::NF; ++q) tmp[q] = 0.0f;
#pragma unroll
		for (int q = 0; q < Rec::NF / 4; ++q) st_v4_if(wr, dst + 4 * q, tmp[4 * q], tmp[4 * q + 1], tmp[4 * q + 2], tmp[4 * q + 3]);
		const uint32_t nn = rs.n + __popc(hm);
		rs.overflow = rs.overflow || (nn > rs.cap);
		rs.n = (nn > rs.cap) ? (uint32_t)rs.cap : nn;
	}
	const float fl = (flraw * f.sL1) * f.sL2;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		bc[j] = bn[j];
		bn[j] = pf[CFG::ROWF + j * 32 + lane];
	}
#pragma unroll
	for (int j = CPL - 1; j >= 0; --j)
	{
		const float left = (j > 0) ? f.fE[j - 1] : fl;
		const float ne = fmaf(f.fE[j], e2, f.fM[j]) * p[j];  // (fM + fE*e2) * p    (NT:146-150, e1 = 1)
		f.fM[j] = left * (p[j] * m1);                        // fE[t][n-1] * p * m1  (NT:143)
		f.fE[j] = ne;
	}
}

// pass 2: forward + posterior + posterior-Viterbi fill.  Returns log2 Zf - log2 Zb (NaN on a range fault).
// WPC > 1: the warps of the CTA also keep the two halves of every block in step (recomputation, forward rows), so at
// any time the whole SM runs one of the two loops; kb_sync = the largest block count among the CTA's reads (a warp whose
// read is shorter keeps passing the barriers)
template <class CFG, int WPC = 1>
DYN_DEV double forward_posterior_pass(Warp<CFG>& w, const SlotScratch& sc, const BatchArgs& args, unsigned char* smem_raw,
	double Z2, float thr, float m1, float e2, uint32_t& nrec_out, bool& overflow, uint32_t kb_sync = 0)
{
	constexpr int CPL = CFG::CPL;
	constexpr int CK = CFG::CK;
	constexpr int RN = CFG::RN;
	constexpr int ROWF = CFG::ROWF;
	constexpr int RCP_UNR = (WPC > 1) ? 3 : 1;
	SmemL<CFG> sm(smem_raw);
	FwdL<CPL> f;
	BwdL<CPL> b;
	const int lane = w.lane;
	const uint32_t T = w.T;
	RecSink rs;
	rs.recs = sc.recs;
	rs.cap = args.rec_cap;
	rs.n = 0;
	rs.overflow = false;
	// posterior = sf * sb * 2^(OF + OB - Z2) = sf * sb * c0 * 2^(OF + OB - Z2i)
	const double Z2f = floor(Z2);
	const int Z2i = (int)Z2f;
	const float c0 = (float)exp2(Z2f - Z2);

	int mid_f = 0;
	float bc[CPL], bn[CPL];
#pragma unroll
	for (int j = 0; j < CPL; ++j) bc[j] = bn[j] = 0.0f;
	w.load_window(0);
	const uint32_t kb = (T - 1) / CK;
	Chunk cur = chunk_load<CFG>(w, 0);
	Chunk nxtc = chunk_load<CFG>(w, 32);

	for (uint32_t k = 0; k <= kb; ++k)
	{
		const uint32_t t_lo = k * CK;
		const uint32_t t_hi = t_lo + CK;
		if (k > 0 && (t_lo & 31u) == 0)
		{
			cur = nxtc;
			nxtc = chunk_load<CFG>(w, t_lo + 32);
		}
		// ---- step a: recompute the backward rows of this block into shared memory -----------------------
		// sm.OB[i] = backward offsets in force for rows t_lo + (i-1)*RN + 1 .. t_lo + i*RN
		const bool from_ckpt = (t_hi <= T - 1);
		const uint32_t src_row = from_ckpt ? t_hi : T - 1;
		int mid_b = (int)band_mid(src_row, w.ratio);
		w.slide_window_up(mid_f, mid_b);
		if (from_ckpt)
		{
			ckpt_load<CFG>(sc, k + 1, lane, b);
			if (t_hi + CK <= T - 1) ckpt_prefetch<CFG>(sc, k + 2, lane);  // pull the next block's checkpoint into L2
		}
		else bwd_init_terminal<CFG>(w, b);
		{
			float* dst = sm.bE + (size_t)(src_row - t_lo) * ROWF;
#pragma unroll
			for (int j = 0; j < CPL; ++j) dst[j * 32 + lane] = b.bE[j];
			sm.OB[((src_row - t_lo + RN - 1) / RN) * 32 + lane] = b.OB;
		}
		{
			int tt = (int)src_row - 1;
#pragma unroll 1
			while (tt >= (int)t_lo)
			{
				const int i = tt & 31;
				if (DYN_RCP_UNROLL && RN % 4 == 0 && (i & 3) == 3 && ((cur.smask >> (i - 3)) & 0xfu) == 0u)
				{
					// an aligned group of four rows without a band slide (see backward_pass)
					float* dst = sm.bE + (size_t)(tt - (int)t_lo) * ROWF + lane;
#pragma unroll RCP_UNR
					for (int q = 0; q < 3; ++q)  // (unrolled for the phase-synchronised CTAs, +1.2 %; see backward_pass)
					{
						bwd_row<CFG>(w, b, __shfl_sync(FULL, cur.xv, i - q), m1, e2);
#pragma unroll
						for (int j = 0; j < CPL; ++j) dst[j * 32] = b.bE[j];
						dst -= ROWF;
					}
					bwd_row<CFG>(w, b, __shfl_sync(FULL, cur.xv, i - 3), m1, e2);
					if (((tt - 3) & (RN - 1)) == 0)
					{
						bwd_renorm<CFG>(w, b);
						sm.OB[((tt - 3 - (int)t_lo) / RN) * 32 + lane] = b.OB;
					}
#pragma unroll
					for (int j = 0; j < CPL; ++j) dst[j * 32] = b.bE[j];
					tt -= 4;
					continue;
				}
				const float x = __shfl_sync(FULL, cur.xv, i);
				const bool sl = (cur.smask >> i) & 1u;
				if (sl)
				{
					// the column that enters the band at row tt: (tt+1, n) is out of band, but its slot kept computing the
					// ungated M-transition term (see bwd_step); zero it in the stored row tt+1 so that the forward pass
					// gets bM[tt][n] = bE[tt+1][n] * p = 0 without a special case
					const int nb = mid_b - 1 - w.bw;
					if (nb >= 0)
					{
						const int q = pmod(nb, CFG::SLOTS);
						if (lane == q / CPL) sm.bE[(size_t)(tt + 1 - (int)t_lo) * ROWF + (q % CPL) * 32 + lane] = 0.0f;
					}
				}
				bwd_step<CFG>(w, b, x, sl, mid_b, m1, e2);
				float* dst = sm.bE + (size_t)(tt - (int)t_lo) * ROWF;
				if ((tt & (RN - 1)) == 0)
				{
					bwd_renorm<CFG>(w, b);
					sm.OB[((tt - (int)t_lo) / RN) * 32 + lane] = b.OB;
				}
#pragma unroll
				for (int j = 0; j < CPL; ++j) dst[j * 32 + lane] = b.bE[j];
				--tt;
			}
		}
		__syncwarp();
		if (DYN_BLOCK_SYNC) cta_sync<WPC>();

		// ---- step b: forward rows t_lo .. min(t_hi, T) - 1 ----------------------------------------------
		uint32_t t = t_lo;
		if (k == 0)
		{
			// row 0: fE[0][0] = 1 (NT:120), VE[0][0] = 1 (NT:336)
#pragma unroll
			for (int j = 0; j < CPL; ++j)
			{
				f.fM[j] = 0.0f;
				f.fE[j] = 0.0f;
				f.VM[j] = 0.0f;
				f.VE[j] = 0.0f;
			}
			f.OF = (lane == 0) ? 0 : -DCPL;
			f.OV = (lane == 0) ? 0 : -DCPL;
			f.fault = false;
			if (lane == 0)
			{
				f.fE[0] = 1.0f;
				f.VE[0] = 1.0f;
			}
			{
				const int ofl = __shfl_sync(FULL, f.OF, (lane + 31) & 31);
				pow2_split(ofl - f.OF, f.sL1, f.sL2);
				const int ovl = __shfl_sync(FULL, f.OV, (lane + 31) & 31);
				pow2_split(ovl - f.OV, f.sV1, f.sV2);
			}
			fwd_row<CFG, false, true>(w, f, sc, rs, thr, 0, __shfl_sync(FULL, cur.xv, 0), cur.smask & 1u, mid_f, bc, bn,
				sm.bE, false, 0.0f, 0.0f, m1, e2);
			f.kap = kappa(f.OF, sm.OB[32 + lane], Z2i, c0);  // rows 1 .. RN
			t = 1;
		}
		const uint32_t t_end = min(t_hi, T - 1);
		if (t < t_end)
		{
			const float* row = sm.bE + (size_t)(t - t_lo) * ROWF;
#pragma unroll
			for (int j = 0; j < CPL; ++j)
			{
				bc[j] = row[j * 32 + lane];
				bn[j] = row[ROWF + j * 32 + lane];
			}
		}
#pragma unroll 1
		while (t < t_end)
		{
			const int i = t & 31;
			const uint32_t r = t - t_lo;
			if (DYN_FWD_FAST && args.fwd_fast && RN % 4 == 0 && (t & 3u) == 0 && t + 4 <= t_end)
			{
				// an aligned group of RN rows inside the block: the branch-free row body (see fwd_row_fast)
				const unsigned sl4 = (cur.smask >> i) & 0xfu;
				// on a renormalisation row bM[t] = bE[t+1] * p lives in the offsets of the next rows
				const float kapN = ((t & (RN - 1)) == 0) ? kappa(f.OF, sm.OB[(r / RN + 1) * 32 + lane], Z2i, c0) : f.kap;
				if ((t & (CFG::RV - 1)) == 0) vit_renorm<CFG>(w, f);
				float kapE = f.kap;
				const float* pf = sm.bE + (size_t)(r + 1) * ROWF;
				if (sl4 == 0u)
				{
#pragma unroll 1
					for (int q = 0; q < 4; ++q)
					{
						const float x = __shfl_sync(FULL, cur.xv, i + q);
						fwd_row_fast<CFG>(w, f, sc, rs, thr, t + q, x, bc, bn, pf, kapE, kapN, m1, e2);
						kapE = kapN;
						pf += ROWF;
					}
				}
				else
				{
					// a band slide between rows t+q and t+q+1: the column entering band(t+q+1) is activated before the row,
					// the column leaving is retired after it; the match posterior of the leaving column is forced through
					// the zero the recomputation wrote into the backward row t+q+1 (see step a)
#pragma unroll 1
					for (int q = 0; q < 4; ++q)
					{
						const bool sl = (sl4 >> q) & 1u;
						const float x = __shfl_sync(FULL, cur.xv, i + q);
						if (sl) w.activate(mid_f + 1 + w.bw);
						fwd_row_fast<CFG>(w, f, sc, rs, thr, t + q, x, bc, bn, pf, kapE, kapN, m1, e2);
						if (sl)
						{
							const int nold = mid_f - w.bw;
							if (nold >= 0)
								with_slot<CPL>(lane, pmod(nold, CFG::SLOTS), SetEmisState<CPL>{w.em, 0.0f, w.b_off(), CNEG, f.fM, f.fE, 0.0f});
							++mid_f;
						}
						kapE = kapN;
						pf += ROWF;
					}
				}
				f.kap = kapN;
				if (((t + 4) & (RN - 1)) == 0) fwd_renorm<CFG>(w, f, bc, sm.OB[((r + 4) / RN) * 32 + lane], Z2i, c0);
				t += 4;
				continue;
			}
			const bool rn_row = (t & (RN - 1)) == 0;
			// on a renormalisation row bM[t] = bE[t+1] * p lives in the offsets of the next rows
			const float kapN = rn_row ? kappa(f.OF, sm.OB[(r / RN + 1) * 32 + lane], Z2i, c0) : f.kap;
			const float x = __shfl_sync(FULL, cur.xv, i);
			fwd_row<CFG, true, true>(w, f, sc, rs, thr, t, x, (cur.smask >> i) & 1u, mid_f, bc, bn,
				sm.bE + (size_t)(r + 1) * ROWF, t + 1 < t_end, f.kap, kapN, m1, e2);
			f.kap = kapN;
			// the forward values are now those of row t+1 (bc: the backward row t+1, or still row t at a block end)
			if (((t + 1) & (RN - 1)) == 0) fwd_renorm<CFG>(w, f, bc, sm.OB[((r + 1) / RN) * 32 + lane], Z2i, c0);
			++t;
		}
		__syncwarp();
		if (DYN_BLOCK_SYNC) cta_sync<WPC>();
	}
	if (DYN_BLOCK_SYNC && WPC > 1)
		for (uint32_t k = kb + 1; k <= kb_sync; ++k)
		{
			cta_sync<WPC>();
			cta_sync<WPC>();
		}
	{
		const float* row = sm.bE + (size_t)((T - 1) - kb * CK) * ROWF;
#pragma unroll
		for (int j = 0; j < CPL; ++j) bc[j] = row[j * 32 + lane];
		fwd_row<CFG, true, false>(w, f, sc, rs, thr, T - 1, 0.0f, false, mid_f, bc, bn, sm.bE, false, f.kap, 0.0f, m1, e2);
	}
	// Zf = fE[T-1][N-1] (NT:285)
	float v = 0.0f;
	with_slot<CPL>(lane, pmod((int)w.N - 1, CFG::SLOTS), GetOne<CPL>{f.fE, v});
	const int ql = pmod((int)w.N - 1, CFG::SLOTS) / CPL;
	double dz = log2((double)v) + (double)f.OF - Z2;
	dz = shfl_f64(dz, ql);
	if (__any_sync(FULL, f.fault)) dz = NAN;
	if (lane == 0) sc.rowptr[T] = rs.n;
	nrec_out = rs.n;
	overflow = rs.overflow;
	return dz;
}

// ------------------------------------------------------------------------------------------------------
// pass 3: traceback (NT:383-456) — the walk over the decision bits is shared with the log2-domain kernels; the
// path posteriors come from LINEAR records and every row's recorded mass is checked
// ------------------------------------------------------------------------------------------------------
// returns 0 ok, 1 incomplete path, 2 range fault (row mass)
template <class CFG>
DYN_DEV int traceback_pass(Warp<CFG>& w, const SlotScratch& sc, const BatchArgs& args, unsigned char* smem_raw,
	const ReadDesc& rd)
{
	constexpr int CPL = CFG::CPL;
	constexpr int SLOTS = CFG::SLOTS;
	typedef LaneRec<CPL> Rec;
	const int lane = w.lane;
	const uint32_t T = w.T;
	uint32_t t_first = 0;
	if (!trace_decisions<CFG>(w, sc, args, smem_raw, rd, t_first)) return 1;

	const Rec* recs = static_cast<const Rec*>(sc.recs);
	bool bad = false;
	// NR rows per lane and iteration: the row's path cell, its record range and its first record (vector loads) are
	// fetched for all NR rows before anything is consumed, so a lane has NR independent chains of dependent loads in
	// flight instead of one (the records of a long read come from DRAM)
	constexpr int NR = 4;
	constexpr int NV = Rec::NF / 4;
	for (uint32_t rb = t_first; rb < T; rb += 32 * NR)
	{
		uint32_t pv[NR], r0[NR], r1[NR];
#pragma unroll
		for (int q = 0; q < NR; ++q)
		{
			const uint32_t r = rb + q * 32 + lane;
			const bool in = r < T;
			pv[q] = in ? sc.pn[r] : 0u;
			r0[q] = in ? sc.rowptr[r] : 0u;
			r1[q] = in ? sc.rowptr[r + 1] : 0u;
		}
		float4 rv[NR][NV];
#pragma unroll
		for (int q = 0; q < NR; ++q)
		{
			const float4* src = reinterpret_cast<const float4*>(recs + r0[q]);
			const bool has = r1[q] > r0[q];
#pragma unroll
			for (int k = 0; k < NV; ++k) rv[q][k] = has ? src[k] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		}
#pragma unroll
		for (int q = 0; q < NR; ++q)
		{
			const uint32_t r = rb + q * 32 + lane;
			if (r >= T) continue;
			const uint32_t col = pv[q] & 0x7fffffffu;
			const bool isM = (pv[q] >> 31) != 0;
			const int sl = (int)(col % SLOTS);
			const int ql = sl / CPL, j = sl - ql * CPL;
			const int want = (isM ? 0 : CPL) + j;
			float mass = 0.0f, lp = 0.0f;
			if (r1[q] > r0[q])
			{
				// first record, from registers: same summation order as the loop below
				float pick = 0.0f;
#pragma unroll
				for (int c = 0; c < 2 * CPL; ++c)
				{
					const float fc = f4_get(rv[q][c >> 2], c & 3);
					mass += fc;
					pick = (c == want) ? fc : pick;
				}
				if (__float_as_int(f4_get(rv[q][(2 * CPL) >> 2], (2 * CPL) & 3)) == ql) lp = pick;
			}
			for (uint32_t i = r0[q] + 1; i < r1[q]; ++i)
			{
				const float* f = recs[i].v;
				for (int c = 0; c < 2 * CPL; ++c) mass += f[c];
				if (__float_as_int(f[2 * CPL]) == ql) lp = f[want];
			}
			if (!(fabsf(mass - 1.0f) <= LIN_MASS_TOL)) bad = true;
			sc.pp[r] = lp / mass;
		}
	}
	// rows before the first path row (t_first > 1 never happens for a complete path, which starts at row 1)
	if (__any_sync(FULL, bad)) return 2;
	__threadfence_block();
	__syncwarp();
	segment_medians<CFG>(w, sc, args, rd);
	return 0;
}

// training statistics (NT:494-514, 641-725) from the linear records
template <class CFG>
DYN_DEV bool train_stats_pass(Warp<CFG>& w, const SlotScratch& sc, const BatchArgs& args, const ReadDesc& rd,
	double& xi_m, double& xi_e)
{
	constexpr int CPL = CFG::CPL;
	typedef LaneRec<CPL> Rec;
	const Rec* recs = static_cast<const Rec*>(sc.recs);
	const int lane = w.lane;
	double sm_ = 0.0, se_ = 0.0;
	// first make sure every row's recorded mass is sane: nothing may be accumulated for a read that is going to be
	// re-run in the log2 domain
	bool bad = false;
	for (uint32_t r = 1 + lane; r < w.T; r += 32)
	{
		const uint32_t r0 = sc.rowptr[r], r1 = sc.rowptr[r + 1];
		float mass = 0.0f;
		for (uint32_t i = r0; i < r1; ++i)
		{
			const float* f = recs[i].v;
			for (int c = 0; c < 2 * CPL; ++c) mass += f[c];
		}
		if (!(fabsf(mass - 1.0f) <= LIN_MASS_TOL)) bad = true;
	}
	if (__any_sync(FULL, bad)) return false;
	for (uint32_t r = 1 + lane; r < w.T; r += 32)
	{
		const uint32_t r0 = sc.rowptr[r], r1 = sc.rowptr[r + 1];
		double mass = 0.0;
		for (uint32_t i = r0; i < r1; ++i)
		{
			const float* f = recs[i].v;
			for (int c = 0; c < 2 * CPL; ++c) mass += (double)f[c];
		}
		const double inv = 1.0 / mass;
		const double xo = (double)w.sig[r - 1];
		const int n0 = (int)band_mid(r, w.ratio) - w.bw;
		for (uint32_t i = r0; i < r1; ++i)
		{
			const float* f = recs[i].v;
			const int rl = __float_as_int(f[2 * CPL]);
			for (int j = 0; j < CPL; ++j)
			{
				const double pm = (double)f[j] * inv, pe = (double)f[CPL + j] * inv;
				const double g = pm + pe;
				if (!(g > 1e-12)) continue;
				const int col = w.col_of_slot(rl * CPL + j, n0);
				atomicAdd(&args.read_w[rd.pc_off + col], g);
				atomicAdd(&args.read_x[rd.pc_off + col], g * xo);
				atomicAdd(&args.read_xx[rd.pc_off + col], g * xo * xo);
				sm_ += pm;
				se_ += pe;
			}
		}
	}
	for (int o = 16; o; o >>= 1)
	{
		sm_ += shfl_f64(sm_, (lane + o) & 31);
		se_ += shfl_f64(se_, (lane + o) & 31);
	}
	xi_m = sm_;
	xi_e = se_ - sm_;
	return true;
}

// CTA-wide barrier between the passes of the phase-synchronised launch shape (WPC warps per CTA, each with its own
// read of nearly the same length): all warps of an SM then run the SAME pass, so they share its code in the instruction
// cache instead of evicting each other's loops.  Every warp of the CTA passes the same number of barriers per read.
template <int WPC>
DYN_DEV void cta_sync()
{
#ifndef DYN_HOST_EMU
	if (WPC > 1) __syncthreads();
#endif
}

// one read, all passes.  A read the linear arithmetic cannot represent leaves with ST_LIN_FAULT.
// live: false for a warp that has no read in this round of its CTA (it only keeps the barriers company)
template <class CFG, int MODE, int WPC = 1>
DYN_DEV void align_read(const BatchArgs& args, const ReadDesc& rd, uint32_t ridx, const SlotScratch& sc,
	unsigned char* smem_raw, int lane, bool live = true, uint32_t kb_sync = 0)
{
	if (WPC > 1 && !live)
	{
		cta_sync<WPC>();
		if (DYN_BLOCK_SYNC && MODE != 0)
			for (uint32_t k = 0; k <= kb_sync; ++k)
			{
				cta_sync<WPC>();
				cta_sync<WPC>();
			}
		cta_sync<WPC>();
		return;
	}
	Warp<CFG> w;
	w.lane = lane;
	w.S = rd.S;
	w.T = rd.S + 1;
	w.N = rd.N;
	w.bw = (int)rd.bw;
	w.ratio = rd.ratio;
	w.sig = args.signal + rd.sig_off;
	w.pc = args.pc + rd.pc_off;
	w.m1 = args.m1;
	w.e2 = args.e2;
	w.ua = args.uni_a;
	w.uc = args.uni_c;
	const float m1 = args.m1_lin, e2 = args.e2_lin;

	ReadOut out;
	out.Z = 0.0;
	out.dZ = 0.0;
	out.nrec = 0;
	out.status = ST_OK;
	out.xi_m = 0.0;
	out.xi_e = 0.0;

	constexpr int UNR = (WPC > 1) ? 4 : 1;
	const double Z2 = (MODE == 0) ? backward_pass<CFG, false, UNR>(w, sc, m1, e2) : backward_pass<CFG, true, UNR>(w, sc, m1, e2);
	out.Z = Z2 * LN2;
	cta_sync<WPC>();
	bool go = false;  // pass 3 follows
	if (!(Z2 > -1.0e30 && Z2 < 1.0e30))
	{
		out.status = ST_LIN_FAULT;  // underflow of every path or NaN/inf: let the log2-domain kernels decide
		if (DYN_BLOCK_SYNC && WPC > 1 && MODE != 0)
			for (uint32_t k = 0; k <= kb_sync; ++k)
			{
				cta_sync<WPC>();
				cta_sync<WPC>();
			}
	}
	else if (MODE != 0)
	{
		uint32_t nrec = 0;
		bool overflow = false;
		const double dz2 = forward_posterior_pass<CFG, WPC>(w, sc, args, smem_raw, Z2, args.thr_lin, m1, e2, nrec, overflow, kb_sync);
		out.nrec = nrec;
		out.dZ = dz2 * LN2;
		if (!(fabs(dz2) <= LIN_Z_TOL)) out.status = ST_LIN_FAULT;
		else if (overflow) out.status = ST_REC_OVERFLOW;
		else go = true;
	}
	cta_sync<WPC>();
	if (go)
	{
		if (MODE == 1)
		{
			const int rc = lin::traceback_pass<CFG>(w, sc, args, smem_raw, rd);
			if (rc) out.status = ST_LIN_FAULT;
		}
		else
		{
			__threadfence_block();
			__syncwarp();
			if (!lin::train_stats_pass<CFG>(w, sc, args, rd, out.xi_m, out.xi_e)) out.status = ST_LIN_FAULT;
		}
	}
	if (lane == 0) args.out[ridx] = out;
}

} // namespace lin
} // namespace dyn
