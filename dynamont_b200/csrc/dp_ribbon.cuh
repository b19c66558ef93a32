// Ribbon kernels: the banded forward / backward / posterior-decoding passes of basic ("NT") mode evaluated on a NARROW
// ADAPTIVE WINDOW of lattice columns instead of the whole reference band.
//
// Why: the reference band (NT_aligner_api.cpp:90-108: 2*200+1 columns around the diagonal t*N/T) is that wide only
// because the true alignment drifts away from the diagonal (a random walk in the dwell times); at any one row the
// forward and backward values that are not negligible occupy 10 - 30 columns (measured on the compiled
// reference).  In the FP32 block-floating-point arithmetic of dp_linear.cuh everything further out is exactly 0 (it
// underflows), so the full-band kernels spend > 80 % of their instructions multiplying zeros.  Here one warp still owns
// one read and the passes are the same three (backward with checkpoints, recomputation + forward + posterior-Viterbi fill,
// traceback + medians), but a lane holds C = 2 (or 4) columns, i.e. the window is 32*C - 1 columns wide, and it FOLLOWS
// the probability mass:
//
//   * pass 1 (backward, t descending) decides once per 32-row chunk by how many columns the window moves down during the
//     chunk (0 .. 32; the alignment path moves at most one column per row, so the window can always keep up): it centres
//     the window on the lanes whose largest value is within 2^-G of the row maximum.  The decision is stored as the
//     window centre of the chunk's first row plus one slide bit per row (8 bytes per 32 rows).
//   * pass 2 (recomputation + forward) replays exactly that schedule, so the recomputed backward rows are bit-identical
//     to pass 1 and forward and backward values of a column always meet in the same ring slot.
//   * nothing is assumed, everything is checked on the device: (i) in both directions the two edge lanes of the window
//     must stay more than G bits below the row maximum at every renormalisation, (ii) the window must lie inside the
//     reference band at every row (so the cells it drops are exactly the ones that are negligible, never ones the
//     reference forces to -inf), (iii) Zf must equal Zb, (iv) the posterior mass of every renormalisation row must be 1
//     (one-sided losses show up there because the two directions are windowed independently of each other's values),
//     (v) the range guard of dp_linear.cuh.  A read that fails any check leaves with ST_LIN_FAULT and the host re-runs it
//     through the full-band kernels (same GPU).  Short reads (band narrower than the window) go there directly.
//
// Other differences from dp_linear.cuh:
//   * posteriors are kept normalised by a closed loop: every RN rows the row mass is measured (one warp reduction) and
//     the posterior factor is corrected by it, which removes the common-mode FP32 drift at the source;
//   * one 16-byte row header {first record, hot-lane mask, decision bits} + one 16-byte record per hot lane replace the
//     64-byte decision words and 112-byte lane records;
//   * training statistics (NT:494-514) are accumulated in registers per ring slot (a slot holds one column, i.e. one
//     kmer, for as long as the column is inside the window), centred on the model mean, and written once when the
//     column leaves the window: no per-cell atomics, no records, no traceback.
#pragma once

#include "dp_linear.cuh"

namespace dyn
{
namespace rib
{

template <int C_, int CK_, int RN_, int RV_>
struct RCfg
{
	static constexpr bool UNI = false;
	static constexpr int CPL = C_;
	static constexpr int SLOTS = 32 * C_;
	static constexpr int HW = (SLOTS - 2) / 2;  // live columns of a row: [mid - HW, mid + HW]; one ring slot stays dead
	static constexpr int CK = CK_;
	static constexpr int RN = RN_;
	static constexpr int RV = RV_;
	static constexpr int NRN = CK_ / RN_ + 1;
	static constexpr int CKF = 2 * C_ * 32;
	static constexpr int ROWF = C_ * 32;
	static constexpr int HDRW = (2 + C_ + 3) / 4 * 4;  // words of a row header: first record, hot-lane mask, C decision words
	static constexpr int RECF = (2 * C_ + 3) / 4 * 4;  // floats of a lane record: C match + C extend posteriors
	static constexpr size_t SMEM_BYTES = (size_t)(CK_ + 2) * ROWF * 4 + (size_t)NRN * 32 * 4;
	static_assert(CK_ % RN_ == 0 && 32 % CK_ == 0, "CK must divide 32 and be a multiple of RN");
	static_assert((RN_ & (RN_ - 1)) == 0 && (RV_ & (RV_ - 1)) == 0, "RN, RV: powers of two");
};

#if defined(DYN_HOST_EMU) && defined(DYN_RIB_DEBUG)
#define RIB_DBG(...) do { if (threadIdx.x == 0) fprintf(stderr, __VA_ARGS__); } while (0)
#else
#define RIB_DBG(...) do { } while (0)
#endif

constexpr int BAND_SLACK = 18;
// Largest offset deficit of a lane against the lanes its inflow can come from (dp_linear.cuh uses 100).  Smaller here
// because a lane that was empty at the last renormalisation can hold the ridge a few rows later (C = 2: a lane is crossed in
// two rows): its values then sit 2^RDC above [1, 2) in BOTH directions, and the posterior factor 2^(OF + OB - Z) must
// still be a normal float (>= 2^-126): 2 * (RDC + 12) < 126.  Cells more than 2^-(126 + RDC) below a neighbouring lane's
// maximum flush to zero — far below the 2^-G the window guard already treats as nothing.
constexpr int RDC = 30;
// closed-loop check of the posterior mass: |sum of the row masses since the last check - number of rows| above this
// is a fault (typical: < 1e-5; a row that lost 1e-4 of its mass to an FP32 range problem trips it)
constexpr float RIB_MASS_TOL = 1e-4f;  // window vs reference band is checked every 32 rows; both centres move <= 17 columns in between

DYN_DEV int warp_max_int(int v)
{
#ifndef DYN_HOST_EMU
	return __reduce_max_sync(FULL, v);
#else
	for (int o = 16; o; o >>= 1) v = max(v, __shfl_sync(FULL, v, (int)((threadIdx.x + o) & 31)));
	return v;
#endif
}

DYN_DEV float warp_sum(float v, int lane)
{
#pragma unroll
	for (int o = 16; o; o >>= 1) v += __shfl_sync(FULL, v, (lane + o) & 31);
	return v;
}

// Where the mass of a row sits inside the window.  key: OB + exponent of the lane's largest value (INT_MIN/2 for a lane
// without a non-zero value).  Returns false when nothing is alive.  first / last: relative lane index (0 = the lane that
// holds the window's lowest column) of the lowest / highest lane within G bits of the row maximum.
template <class RC>
DYN_DEV bool mass_extent(int key, bool alive, int mid, int G, int& first, int& last)
{
	const int kmax = warp_max_int(alive ? key : -(1 << 30));
	const unsigned m = __ballot_sync(FULL, alive && key >= kmax - G);
	if (m == 0u) return false;
	const int lane_lo = pmod(mid - RC::HW, RC::SLOTS) / RC::CPL;
	const unsigned r = __funnelshift_r(m, m, lane_lo);  // bit k <=> lane (lane_lo + k) & 31
	first = __ffs(r) - 1;
	last = 31 - __clz(r);
	return true;
}

// centre column (times 2) of relative lanes first .. last of the window centred at mid
template <class RC>
DYN_DEV int extent_centre2(int mid, int first, int last)
{
	const int lo = mid - RC::HW;
	const int ub = lo - pmod(lo, RC::SLOTS) % RC::CPL;  // unwrapped column of the first slot of the window's lowest lane
	return 2 * ub + (first + last) * RC::CPL + RC::CPL - 1;
}

DYN_DEV int fexp(float v) { return ((__float_as_int(v) >> 23) & 0xff) - 127; }

// New integer offset of a lane at a renormalisation (block floating point, one exponent per lane).
//   lm   largest value of the lane (0: nothing alive)     O  current offset     e0  exponent the maximum is brought to
//   DIR  +1: values flow in from lane+1 (backward), -1: from lane-1 (forward, posterior-Viterbi)
// A lane's own maximum goes to [2^e0, 2^(e0+1)), but its offset is never more than RDC below the largest candidate of
// the EIGHT source-side lanes: with C columns per lane a value crosses a lane boundary every C rows, i.e. up to RN / C
// lanes between two renormalisations, and every hop multiplies by 2^(offset difference) — bounding the deficit against
// each lane it can reach (not against the direct neighbour only, as dp_linear.cuh can afford with 13 columns per lane)
// keeps whatever flows in below 2^(RDC + growth of RN rows).  A lane with nothing alive within reach parks
// RDC below the row maximum.  All 32 lanes must call.
template <int DIR>
DYN_DEV int ring_offset(float lm, int O, int e0, int lane)
{
	constexpr int NONE = -(1 << 29);
	const bool alive = lm > 0.0f && lm < 3.0e38f;
	const int cand = alive ? O + fexp(lm) - e0 : NONE;
	int r = __shfl_sync(FULL, cand, (lane + DIR) & 31);
	r = max(r, __shfl_sync(FULL, r, (lane + DIR) & 31));
	r = max(r, __shfl_sync(FULL, r, (lane + 2 * DIR) & 31));
	r = max(r, __shfl_sync(FULL, r, (lane + 4 * DIR) & 31));
	const int kmax = warp_max_int(cand);
	int nO = max(cand, r - RDC);
	if (nO < NONE / 2) nO = (kmax < NONE / 2) ? O : kmax - RDC;
	return nO;
}

// lane-local renormalisation of the backward values by an exact power of two
template <class RC>
DYN_DEV void bwd_renorm(Warp<RC>& w, lin::BwdL<RC::CPL>& b)
{
	constexpr int C = RC::CPL;
	float lm = b.bE[0];
#pragma unroll
	for (int j = 1; j < C; ++j) lm = fmaxf(lm, b.bE[j]);
	const int nO = ring_offset<+1>(lm, b.OB, 0, w.lane);
	const float sc = lin::pow2i(b.OB - nO);
#pragma unroll
	for (int j = 0; j < C; ++j)
	{
		b.bM[j] *= sc;
		b.bE[j] *= sc;
	}
	b.OB = nO;
	const int obr = __shfl_sync(FULL, nO, (w.lane + 1) & 31);
	lin::pow2_split(obr - nO, b.sR1, b.sR2);
}

template <class RC>
DYN_DEV void bwd_init_terminal(Warp<RC>& w, lin::BwdL<RC::CPL>& b)
{
	constexpr int C = RC::CPL;
#pragma unroll
	for (int j = 0; j < C; ++j)
	{
		b.bM[j] = 0.0f;
		b.bE[j] = 0.0f;
	}
	const int q = pmod((int)w.N - 1, RC::SLOTS);
	with_slot<C>(w.lane, q, SetOne<C>{b.bE, 1.0f});  // bE[T-1][N-1] = 1 (NT:170)
	b.OB = (w.lane == q / C) ? 0 : -RDC;
	const int obr = __shfl_sync(FULL, b.OB, (w.lane + 1) & 31);
	lin::pow2_split(obr - b.OB, b.sR1, b.sR2);
}

template <class RC>
DYN_DEV void vit_renorm(Warp<RC>& w, lin::FwdL<RC::CPL>& f)
{
	constexpr int C = RC::CPL;
	float lm = 0.0f;
#pragma unroll
	for (int j = 0; j < C; ++j) lm = lin::max3f(lm, f.VM[j], f.VE[j]);
	// every posterior-Viterbi score underflowed (or is NaN): the decision bits from here on would be meaningless
	if (!__any_sync(FULL, lm > 0.0f && lm < 3.0e38f)) f.fault = true;
	const int nO = ring_offset<-1>(lm, f.OV, lin::E0V, w.lane);
	const float sc = lin::pow2i(f.OV - nO);
#pragma unroll
	for (int j = 0; j < C; ++j)
	{
		f.VM[j] *= sc;
		f.VE[j] *= sc;
	}
	f.OV = nO;
	const int ovl = __shfl_sync(FULL, nO, (w.lane + 31) & 31);
	lin::pow2_split(ovl - nO, f.sV1, f.sV2);
}

struct Sched
{
	unsigned smask;  // bit i <=> the window centre of row base+i differs from that of row base+i+1
	float xv;        // sample base + lane
};

// ------------------------------------------------------------------------------------------------------
// pass 1: backward over the whole read; decides and stores the window schedule.  Returns log2 Zb.
// ------------------------------------------------------------------------------------------------------
template <class RC, bool STORE>
DYN_DEV double backward_pass(Warp<RC>& w, const SlotScratch& sc, const ReadDesc& rd, const BatchArgs& args, bool& fault)
{
	constexpr int C = RC::CPL;
	const float m1 = args.m1_lin, e2 = args.e2_lin;
	const int lane = w.lane;
	const int G = args.rib_guard;
	const int band_margin = (int)rd.bw - RC::HW - BAND_SLACK;  // >= 0 (host)
	lin::BwdL<C> b;
	int mid = (int)w.N - 1;  // = the reference's band centre of row T-1
	w.load_window(mid);
	rib::bwd_init_terminal<RC>(w, b);
	if (STORE && ((w.T - 1) & (RC::CK - 1)) == 0) lin::ckpt_store<RC>(sc, (w.T - 1) / RC::CK, lane, b);
	// row T-1 alone in its 32-row chunk: no row of that chunk is computed below, its schedule entry is the start state
	if (STORE && ((w.T - 1) & 31u) == 0 && lane == 0) sc.sched[(w.T - 1) >> 5] = make_uint2((unsigned)mid, 0u);

	int t = (int)w.T - 2;
	float xnext;
	{
		const uint32_t r = ((uint32_t)t & ~31u) + lane;
		xnext = (r < w.S) ? w.sig[r] : 0.0f;
	}
	while (t >= 0)
	{
		const uint32_t base = (uint32_t)t & ~31u;
		const float xv = xnext;
		if (base >= 32) xnext = w.sig[base - 32 + lane];
		const int nrows = t - (int)base + 1;  // rows base .. t are computed in this chunk, from the state of row t+1
		// ---- controller: how far the window moves down during this chunk ------------------------------------
		unsigned smask = 0u;
		{
			float lm = b.bE[0];
#pragma unroll
			for (int j = 1; j < C; ++j) lm = fmaxf(lm, b.bE[j]);
			const bool alive = lm > 0.0f && lm < 3.0e38f;
			int first = 0, last = 0;
			if (!mass_extent<RC>(b.OB + fexp(lm), alive, mid, G, first, last)) { fault = true; RIB_DBG("p1 t=%d nothing alive\n", t); }
			else
			{
				if (first == 0 || last == 31) { fault = true; RIB_DBG("p1 t=%d mid=%d edge first=%d last=%d\n", t, mid, first, last); }  // the mass touches an edge lane of the window
				const int c2 = extent_centre2<RC>(mid, first, last);
				int s = (2 * mid - c2) / 2;  // window centre above the mass centre: slide down
				s = max(0, min(s, min(nrows, mid)));
				// s slides spread evenly over the chunk's rows (row index i = nrows-1 .. 0)
				const bool bit = lane < nrows && ((lane + 1) * s) / nrows != (lane * s) / nrows;
				smask = __ballot_sync(FULL, bit);
			}
			// the window must stay inside the reference band (NT:96-106) at every row of the chunk
			const int dref = mid - (int)band_mid((uint32_t)t + 1u, w.ratio);
			if (dref > band_margin || dref < -band_margin) { fault = true; RIB_DBG("p1 t=%d mid=%d dref=%d margin=%d\n", t, mid, dref, band_margin); }
			RIB_DBG("p1 t=%d mid=%d first=%d last=%d smask=%08x\n", t, mid, first, last, smask);
		}
		if (fault) return NAN;
#pragma unroll 1
		for (int i = nrows - 1; i >= 0; --i)
		{
			const uint32_t tt = base + i;
			const float x = __shfl_sync(FULL, xv, i);
			lin::bwd_step<RC>(w, b, x, (smask >> i) & 1u, mid, m1, e2);
			if ((tt & (RC::RN - 1)) == 0)
			{
				rib::bwd_renorm<RC>(w, b);
				if (STORE && (tt & (RC::CK - 1)) == 0) lin::ckpt_store<RC>(sc, tt / RC::CK, lane, b);
			}
		}
		if (STORE && lane == 0) sc.sched[base >> 5] = make_uint2((unsigned)mid, smask);
		t = (int)base - 1;
	}
	// column 0 must be inside the window of row 0, and the window inside the reference band there as well
	if (mid > RC::HW || mid > band_margin) { fault = true; RIB_DBG("p1 end mid=%d\n", mid); }
	if (fault) return NAN;
	// Zb = bE[0][0] (NT:286): column 0 is ring slot 0 = lane 0, j 0
	const double z = log2((double)b.bE[0]) + (double)b.OB;
	return shfl_f64(z, 0);
}

// window centre of row t from the stored schedule
DYN_DEV int sched_mid(const SlotScratch& sc, uint32_t t)
{
	const uint2 s = sc.sched[t >> 5];
	return (int)s.x + __popc(s.y & ((1u << (t & 31u)) - 1u));
}

template <class RC>
DYN_DEV Sched sched_load(const Warp<RC>& w, const SlotScratch& sc, uint32_t base)
{
	Sched c;
	const uint32_t r = base + w.lane;
	c.smask = (base < w.T) ? sc.sched[base >> 5].y : 0u;
	c.xv = (r < w.S) ? w.sig[r] : 0.0f;
	return c;
}

// ------------------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------------------
template <int C>
struct TrainAcc
{
	float gw[C], gx[C], gxx[C];  // per ring slot: sum gamma, sum gamma*(x - mu), sum gamma*(x - mu)^2
	float mu[C];                 // centre: the model mean of the slot's kmer (FP32)
	float sM, sE;                // chunk sums of the match / extend posteriors of this lane
	double dM, dE;
};

struct RowSink
{
	uint32_t n;     // records written so far
	uint32_t cap;
	bool overflow;
};

template <class RC>
struct SmemR
{
	float* bE;  // [(CK+2)][C][32]
	int* OB;    // [NRN][32]  backward lane offsets in force for the rows up to and including t_lo + i*RN
	DYN_DEV explicit SmemR(unsigned char* p)
	{
		OB = reinterpret_cast<int*>(p);
		bE = reinterpret_cast<float*>(p + (size_t)RC::NRN * 32 * 4);
	}
};

// write the training statistics of ring slot j of this lane for lattice column col and clear them
template <class RC>
DYN_DEV void flush_slot(const BatchArgs& args, const ReadDesc& rd, TrainAcc<RC::CPL>& a, int j, int col)
{
	constexpr int C = RC::CPL;
#pragma unroll
	for (int jj = 0; jj < C; ++jj)
		if (jj == j)
		{
			if (a.gw[jj] > 0.0f)
			{
				// un-centre in double: sum g*x = gx + mu*gw, sum g*x^2 = gxx + 2*mu*gx + mu^2*gw
				const double mu = (double)a.mu[jj], gw = (double)a.gw[jj], gx = (double)a.gx[jj], gxx = (double)a.gxx[jj];
				args.read_w[rd.pc_off + col] = gw;
				args.read_x[rd.pc_off + col] = gx + mu * gw;
				args.read_xx[rd.pc_off + col] = gxx + 2.0 * mu * gx + mu * mu * gw;
			}
			a.gw[jj] = 0.0f;
			a.gx[jj] = 0.0f;
			a.gxx[jj] = 0.0f;
		}
}

// forward renormalisation (own maximum -> [1, 2), coupled to the left lane), range guard and window-edge guard
template <class RC>
DYN_DEV void fwd_renorm(Warp<RC>& w, lin::FwdL<RC::CPL>& f, const float (&brow)[RC::CPL], int OB, int Z2i, float c0, int mid, int G)
{
	constexpr int C = RC::CPL;
	float lm = 0.0f, bm = 0.0f;
#pragma unroll
	for (int j = 0; j < C; ++j) lm = lin::max3f(lm, f.fM[j], f.fE[j]);
#pragma unroll
	for (int j = 0; j < C; ++j) bm = fmaxf(bm, brow[j]);
	const bool alive = lm > 0.0f && lm < 3.0e38f;
	if (alive && bm > 0.0f)
	{
		const int g = fexp(lm) + fexp(bm) + f.OF + OB - Z2i;
		if (g > lin::LIN_GUARD_BITS) f.fault = true;
	}
	int first = 0, last = 0;
	if (!mass_extent<RC>(f.OF + fexp(lm), alive, mid, G, first, last)) { f.fault = true; RIB_DBG("p2 nothing alive mid=%d\n", mid); }
	else if (first == 0 || last == 31) { f.fault = true; RIB_DBG("p2 edge mid=%d first=%d last=%d\n", mid, first, last); }
	const int nO = ring_offset<-1>(lm, f.OF, 0, w.lane);
	const float sc = lin::pow2i(f.OF - nO);
#pragma unroll
	for (int j = 0; j < C; ++j)
	{
		f.fM[j] *= sc;
		f.fE[j] *= sc;
	}
	f.OF = nO;
	const int ofl = __shfl_sync(FULL, nO, (w.lane + 31) & 31);
	lin::pow2_split(ofl - nO, f.sL1, f.sL2);
	f.kap = lin::kappa(nO, OB, Z2i, c0);
}

// One row of pass 2.  On entry f holds the forward values of row t and the Viterbi values of row t-1.
//   DO_V:    posteriors of row t, posterior-Viterbi update (NT:357-362), decision bits + records (MODE 1) or training
//            statistics (MODE 2; xprev = x[t-1], the sample the posteriors of row t weigh, NT:509-512)
//   DO_STEP: forward recurrence to row t+1 (NT:141-150) incl. the window slide between t and t+1
// Returns the posterior mass of this lane's cells.
template <class RC, int MODE, bool DO_V, bool DO_STEP>
DYN_DEV float fwd_row(Warp<RC>& w, lin::FwdL<RC::CPL>& f, const SlotScratch& sc, RowSink& rs, TrainAcc<RC::CPL>& ta,
	const BatchArgs& args, const ReadDesc& rd, uint32_t t, float x, float xprev, bool slide, int& mid_f,
	const float (&bc)[RC::CPL], const float (&bn)[RC::CPL], float kapE, float kapM)
{
	constexpr int C = RC::CPL;
	constexpr int HW = RC::HW;
	const int lane = w.lane;
	const float m1 = args.m1_lin, e2 = args.e2_lin;
	// the values the right lane needs are those of the previous row: send them first, consume them last
	const float vlraw = (DO_V && MODE == 1) ? __shfl_sync(FULL, f.VE[C - 1], (lane + 31) & 31) : 0.0f;
	const float flraw = DO_STEP ? __shfl_sync(FULL, f.fE[C - 1], (lane + 31) & 31) : 0.0f;
	if (DO_STEP && slide)
	{
		// column entering window(t+1): its ring slot was the dead one
		const int nin = mid_f + 1 + HW;
		if (w.valid_col(nin))
		{
			const PosConst v = w.pc[nin];
			with_slot<C>(lane, pmod(nin, RC::SLOTS), [&](int j) {
#pragma unroll
				for (int jj = 0; jj < C; ++jj)
					if (jj == j)
					{
						w.em.a[jj] = v.a;
						w.em.b[jj] = v.b;
						w.em.c[jj] = v.c;
						if (MODE == 2) ta.mu[jj] = v.pad;
					}
			});
		}
	}
	float p[C], PM[C], PE[C];
	float msum = 0.0f;
#pragma unroll
	for (int j = 0; j < C; ++j) p[j] = DO_STEP ? ex2(w.emis(j, x)) : 0.0f;
	if (DO_V)
	{
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			PE[j] = f.fE[j] * (bc[j] * kapE);
			// bM[t][n] = bE[t+1][n] * p(t,n) (NT:200); the last row has no match state
			PM[j] = DO_STEP ? f.fM[j] * ((bn[j] * p[j]) * kapM) : 0.0f;
		}
		if (DO_STEP && slide && mid_f - HW >= 0)
		{
			// the window's lowest column is outside window(t+1): bE[t+1][lo] counts as 0 there
			with_slot<C>(lane, pmod(mid_f - HW, RC::SLOTS), [&](int j) {
#pragma unroll
				for (int jj = 0; jj < C; ++jj)
					if (jj == j) PM[jj] = 0.0f;
			});
		}
#pragma unroll
		for (int j = 0; j < C; ++j) msum += PM[j] + PE[j];
		if (MODE == 1)
		{
			// posterior-Viterbi fill (NT:357-362) as a max-product, in place from the highest slot down;
			// decision bit set <=> the E state of this cell is entered from E (the test of NT:448 at fill time)
			const float vl = (vlraw * f.sV1) * f.sV2;
			unsigned bits[C];
			float lmax = 0.0f;
#pragma unroll
			for (int j = C - 1; j >= 0; --j)
			{
				const float vmx = fmaxf(f.VM[j], f.VE[j]);
				bits[j] = __ballot_sync(FULL, f.VM[j] < f.VE[j]);
				const float left = (j > 0) ? f.VE[j - 1] : vl;
				f.VM[j] = left * PM[j];
				f.VE[j] = vmx * PE[j];
				lmax = lin::max3f(lmax, PM[j], PE[j]);
			}
			if ((t & (RC::RV - 1)) == 0) rib::vit_renorm<RC>(w, f);
			// sparse posterior records: one per lane that holds a posterior above the threshold (NaN counts as hot)
			const bool hot = !(lmax <= args.thr_rib);
			const unsigned hm = __ballot_sync(FULL, hot);
			const uint32_t pos = rs.n + __popc(hm & ((1u << lane) - 1u));
			if (hot && pos < rs.cap)
			{
				float tmp[RC::RECF];
#pragma unroll
				for (int j = 0; j < C; ++j)
				{
					tmp[j] = PM[j];
					tmp[C + j] = PE[j];
				}
#pragma unroll
				for (int q = 2 * C; q < RC::RECF; ++q) tmp[q] = 0.0f;
				float4* dst = reinterpret_cast<float4*>(static_cast<float*>(sc.recs) + (size_t)pos * RC::RECF);
#pragma unroll
				for (int q = 0; q < RC::RECF / 4; ++q) dst[q] = make_float4(tmp[4 * q], tmp[4 * q + 1], tmp[4 * q + 2], tmp[4 * q + 3]);
			}
			if (lane == 0)
			{
				uint32_t h[RC::HDRW];
				h[0] = rs.n;
				h[1] = hm;
#pragma unroll
				for (int j = 0; j < C; ++j) h[2 + j] = bits[j];
#pragma unroll
				for (int q = 2 + C; q < RC::HDRW; ++q) h[q] = 0u;
				uint4* dst = reinterpret_cast<uint4*>(sc.hdr + (size_t)t * RC::HDRW);
#pragma unroll
				for (int q = 0; q < RC::HDRW / 4; ++q) dst[q] = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
			}
			const uint32_t nn = rs.n + __popc(hm);
			rs.overflow = rs.overflow || (nn > rs.cap);
			rs.n = (nn > rs.cap) ? rs.cap : nn;
		}
		else
		{
			// training statistics (NT:494-514): gamma = pM + pE weighs sample x[t-1] for the kmer of the cell's column
			float sm_ = 0.0f, se_ = 0.0f;
#pragma unroll
			for (int j = 0; j < C; ++j)
			{
				const float g = PM[j] + PE[j];
				const float dx = xprev - ta.mu[j];
				const float gd = g * dx;
				ta.gw[j] += g;
				ta.gx[j] += gd;
				ta.gxx[j] = fmaf(gd, dx, ta.gxx[j]);
				sm_ += PM[j];
				se_ += PE[j];
			}
			ta.sM += sm_;
			ta.sE += se_;
		}
	}
	if (DO_STEP)
	{
		const float fl = (flraw * f.sL1) * f.sL2;
#pragma unroll
		for (int j = C - 1; j >= 0; --j)
		{
			const float left = (j > 0) ? f.fE[j - 1] : fl;
			const float ne = fmaf(f.fE[j], e2, f.fM[j]) * p[j];  // (fM + fE*e2) * p    (NT:146-150, e1 = 1)
			f.fM[j] = left * (p[j] * m1);                        // fE[t][n-1] * p * m1  (NT:143)
			f.fE[j] = ne;
		}
		if (slide)
		{
			const int nold = mid_f - HW;  // column of window(t) that is not in window(t+1)
			if (nold >= 0)
				with_slot<C>(lane, pmod(nold, RC::SLOTS), [&](int j) {
					if (MODE == 2 && nold >= 1) flush_slot<RC>(args, rd, ta, j, nold);
#pragma unroll
					for (int jj = 0; jj < C; ++jj)
						if (jj == j)
						{
							w.em.a[jj] = 0.0f;
							w.em.b[jj] = 0.0f;
							w.em.c[jj] = CNEG;
							f.fM[jj] = 0.0f;
							f.fE[jj] = 0.0f;
						}
				});
			++mid_f;
		}
	}
	return msum;
}

// pass 2: recomputation + forward + posterior (+ posterior-Viterbi fill | training statistics).
// Returns log2 Zf - log2 Zb (NaN on a fault).
template <class RC, int MODE>
DYN_DEV double forward_pass(Warp<RC>& w, const SlotScratch& sc, const BatchArgs& args, const ReadDesc& rd,
	unsigned char* smem_raw, double Z2, uint32_t& nrec_out, bool& overflow, double& xi_m, double& xi_e)
{
	constexpr int C = RC::CPL;
	constexpr int CK = RC::CK;
	constexpr int RN = RC::RN;
	constexpr int ROWF = RC::ROWF;
	constexpr int HW = RC::HW;
	const float m1 = args.m1_lin, e2 = args.e2_lin;
	const int G = args.rib_guard;
	SmemR<RC> sm(smem_raw);
	lin::FwdL<C> f;
	lin::BwdL<C> b;
	TrainAcc<C> ta;
	const int lane = w.lane;
	const uint32_t T = w.T;
	RowSink rs;
	rs.n = 0;
	rs.cap = (uint32_t)args.rec_cap;
	rs.overflow = false;
	// posterior = sf * sb * 2^(OF + OB - Z2) = sf * sb * c0 * 2^(OF + OB - Z2i); c0 absorbs the closed-loop correction
	const double Z2f = floor(Z2);
	const int Z2i = (int)Z2f;
	float c0 = (float)exp2(Z2f - Z2);
	bool massfault = false;
	float macc = 0.0f, mcnt = 0.0f;  // posterior mass of this lane's cells / rows since the last closed-loop check

	int mid_f = sched_mid(sc, 0);
	float bc[C], bn[C];
#pragma unroll
	for (int j = 0; j < C; ++j)
	{
		bc[j] = bn[j] = 0.0f;
		ta.gw[j] = ta.gx[j] = ta.gxx[j] = 0.0f;
		ta.mu[j] = 0.0f;
	}
	ta.sM = ta.sE = 0.0f;
	ta.dM = ta.dE = 0.0;
	w.load_window(mid_f);
	if (MODE == 2)
	{
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			const int n = w.col_of_slot(lane * C + j, mid_f - HW);
			ta.mu[j] = (n >= 0 && n <= min(mid_f + HW, (int)w.N - 1)) ? w.pc[n].pad : 0.0f;
		}
	}
	const uint32_t kb = (T - 1) / CK;
	Sched cur = sched_load<RC>(w, sc, 0);
	Sched nxtc = sched_load<RC>(w, sc, 32);
	float xprev = 0.0f;

	for (uint32_t k = 0; k <= kb; ++k)
	{
		const uint32_t t_lo = k * CK;
		const uint32_t t_hi = t_lo + CK;
		if (k > 0 && (t_lo & 31u) == 0)
		{
			if (MODE == 2)
			{
				ta.dM += (double)ta.sM;
				ta.dE += (double)ta.sE;
				ta.sM = ta.sE = 0.0f;
			}
			cur = nxtc;
			nxtc = sched_load<RC>(w, sc, t_lo + 32);
		}
		// ---- step a: recompute the backward rows of this block into shared memory -----------------------
		// sm.OB[i] = backward offsets in force for rows t_lo + (i-1)*RN + 1 .. t_lo + i*RN
		const bool from_ckpt = (t_hi <= T - 1);
		const uint32_t src_row = from_ckpt ? t_hi : T - 1;
		int mid_b = sched_mid(sc, src_row);
		w.slide_window_up(mid_f, mid_b);
		if (from_ckpt) lin::ckpt_load<RC>(sc, k + 1, lane, b);
		else rib::bwd_init_terminal<RC>(w, b);
		{
			float* dst = sm.bE + (size_t)(src_row - t_lo) * ROWF;
#pragma unroll
			for (int j = 0; j < C; ++j) dst[j * 32 + lane] = b.bE[j];
			sm.OB[((src_row - t_lo + RN - 1) / RN) * 32 + lane] = b.OB;
		}
#pragma unroll 1
		for (int tt = (int)src_row - 1; tt >= (int)t_lo; --tt)
		{
			const int i = tt & 31;
			const float x = __shfl_sync(FULL, cur.xv, i);
			const bool sl = (cur.smask >> i) & 1u;
			if (sl)
			{
				// the column that enters the window at row tt: (tt+1, n) is outside the window, but its ring slot (the dead
				// one) kept the ungated M-transition term (dp_linear.cuh bwd_step); zero it in the stored row tt+1
				const int nb = mid_b - 1 - HW;
				if (nb >= 0)
				{
					const int q = pmod(nb, RC::SLOTS);
					if (lane == q / C) sm.bE[(size_t)(tt + 1 - (int)t_lo) * ROWF + (q % C) * 32 + lane] = 0.0f;
				}
			}
			lin::bwd_step<RC>(w, b, x, sl, mid_b, m1, e2);
			if ((tt & (RN - 1)) == 0)
			{
				rib::bwd_renorm<RC>(w, b);
				sm.OB[((tt - (int)t_lo) / RN) * 32 + lane] = b.OB;
			}
			float* dst = sm.bE + (size_t)(tt - (int)t_lo) * ROWF;
#pragma unroll
			for (int j = 0; j < C; ++j) dst[j * 32 + lane] = b.bE[j];
		}
		__syncwarp();
		// the emission window is window(t_lo) again (mid_b == mid_f); the training centres follow the emission constants
		if (MODE == 2 && mid_b != mid_f) massfault = true;  // cannot happen: the schedule is replayed exactly

		// ---- step b: forward rows t_lo .. min(t_hi, T-1) - 1 ----------------------------------------------
		uint32_t t = t_lo;
		if (k == 0)
		{
			// row 0: fE[0][0] = 1 (NT:120), VE[0][0] = 1 (NT:336)
#pragma unroll
			for (int j = 0; j < C; ++j)
			{
				f.fM[j] = 0.0f;
				f.fE[j] = 0.0f;
				f.VM[j] = 0.0f;
				f.VE[j] = 0.0f;
			}
			f.OF = (lane == 0) ? 0 : -RDC;
			f.OV = (lane == 0) ? 0 : -RDC;
			f.fault = false;
			if (lane == 0)
			{
				f.fE[0] = 1.0f;
				f.VE[0] = 1.0f;
			}
			{
				const int ofl = __shfl_sync(FULL, f.OF, (lane + 31) & 31);
				lin::pow2_split(ofl - f.OF, f.sL1, f.sL2);
				const int ovl = __shfl_sync(FULL, f.OV, (lane + 31) & 31);
				lin::pow2_split(ovl - f.OV, f.sV1, f.sV2);
			}
			const float x0 = __shfl_sync(FULL, cur.xv, 0);
			rib::fwd_row<RC, MODE, false, true>(w, f, sc, rs, ta, args, rd, 0, x0, 0.0f, cur.smask & 1u, mid_f, bc, bn, 0.0f, 0.0f);
			xprev = x0;
			f.kap = lin::kappa(f.OF, sm.OB[32 + lane], Z2i, c0);  // rows 1 .. RN
			t = 1;
		}
		const uint32_t t_end = min(t_hi, T - 1);
#pragma unroll 1
		while (t < t_end)
		{
			const int i = t & 31;
			const uint32_t r = t - t_lo;
			const bool rn_row = (t & (RN - 1)) == 0;
			const float* row = sm.bE + (size_t)r * ROWF;
#pragma unroll
			for (int j = 0; j < C; ++j)
			{
				bc[j] = row[j * 32 + lane];
				bn[j] = row[ROWF + j * 32 + lane];
			}
			// on a renormalisation row bM[t] = bE[t+1] * p lives in the offsets of the next rows
			float kapN = rn_row ? lin::kappa(f.OF, sm.OB[(r / RN + 1) * 32 + lane], Z2i, c0) : f.kap;
			const float x = __shfl_sync(FULL, cur.xv, i);
			const float ms = rib::fwd_row<RC, MODE, true, true>(w, f, sc, rs, ta, args, rd, t, x, xprev, (cur.smask >> i) & 1u, mid_f,
				bc, bn, f.kap, kapN);
			xprev = x;
			macc += ms;
			mcnt += 1.0f;
			if (rn_row)
			{
				// closed loop: the posterior mass of every row is 1.  The masses of the rows since the last check are summed
				// per lane and reduced once: any row that lost more than RIB_MASS_TOL of its mass is a fault, and the mean
				// deviation (slow common-mode FP32 drift) is folded into the posterior factor
				const float mass = warp_sum(macc, lane);
				if (!(fabsf(mass - mcnt) <= RIB_MASS_TOL)) { massfault = true; RIB_DBG("p2 t=%u mass=%g of %g\n", t, mass, mcnt); }
				const float inv = mcnt / mass;
				c0 *= inv;
				kapN *= inv;
				macc = 0.0f;
				mcnt = 0.0f;
			}
			f.kap = kapN;
			// the forward values are now those of row t+1
			if (((t + 1) & (RN - 1)) == 0)
			{
				const float* rown = sm.bE + (size_t)(r + 1) * ROWF;
				float br[C];
#pragma unroll
				for (int j = 0; j < C; ++j) br[j] = rown[j * 32 + lane];
				rib::fwd_renorm<RC>(w, f, br, sm.OB[((r + 1) / RN) * 32 + lane], Z2i, c0, mid_f, G);
			}
			++t;
		}
		__syncwarp();
	}
	{
		const float* row = sm.bE + (size_t)((T - 1) - kb * CK) * ROWF;
#pragma unroll
		for (int j = 0; j < C; ++j) bc[j] = row[j * 32 + lane];
		const float ms = rib::fwd_row<RC, MODE, true, false>(w, f, sc, rs, ta, args, rd, T - 1, 0.0f, xprev, false, mid_f, bc, bn, f.kap, 0.0f);
		macc += ms;
		mcnt += 1.0f;
		const float mass = warp_sum(macc, lane);
		if (!(fabsf(mass - mcnt) <= RIB_MASS_TOL)) massfault = true;
	}
	// Zf = fE[T-1][N-1] (NT:285)
	float v = 0.0f;
	with_slot<C>(lane, pmod((int)w.N - 1, RC::SLOTS), GetOne<C>{f.fE, v});
	const int ql = pmod((int)w.N - 1, RC::SLOTS) / C;
	double dz = log2((double)v) + (double)f.OF - Z2;
	dz = shfl_f64(dz, ql);
	RIB_DBG("p2 end dz=%g\n", dz);
	if (__any_sync(FULL, f.fault || massfault)) dz = NAN;
	if (MODE == 2)
	{
		// columns still inside the window
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			const int n = w.col_of_slot(lane * C + j, mid_f - HW);
			if (n >= 1 && n <= min(mid_f + HW, (int)w.N - 1)) flush_slot<RC>(args, rd, ta, j, n);
		}
		double sm_ = ta.dM + (double)ta.sM, se_ = ta.dE + (double)ta.sE;
		for (int o = 16; o; o >>= 1)
		{
			sm_ += shfl_f64(sm_, (lane + o) & 31);
			se_ += shfl_f64(se_, (lane + o) & 31);
		}
		// #(E->M) = sum pM; #(E->E) = sum pE - sum pM (dp_kernels.cuh train_stats_pass)
		xi_m = sm_;
		xi_e = se_ - sm_;
	}
	nrec_out = rs.n;
	overflow = rs.overflow;
	return dz;
}

// ------------------------------------------------------------------------------------------------------
// pass 3: traceback over the decision bits (NT:383-456), posterior of every path cell, per-segment medians
// ------------------------------------------------------------------------------------------------------
template <class RC>
DYN_DEV bool traceback_pass(Warp<RC>& w, const SlotScratch& sc, const BatchArgs& args, const ReadDesc& rd)
{
	constexpr int C = RC::CPL;
	constexpr int SLOTS = RC::SLOTS;
	constexpr int HDRW = RC::HDRW;
	const int lane = w.lane;
	const uint32_t T = w.T, N = w.N;
	uint32_t* border = args.out_sigpos + rd.out_off;  // Kc = N-1 entries
	const float* recs = static_cast<const float*>(sc.recs);
	int t = (int)T - 1, n = (int)N - 1;
	int inM = 0;
	bool done = false;
	while (!done)
	{
		const int cbase = t & ~31;
		const uint32_t row = (uint32_t)cbase + lane;
		// every lane holds the header of one row of the chunk
		uint32_t h[HDRW];
		{
			const uint4* src = reinterpret_cast<const uint4*>(sc.hdr + (size_t)row * HDRW);
#pragma unroll
			for (int q = 0; q < HDRW / 4; ++q)
			{
				const uint4 v = (row < T && row >= 1) ? src[q] : make_uint4(0u, 0u, 0u, 0u);
				h[4 * q] = v.x;
				h[4 * q + 1] = v.y;
				h[4 * q + 2] = v.z;
				h[4 * q + 3] = v.w;
			}
		}
		uint32_t mycell = 0xffffffffu;  // path cell of this lane's row: column | match state << 31
		while (t >= cbase)
		{
			if (t == 0 || n == 0)
			{
				done = true;
				break;
			}
			if (inM)
			{
				// match state at (t, n): emits the segment border (NT:416-440)
				if (lane == t - cbase) mycell = (uint32_t)n | 0x80000000u;
				if (lane == 0) border[n - 1] = (uint32_t)t - 1;  // Segment.signalPosition (NT:424-430)
				--t;
				--n;
				inM = 0;
				continue;
			}
			// extension run in column n: rows t, t-1, ... down to the row whose E cell was entered from M (NT:443-451)
			const int q = n % SLOTS;
			const int ql = q / C, j = q - ql * C;
			uint32_t word = 0u;
#pragma unroll
			for (int jj = 0; jj < C; ++jj)
				if (jj == j) word = h[2 + jj];
			const bool valid = (int)row <= t && row >= 1;
			const bool from_m = valid && (((word >> ql) & 1u) == 0u);
			const unsigned mk = __ballot_sync(FULL, from_m);
			int tstar;
			if (mk)
			{
				tstar = cbase + (31 - __clz(mk));
				inM = 1;
			}
			else
				tstar = max(cbase, 1);
			if ((int)row >= tstar && (int)row <= t) mycell = (uint32_t)n;
			t = tstar - 1;
		}
		if (mycell != 0xffffffffu)
		{
			const uint32_t col = mycell & 0x7fffffffu;
			const bool isM = (mycell >> 31) != 0;
			const int q = (int)(col % SLOTS);
			const int ql = q / C, j = q - ql * C;
			float pv = 0.0f;
			if ((h[1] >> ql) & 1u)
			{
				const uint32_t idx = h[0] + __popc(h[1] & ((1u << ql) - 1u));
				pv = recs[(size_t)idx * RC::RECF + (isM ? 0 : C) + j];
			}
			sc.pp[row] = pv;
		}
		__syncwarp();
	}
	if (!((n == 0) && !inM)) return false;
	__threadfence_block();
	__syncwarp();
	segment_medians<RC>(w, sc, args, rd);
	return true;
}

// one read, all passes.  A read the ribbon cannot represent leaves with ST_LIN_FAULT (host: full-band kernels).
template <class RC, int MODE>
DYN_DEV void ribbon_read(const BatchArgs& args, const ReadDesc& rd, uint32_t ridx, const SlotScratch& sc,
	unsigned char* smem_raw, int lane)
{
	Warp<RC> w;
	w.lane = lane;
	w.S = rd.S;
	w.T = rd.S + 1;
	w.N = rd.N;
	w.bw = RC::HW;
	w.ratio = rd.ratio;
	w.sig = args.signal + rd.sig_off;
	w.pc = args.pc + rd.pc_off;
	w.m1 = args.m1;
	w.e2 = args.e2;
	w.ua = 0.0f;
	w.uc = 0.0f;

	ReadOut out;
	out.Z = 0.0;
	out.dZ = 0.0;
	out.nrec = 0;
	out.status = ST_OK;
	out.xi_m = 0.0;
	out.xi_e = 0.0;

	bool fault = false;
	const double Z2 = (MODE == 0) ? rib::backward_pass<RC, false>(w, sc, rd, args, fault) : rib::backward_pass<RC, true>(w, sc, rd, args, fault);
	out.Z = Z2 * LN2;
	if (fault || !(Z2 > -1.0e30 && Z2 < 1.0e30)) out.status = ST_LIN_FAULT;
	else if (MODE != 0)
	{
		__threadfence_block();
		__syncwarp();
		uint32_t nrec = 0;
		bool overflow = false;
		const double dz2 = rib::forward_pass<RC, MODE>(w, sc, args, rd, smem_raw, Z2, nrec, overflow, out.xi_m, out.xi_e);
		out.nrec = nrec;
		out.dZ = dz2 * LN2;
		if (!(fabs(dz2) <= lin::LIN_Z_TOL)) out.status = ST_LIN_FAULT;
		else if (overflow) out.status = ST_LIN_FAULT;  // pathological record density: the full-band kernels have their own retry
		else if (MODE == 1)
		{
			__threadfence_block();
			__syncwarp();
			if (!rib::traceback_pass<RC>(w, sc, args, rd)) out.status = ST_LIN_FAULT;
		}
	}
	if (lane == 0) args.out[ridx] = out;
}

} // namespace rib
} // namespace dyn
