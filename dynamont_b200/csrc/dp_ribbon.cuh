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
// the probability mass.
//
// Everything is organised in GROUPS of 8 samples (sample t = the transition between lattice rows t and t+1; group g =
// samples 8g .. 8g+7, i.e. rows 8g .. 8g+8): the window is constant inside a group and moves by 0 .. 8 columns at a group
// boundary (the alignment path moves at most one column per row, so the window can always keep up); renormalisation,
// the window controller, every consistency check and the checkpoints happen at group boundaries only, so the eight
// rows of a group are one straight-line block of code.
//
//   * pass 1 (backward, groups descending) decides at the end of every group where the next window goes: it centres it
//     on the lanes whose largest value is within 2^-G of the row maximum.  The window centre of every group is stored
//     (4 bytes per 8 rows) together with the checkpoint of the group's top row, already moved into that window.
//   * pass 2 replays exactly that schedule group by group: recompute the backward rows of the group from its
//     checkpoint into shared memory (bit-identical to pass 1), then the forward rows with the posteriors, the
//     posterior-Viterbi fill (decision bits + sparse posterior records) or the training statistics.
//   * nothing is assumed, everything is checked on the device: (i) in both directions the two edge lanes of the window
//     must stay more than G bits below the row maximum at every group boundary, (ii) the window must lie inside the
//     reference band (so the cells it drops are negligible ones, never cells the reference forces to -inf), (iii) Zf must
//     equal Zb, (iv) the posterior mass of every row must be 1 (summed per group; one-sided losses show up there because
//     the two directions are windowed independently of each other's values), (v) largest forward x largest backward
//     value of a row must stay below 2^LIN_GUARD_BITS * Z (what a flushed cell can have carried), (vi) the
//     posterior-Viterbi scores must not die.  A read that fails any check leaves with ST_LIN_FAULT and the host re-runs it
//     through the full-band kernels (same GPU).  Short reads (band narrower than the window) go there directly.
//
// Arithmetic: the linear-domain recurrences of dp_linear.cuh (one MUFU per cell-update), block floating point with one
// integer exponent per lane, but every lane's offset is kept within RDC bits of the row's largest: a lane that was empty
// at one boundary can hold the ridge eight rows later (C = 2: a lane is crossed in two rows), so what flows in must
// fit whatever the lane's scale is.  Posteriors are kept normalised by a closed loop (the measured mass of a group's
// rows corrects the posterior factor of the next group), which removes the common-mode FP32 drift.
// Training statistics (NT:494-514) are accumulated in registers per ring slot (a slot holds one column, i.e. one kmer,
// for as long as the column is inside the window), centred on the model mean, and written once when the column leaves
// the window: no per-cell atomics, no records, no traceback.
#pragma once

#include "dp_linear.cuh"

namespace dyn
{
namespace rib
{

// LOG_: the same passes in the LOG2 DOMAIN (stored value L = log2(true value) - lane offset, "zero" = NEG, sums by
// log-sum-exp on MUFU): the tier for reads the linear-domain ribbon cannot represent — an alignment that leaves the
// reference band is forced through emissions of 2^-600 per row, which no FP32 product survives.  Rows run one by one with
// a lane-local renormalisation after every row; align only, records-free layout (MODE 0 / 3 / 4).
template <int C_, bool LOG_ = false, int GR_ = 8>
struct RCfg
{
	static constexpr bool LOGD = LOG_;
	static constexpr int CPL = C_;
	static constexpr int SLOTS = 32 * C_;
	static constexpr int HW = (SLOTS - 2) / 2;  // live columns of a group: [mid - HW, mid + HW]; one ring slot stays dead
	static constexpr int GR = GR_;              // samples per group (8 or 16: lanes 0 .. GR hold a group's samples / path cells)
	static_assert(GR_ == 8 || GR_ == 16, "group size");
	static constexpr int CKF = 2 * C_ * 32;     // floats per checkpoint
	static constexpr int ROWF = C_ * 32;        // floats per shared-memory row
	static constexpr int HDRW = (2 + C_ + 3) / 4 * 4;  // words of a row header: first record, hot-lane mask, C decision words
	static constexpr int RECF = (2 * C_ + 3) / 4 * 4;  // floats of a lane record: C match + C extend posteriors
	// backward rows 8g .. 8g+8 of the current group (log2 domain: + the lane offsets in force for every row)
	static constexpr size_t SMEM_BYTES = (size_t)(GR + 1) * (ROWF * 4 + (LOG_ ? 32 * 4 : 0));
};

#if defined(DYN_HOST_EMU) && defined(DYN_RIB_DEBUG)
#define RIB_DBG(...) do { if (threadIdx.x == 0) fprintf(stderr, __VA_ARGS__); } while (0)
#else
#define RIB_DBG(...) do { } while (0)
#endif

// The reference forces every cell outside its band (NT:96-106) to -inf.  The window may reach beyond the band as long as
// the MASS does not: the lanes within 2^-G of the row maximum must lie inside the band at every group boundary with this
// many columns to spare (the alignment moves <= 8 columns and the band centre <= 5 during a group), so that whatever
// the window holds outside the band is below 2^-G of the row maximum — the same level the window edges are held to.
constexpr int BAND_SLACK = 14;

// columns covered by relative lanes first .. last of the window centred at mid
template <class RC>
DYN_DEV void extent_columns(int mid, int first, int last, int& col_lo, int& col_hi)
{
	const int lo = mid - RC::HW;
	const int ub = lo - pmod(lo, RC::SLOTS) % RC::CPL;
	col_lo = ub + first * RC::CPL;
	col_hi = ub + last * RC::CPL + RC::CPL - 1;
}
// Largest offset deficit of a lane against the row's largest lane (forward / backward values).  The posterior factor
// 2^(OF + OB - Z) must stay a normal float when a lane that was empty at the last boundary holds the ridge in BOTH
// directions: 2 * (RDC + 12) < 126.  Cells more than 2^-(126 + RDC) below the row maximum flush to zero — far below the
// 2^-G the window guard already treats as nothing; what a flushed cell can have carried is bounded by check (v).
constexpr int RDC = 30;
constexpr int RDCV = 60;  // the same for the posterior-Viterbi scores (products of posteriors: they only shrink)
// |sum of the posterior masses of a group's rows - number of rows| above this is a fault (typical: < 1e-5; a row that
// lost 1e-4 of its mass to an FP32 range problem trips it)
constexpr float RIB_MASS_TOL = 1e-4f;
constexpr int NONE = -(1 << 29);

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

// packed FP32x2 arithmetic (sm_100a FMUL2 / FFMA2: one issue slot for two lattice columns; the kernels are bound by issue
// slots).  Plain IEEE round-to-nearest per component: bit-identical to the scalar instructions they replace.
#ifndef DYN_HOST_EMU
DYN_DEV float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
DYN_DEV float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
DYN_DEV float2 mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
DYN_DEV float2 fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif

// a copy of v in a register of its own: a value used straight out of a 128-bit load stays tied to that load's register
// quad, and the FFMA2 that wants it as half of an aligned pair then pays two moves per lattice row
#ifndef DYN_HOST_EMU
DYN_DEV float untie(float v) { return __fadd_rn(v, 0.0f); }  // (v is a positive constant: no -0 to lose)
#else
DYN_DEV float untie(float v) { return v; }
#endif

DYN_DEV int fexp(float v) { return ((__float_as_int(v) >> 23) & 0xff) - 127; }
DYN_DEV bool is_alive(float v) { return v > 0.0f && v < 3.0e38f; }

// the two arithmetic domains behind one set of passes
template <class RC> DYN_DEV float zero_v() { return RC::LOGD ? NEG : 0.0f; }
template <class RC> DYN_DEV float one_v() { return RC::LOGD ? 0.0f : 1.0f; }
template <class RC> DYN_DEV bool alive(float v) { return RC::LOGD ? (v > DEADT) : is_alive(v); }
template <class RC> DYN_DEV int vexp(float v) { return RC::LOGD ? (int)floorf(v) : fexp(v); }  // floor(log2(value))
// value expressed against an offset that is d smaller (d = old offset - new offset)
template <class RC> DYN_DEV float rescale(float v, int d, float pw) { return RC::LOGD ? v + (float)d : v * pw; }
// factor that converts a neighbour lane's value into this lane's scale (dn = neighbour's offset - own offset)
template <class RC> DYN_DEV float nb_factor(int dn) { return RC::LOGD ? (float)dn : lin::pow2i(dn); }

template <int C>
struct Bw
{
	float bM[C], bE[C];
	int OB;    // true value = stored * 2^OB
	float sR;  // 2^(OB(right lane) - OB(this lane))
};

template <int C>
struct Fw
{
	float fM[C], fE[C];  // (true value) * 2^-OF
	float VM[C], VE[C];  // posterior-Viterbi products, (true value) * 2^-OV
	int OF, OV;
	float sL, sV;        // 2^(OF(left lane) - OF(this lane)), the same for OV
};

template <int C>
struct TrainAcc
{
	float gw[C], gx[C], gxx[C];  // per ring slot: sum gamma, sum gamma*(x - mu), sum gamma*(x - mu)^2
	float mu[C];                 // centre: the model mean of the slot's kmer (FP32)
	float sM, sE;                // sums of the match / extend posteriors of this lane since the last group boundary
	double dM, dE;
};

struct RowSink
{
	uint32_t n;     // records written so far
	uint32_t cap;
	bool overflow;
};

// The warp-uniform view of one read plus this lane's emission constants.
template <class RC>
struct RWarp
{
	static constexpr int C = RC::CPL;
	int lane;
	uint32_t T, N, S;
	int bw_ref;       // the reference's half band width (NT:243)
	double ratio;
	float ratio_f;
	const float* sig;
	const PosConst* pc;
	// log2 N(x; mu, sigma) = c - (x*a - b)^2 of the column in each ring slot (dead slot: c = CNEG), kept as a, -b, -c so that
	// a pair of columns is two FFMA2: z = x*a + (-b), -e = z*z + (-c)
	float a[C], nb[C], nc[C];

	// log2 emission of slot j (dead slot: -1e25)
	DYN_DEV float emis_log(int j, float x) const
	{
		const float z = fmaf(x, a[j], nb[j]);
		return -fmaf(z, z, nc[j]);
	}
	// emission probabilities 2^e of the slot pair h (slots 2h, 2h+1) for sample x
	DYN_DEV float2 emis_pair(int h, float2 x2) const
	{
		const float2 z = fma2(x2, make_float2(a[2 * h], a[2 * h + 1]), make_float2(nb[2 * h], nb[2 * h + 1]));
		const float2 ne = fma2(z, z, make_float2(nc[2 * h], nc[2 * h + 1]));  // = -fma(-z, z, c) exactly
		return make_float2(ex2(-ne.x), ex2(-ne.y));
	}
	// lattice column held by ring slot q when the window starts at column lo (may be negative)
	DYN_DEV static int col_of_slot(int q, int lo) { return lo + pmod(q - lo, RC::SLOTS); }
	DYN_DEV void set_dead(int j)
	{
#pragma unroll
		for (int jj = 0; jj < C; ++jj)
			if (jj == j)
			{
				a[jj] = 0.0f;
				nb[jj] = 0.0f;
				nc[jj] = -CNEG;
			}
	}
	DYN_DEV void set_col(int j, const PosConst& v)
	{
#pragma unroll
		for (int jj = 0; jj < C; ++jj)
			if (jj == j)
			{
				a[jj] = untie(v.a);
				nb[jj] = -v.b;
				nc[jj] = -v.c;
			}
	}
	// emission constants of every column of the window centred at mid; everything else dead
	DYN_DEV void load_window(int mid)
	{
		const int lo = mid - RC::HW;
		const int nlast = min(mid + RC::HW, (int)N - 1);
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			const int n = col_of_slot(lane * C + j, lo);
			if (n >= 0 && n <= nlast)
			{
				const PosConst v = pc[n];
				a[j] = untie(v.a);
				nb[j] = -v.b;
				nc[j] = -v.c;
			}
			else
			{
				a[j] = 0.0f;
				nb[j] = 0.0f;
				nc[j] = -CNEG;
			}
		}
	}
};

template <int C>
DYN_DEV void zero_slot(float (&p)[C], float (&q)[C], int j, float z = 0.0f)
{
#pragma unroll
	for (int jj = 0; jj < C; ++jj)
		if (jj == j)
		{
			p[jj] = z;
			q[jj] = z;
		}
}

// Where the mass of a row sits inside the window: relative lane index (0 = the lane that holds the window's lowest
// column) of the lowest / highest lane whose candidate exponent is within G bits of the row maximum kmax.
template <class RC>
DYN_DEV void mass_extent(int cand, int kmax, int mid, int G, int& first, int& last)
{
	const unsigned m = __ballot_sync(FULL, cand >= kmax - G);
	const int lane_lo = pmod(mid - RC::HW, RC::SLOTS) / RC::CPL;
	const unsigned r = __funnelshift_r(m, m, lane_lo);  // bit k <=> lane (lane_lo + k) & 31
	first = __ffs(r) - 1;
	last = 31 - __clz(r);
}

// centre column (times 2) of relative lanes first .. last of the window centred at mid
template <class RC>
DYN_DEV int extent_centre2(int mid, int first, int last)
{
	const int lo = mid - RC::HW;
	const int ub = lo - pmod(lo, RC::SLOTS) % RC::CPL;  // unwrapped column of the first slot of the window's lowest lane
	return 2 * ub + (first + last) * RC::CPL + RC::CPL - 1;
}

// The reference forces every cell outside its band (row t: columns band_mid(t) +- bw, NT:96-106) to -inf.  A group whose
// window lies inside the band of all its rows needs nothing; otherwise (the alignment has drifted to the band's edge,
// or the band is narrower than the window) its rows run one by one and every row is clipped to the band exactly.
template <class RC>
DYN_DEV bool group_needs_clip(const RWarp<RC>& w, int mid, int g)
{
	const uint32_t t0 = (uint32_t)RC::GR * (uint32_t)g, t1 = min(t0 + (uint32_t)RC::GR, w.T - 1u);
	// cheap FP32 estimate of the band centre first (t * ratio < 2^17 for any read that fits the device: off by << 1 column);
	// the exact double-precision centres of the reference only when the window is within two columns of the band's edge
	const float c0 = (float)t0 * w.ratio_f;
	const int lo_room = (mid - RC::HW) - ((int)c0 - w.bw_ref), hi_room = ((int)c0 + w.bw_ref) - (mid + RC::HW);
	// (the estimate is within one column of band_mid(t0); the centre moves <= GR/2 + 1 columns over the group's rows: ratio <= 1/2)
	if (lo_room > RC::GR / 2 + 4 && hi_room > 3) return false;
	const int m0 = (int)band_mid(t0, w.ratio), m1 = (int)band_mid(t1, w.ratio);  // band centres are non-decreasing in t
	return (mid + RC::HW > m0 + w.bw_ref) || (mid - RC::HW < m1 - w.bw_ref);
}

// zero the cells of row t that lie outside the reference band
template <class RC>
DYN_DEV void clip_row(const RWarp<RC>& w, float (&p)[RC::CPL], float (&q)[RC::CPL], int mid, uint32_t t)
{
	const int m = (int)band_mid(t, w.ratio);
	const int lo = m - w.bw_ref, hi = m + w.bw_ref;
#pragma unroll
	for (int j = 0; j < RC::CPL; ++j)
	{
		const int col = RWarp<RC>::col_of_slot(w.lane * RC::CPL + j, mid - RC::HW);
		if (col < lo || col > hi)
		{
			p[j] = zero_v<RC>();
			q[j] = zero_v<RC>();
		}
	}
}

// ------------------------------------------------------------------------------------------------------
// backward recurrence (NT_aligner_api.cpp:158-207), linear domain: one row
// ------------------------------------------------------------------------------------------------------
template <class RC>
DYN_DEV void bwd_row(const RWarp<RC>& w, Bw<RC::CPL>& b, float x, float m1, float e2)
{
	constexpr int C = RC::CPL, H = C / 2;
	static_assert(C % 2 == 0, "columns per lane come in pairs (FMUL2 / FFMA2)");
	if constexpr (RC::LOGD)
	{
		// log2 domain (m1, e2 = log2 transition scores): products are sums, the one sum of the recurrence is a log-sum-exp
		float e[C], A[C];
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			e[j] = w.emis_log(j, x);
			A[j] = b.bM[j] + (e[j] + m1);
		}
		const float Araw = __shfl_sync(FULL, A[0], (w.lane + 1) & 31);
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			const float nm = b.bE[j] + e[j];
			const float in = (j + 1 < C) ? A[(j + 1 < C) ? j + 1 : j] : Araw + b.sR;
			b.bE[j] = logplus2(nm + e2, in);
			b.bM[j] = nm;
		}
		return;
	}
	const float2 x2 = make_float2(x, x), m2 = make_float2(m1, m1), e22 = make_float2(e2, e2);
	float2 p[H], A[H];
	// A[n] = bM[t+1][n] * p(t,n) * m1 is consumed by column n-1; pair 0 first: its shuffle has the rest of the row to complete
	p[0] = w.emis_pair(0, x2);
	A[0] = mul2(make_float2(b.bM[0], b.bM[1]), mul2(p[0], m2));
	const float Araw = __shfl_sync(FULL, A[0].x, (w.lane + 1) & 31);
#pragma unroll
	for (int h = 1; h < H; ++h)
	{
		p[h] = w.emis_pair(h, x2);
		A[h] = mul2(make_float2(b.bM[2 * h], b.bM[2 * h + 1]), mul2(p[h], m2));
	}
#pragma unroll
	for (int h = 0; h < H; ++h)
	{
		const float2 nm = mul2(make_float2(b.bE[2 * h], b.bE[2 * h + 1]), p[h]);  // bM[t][n] = bE[t+1][n] * p           (NT:200)
		const float2 in = make_float2(A[h].y, (h + 1 < H) ? A[(h + 1 < H) ? h + 1 : h].x : Araw * b.sR);
		const float2 ne = fma2(nm, e22, in);                                        //                                     (NT:194,201,204)
		b.bM[2 * h] = nm.x;
		b.bM[2 * h + 1] = nm.y;
		b.bE[2 * h] = ne.x;
		b.bE[2 * h + 1] = ne.y;
	}
}

// window one column down (backward direction): the top column retires, the column below the window takes the dead slot
template <class RC>
DYN_DEV void slide_down(RWarp<RC>& w, Bw<RC::CPL>& b, int& mid)
{
	constexpr int C = RC::CPL;
	const int hi = mid + RC::HW;
	if (hi < (int)w.N)
		with_slot<C>(w.lane, pmod(hi, RC::SLOTS), [&](int j) {
			w.set_dead(j);
			zero_slot<C>(b.bM, b.bE, j, zero_v<RC>());
		});
	const int nl = mid - 1 - RC::HW;
	{
		// the dead slot kept the ungated M-transition term of its right neighbour (dp_linear.cuh bwd_step): clear it
		PosConst v;
		v.a = 0.0f; v.b = 0.0f; v.c = CNEG; v.pad = 0.0f;
		if (nl >= 0) v = w.pc[nl];
		with_slot<C>(w.lane, pmod(nl, RC::SLOTS), [&](int j) {
			w.set_col(j, v);
			zero_slot<C>(b.bM, b.bE, j, zero_v<RC>());
		});
	}
	--mid;
}

template <class RC>
DYN_DEV void ckpt_put(float* base, int* ob, uint32_t idx, int lane, const Bw<RC::CPL>& b)
{
	constexpr int C = RC::CPL;
	float* f = base + (size_t)idx * RC::CKF;
#pragma unroll
	for (int j = 0; j < C; ++j)
	{
		f[j * 32 + lane] = b.bM[j];
		f[(C + j) * 32 + lane] = b.bE[j];
	}
	ob[(size_t)idx * 32 + lane] = b.OB;
}

template <class RC>
DYN_DEV void ckpt_store(const SlotScratch& sc, uint32_t g, int lane, const Bw<RC::CPL>& b)
{
	ckpt_put<RC>(sc.ckpt, reinterpret_cast<int*>(sc.ckpt_ob), g, lane, b);
}

// Two-level checkpointing (long reads): pass 1 keeps the checkpoint of every SG-th group only (plus the last group's);
// pass 2 replays the backward pass over one super-group at a time and parks the 8 group checkpoints in a small ring.
constexpr int SG = 8;

// largest value / candidate exponent of this lane and the row maximum, as the group boundary of pass 1 computes them
template <class RC>
DYN_DEV void bwd_stats(const Bw<RC::CPL>& b, int& cand, int& kmax)
{
	float lm = b.bE[0];
#pragma unroll
	for (int j = 1; j < RC::CPL; ++j) lm = fmaxf(lm, b.bE[j]);
	cand = alive<RC>(lm) ? b.OB + vexp<RC>(lm) : NONE;
	kmax = warp_max_int(cand);
}

// log2 domain: after every row a lane takes the integer part of its largest value into its offset (no cross-lane
// traffic except the neighbour's new offset), so the stored values stay in [0, 1) where FP32 resolves 1e-7
template <class RC>
DYN_DEV void lane_renorm_b(const RWarp<RC>& w, Bw<RC::CPL>& b)
{
	float lm = NEG;
#pragma unroll
	for (int j = 0; j < RC::CPL; ++j) lm = lin::max3f(lm, b.bM[j], b.bE[j]);
	if (lm > DEADT)
	{
		const float d = floorf(lm);
#pragma unroll
		for (int j = 0; j < RC::CPL; ++j)
		{
			b.bM[j] -= d;
			b.bE[j] -= d;
		}
		b.OB += (int)d;
	}
	b.sR = (float)(__shfl_sync(FULL, b.OB, (w.lane + 1) & 31) - b.OB);
}

template <class RC>
DYN_DEV void slide_down(RWarp<RC>& w, Bw<RC::CPL>& b, int& mid);

// group boundary of the backward direction: renormalise, move the window down to the next group's centre
template <class RC>
DYN_DEV void bwd_boundary(RWarp<RC>& w, Bw<RC::CPL>& b, int cand, int kmax, int& mid, int target)
{
	constexpr int C = RC::CPL;
	const int nO = max(cand, kmax - RDC);
	const float scl = lin::pow2i(b.OB - nO);
#pragma unroll
	for (int j = 0; j < C; ++j)
	{
		b.bM[j] = rescale<RC>(b.bM[j], b.OB - nO, scl);
		b.bE[j] = rescale<RC>(b.bE[j], b.OB - nO, scl);
	}
	b.OB = nO;
#pragma unroll 1
	while (mid > target) slide_down<RC>(w, b, mid);
	b.sR = nb_factor<RC>(__shfl_sync(FULL, nO, (w.lane + 1) & 31) - nO);
}

// ------------------------------------------------------------------------------------------------------
// pass 1: backward over the whole read; decides and stores the window schedule.  Returns log2 Zb.
// ------------------------------------------------------------------------------------------------------
template <class RC, bool STORE, bool TL>
DYN_DEV double backward_pass(RWarp<RC>& w, const SlotScratch& sc, const BatchArgs& args, int& fault)
{
	constexpr int C = RC::CPL;
	const float m1 = RC::LOGD ? args.m1 : args.m1_lin, e2 = RC::LOGD ? args.e2 : args.e2_lin;
	const int lane = w.lane;
	const int G = args.rib_guard;
	const int S = (int)w.S;
	constexpr int GR = RC::GR;
	const int gl = (S - 1) / GR;  // last group
	Bw<C> b;
	int mid = (int)w.N - 1;  // = the reference's band centre of row T-1
	w.load_window(mid);
#pragma unroll
	for (int j = 0; j < C; ++j) b.bM[j] = b.bE[j] = zero_v<RC>();
	{
		const int q = pmod((int)w.N - 1, RC::SLOTS);
		with_slot<C>(lane, q, SetOne<C>{b.bE, one_v<RC>()});  // bE[T-1][N-1] = 1 (NT:170)
		b.OB = (lane == q / C) ? 0 : -RDC;
		b.sR = nb_factor<RC>(__shfl_sync(FULL, b.OB, (lane + 1) & 31) - b.OB);
	}
	if (STORE)
	{
		ckpt_store<RC>(sc, (uint32_t)(TL ? gl / SG : gl), lane, b);
		if (lane == 0) sc.sched[gl].x = (unsigned)mid;
	}
	float x8;
	{
		const int i = GR * gl + (lane & (GR - 1));
		x8 = (i < S) ? w.sig[i] : 0.0f;
	}
	for (int g = gl; g >= 0; --g)
	{
		const float xg = x8;
		if (g > 0) x8 = w.sig[GR * (g - 1) + (lane & (GR - 1))];
		const bool clip = group_needs_clip<RC>(w, mid, g);
		__syncwarp();  // (the compiler then knows the warp is converged: plain SHFL instead of WARPSYNC + SHFL + ENDCOLLECTIVE per row)
		if (!RC::LOGD && g < gl && !clip)
		{
#pragma unroll
			for (int k = GR - 1; k >= 0; --k) bwd_row<RC>(w, b, __shfl_sync(FULL, xg, k), m1, e2);
		}
		else
		{
#pragma unroll 1
			for (int k = ((g < gl) ? GR : S - GR * gl) - 1; k >= 0; --k)
			{
				bwd_row<RC>(w, b, __shfl_sync(FULL, xg, k), m1, e2);
				if (clip) clip_row<RC>(w, b.bM, b.bE, mid, (uint32_t)GR * (uint32_t)g + (uint32_t)k);
				if (RC::LOGD) lane_renorm_b<RC>(w, b);
			}
		}
		if (g == 0) break;
		// ---- group boundary: row 8g.  Renormalise, decide the window of group g-1, move there, checkpoint -----
		int cand, kmax;
		bwd_stats<RC>(b, cand, kmax);
		// (inside a group clipped to the reference band that is reason 3: the alignment has left the band — the forward and
		// backward ridges then separate by up to the band's width, which no window holds: full-band kernels)
		if (kmax == NONE) { fault = clip ? 3 : 1; RIB_DBG("p1 g=%d nothing alive (clip %d)\n", g, (int)clip); return NAN; }
		int first, last;
		mass_extent<RC>(cand, kmax, mid, G, first, last);
		if ((first == 0 || last == 31) && !fault) { fault = 2; RIB_DBG("p1 g=%d mid=%d edge first=%d last=%d\n", g, mid, first, last); }
		int s = (2 * mid - extent_centre2<RC>(mid, first, last)) / 2;  // window centre above the mass centre: move down
		s = max(0, min(s, min(GR, mid)));
		bwd_boundary<RC>(w, b, cand, kmax, mid, mid - s);
		if (fault) return NAN;
		if (STORE)
		{
			if (!TL) ckpt_store<RC>(sc, (uint32_t)(g - 1), lane, b);
			else if ((g - 1) % SG == SG - 1) ckpt_store<RC>(sc, (uint32_t)((g - 1) / SG), lane, b);
			if (lane == 0) sc.sched[g - 1].x = (unsigned)mid;
		}
	}
	// column 0 must be inside the window of row 0
	if (mid > RC::HW && !fault) { fault = 4; RIB_DBG("p1 end mid=%d\n", mid); }
	if (fault) return NAN;
	// Zb = bE[0][0] (NT:286): column 0 is ring slot 0 = lane 0, j 0
	const double z = (RC::LOGD ? (double)b.bE[0] : log2((double)b.bE[0])) + (double)b.OB;
	return shfl_f64(z, 0);
}

// ------------------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------------------
// write the training statistics of ring slot j of this lane for lattice column col and clear them
template <class RC>
DYN_DEV void flush_slot(const BatchArgs& args, uint64_t pc_off, TrainAcc<RC::CPL>& a, int j, int col)
{
	constexpr int C = RC::CPL;
#pragma unroll
	for (int jj = 0; jj < C; ++jj)
		if (jj == j)
		{
			if (col >= 1 && a.gw[jj] > 0.0f)  // column 0 scores no kmer (its slot only ever sees the posterior of cell (0,0))
			{
				// un-centre in double: sum g*x = gx + mu*gw, sum g*x^2 = gxx + 2*mu*gx + mu^2*gw
				const double mu = (double)a.mu[jj], gw = (double)a.gw[jj], gx = (double)a.gx[jj], gxx = (double)a.gxx[jj];
				args.read_w[pc_off + col] = gw;
				args.read_x[pc_off + col] = gx + mu * gw;
				args.read_xx[pc_off + col] = gxx + 2.0 * mu * gx + mu * mu * gw;
			}
			a.gw[jj] = 0.0f;
			a.gx[jj] = 0.0f;
			a.gxx[jj] = 0.0f;
		}
}

// window one column up (forward direction): the lowest column retires, the column above the window takes the dead slot
template <class RC, int MODE>
DYN_DEV void slide_up(RWarp<RC>& w, Fw<RC::CPL>& f, TrainAcc<RC::CPL>& ta, const BatchArgs& args, uint64_t pc_off, int& mid)
{
	constexpr int C = RC::CPL;
	const int lo = mid - RC::HW;
	if (lo >= 0)
		with_slot<C>(w.lane, pmod(lo, RC::SLOTS), [&](int j) {
			if (MODE == 2) flush_slot<RC>(args, pc_off, ta, j, lo);
			w.set_dead(j);
			zero_slot<C>(f.fM, f.fE, j, zero_v<RC>());
			zero_slot<C>(f.VM, f.VE, j, zero_v<RC>());
		});
	const int nh = mid + 1 + RC::HW;
	if (nh < (int)w.N)
	{
		const PosConst v = w.pc[nh];
		with_slot<C>(w.lane, pmod(nh, RC::SLOTS), [&](int j) {
			w.set_col(j, v);
			zero_slot<C>(f.fM, f.fE, j, zero_v<RC>());
			zero_slot<C>(f.VM, f.VE, j, zero_v<RC>());
			if (MODE == 2)
			{
#pragma unroll
				for (int jj = 0; jj < C; ++jj)
					if (jj == j) ta.mu[jj] = v.pad;
			}
		});
	}
	++mid;
}

// One row of pass 2.  On entry f holds the forward values of row t and the Viterbi values of row t-1; bc / bn: this lane's
// backward values of rows t / t+1, already multiplied by the posterior factor of the group.
//   posteriors of row t, then MODE 1: posterior-Viterbi update (NT:357-362), decision bits, sparse posterior records;
//                             MODE 2: training statistics (xprev = x[t-1], the sample the posteriors of row t weigh, NT:509-512)
//   STEP: forward recurrence to row t+1 (NT:141-150)
// hdr_row / recs: where row t's header / the records go.  Returns the posterior mass of this lane's cells.
//   MODE 3: as MODE 1 without records (the row header holds the C decision words only); MODE 4: the gather sweep that
//   follows the traceback of MODE 3 — `cell` = path cell of row t (column | match state << 31), whose posterior goes to
//   hdr_row[0] (= pp[t]); no posterior-Viterbi state.
// The same row in the log2 domain (MODE 3: fill, decision words only; MODE 4: gather).  f, bc, bn hold log2 values against
// their lane offsets; kc / kn = (OF + OB(row t / t+1) - floor(Z2)) + log2 c0 of this lane, so that fE + bc + kc and
// fM + bn + e + kn ARE the log2 posteriors.  m1, e2: log2 transition scores.
template <class RC, int MODE, bool STEP>
DYN_DEV float fwd_row_log(const RWarp<RC>& w, Fw<RC::CPL>& f, TrainAcc<RC::CPL>& ta, uint32_t* hdr_row, float x, float xprev,
	const float (&bc)[RC::CPL], const float (&bn)[RC::CPL], float kc, float kn, float m1, float e2, uint32_t cell)
{
	constexpr int C = RC::CPL;
	static_assert(MODE == 2 || MODE == 3 || MODE == 4, "the log2-domain ribbon: training statistics, or alignment with the records-free layout");
	const int lane = w.lane;
	const float vlraw = (MODE == 3) ? __shfl_sync(FULL, f.VE[C - 1], (lane + 31) & 31) : NEG;
	const float flraw = STEP ? __shfl_sync(FULL, f.fE[C - 1], (lane + 31) & 31) : NEG;
	float e[C], PM[C], PE[C];
	float msum = 0.0f;
#pragma unroll
	for (int j = 0; j < C; ++j)
	{
		e[j] = STEP ? w.emis_log(j, x) : 0.0f;
		PE[j] = f.fE[j] + (bc[j] + kc);
		PM[j] = STEP ? f.fM[j] + ((bn[j] + e[j]) + kn) : NEG;  // bM[t][n] = bE[t+1][n] * p(t,n) (NT:200); no match state in the last row
		const float gm = ex2(PM[j]), ge = ex2(PE[j]);
		msum += gm + ge;
		if (MODE == 2)
		{
			// training statistics (NT:494-514), as in the linear domain: gamma = pM + pE weighs sample x[t-1]
			const float g = gm + ge;
			const float dx = xprev - ta.mu[j];
			const float gd = g * dx;
			ta.gw[j] += g;
			ta.gx[j] += gd;
			ta.gxx[j] = fmaf(gd, dx, ta.gxx[j]);
			ta.sM += gm;
			ta.sE += ge;
		}
	}
	if (MODE == 3)
	{
		// posterior-Viterbi fill (NT:357-362) as a max-SUM of log2 posteriors; the decision bits as in the linear domain
		const float vl = vlraw + f.sV;
		unsigned bits[C];
		float vmx[C], left[C];
#pragma unroll
		for (int j = C - 1; j >= 0; --j)
		{
			vmx[j] = fmaxf(f.VM[j], f.VE[j]);
			bits[j] = __ballot_sync(FULL, f.VM[j] < f.VE[j]);
			left[j] = (j > 0) ? f.VE[(j > 0) ? j - 1 : 0] : vl;
		}
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			f.VM[j] = fmaxf(left[j] + PM[j], NEG);
			f.VE[j] = fmaxf(vmx[j] + PE[j], NEG);
		}
		if (lane == 0)
		{
			if (C == 2) *reinterpret_cast<uint2*>(hdr_row) = make_uint2(bits[0], bits[C - 1]);
			else
			{
#pragma unroll
				for (int j = 0; j < C; ++j) hdr_row[j] = bits[j];
			}
		}
	}
	else if (MODE == 4)
	{
		const int q = (int)((cell & 0x7fffffffu) % (uint32_t)RC::SLOTS);
		const int ql = q / C, jq = q - ql * C;
		const bool isM = (cell >> 31) != 0u;
		float pv = NEG;
#pragma unroll
		for (int j = 0; j < C; ++j)
			if (j == jq) pv = isM ? PM[j] : PE[j];
		if (lane == ql && cell != 0xffffffffu) *reinterpret_cast<float*>(hdr_row) = ex2(pv);
	}
	if (STEP)
	{
		const float fl = flraw + f.sL;
		float left[C];
#pragma unroll
		for (int j = 0; j < C; ++j) left[j] = (j > 0) ? f.fE[(j > 0) ? j - 1 : 0] : fl;
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			const float ne = logplus2(f.fE[j] + e2, f.fM[j]) + e[j];  // (fM + fE*e2) * p    (NT:146-150, e1 = 1)
			f.fM[j] = fmaxf(left[j] + (e[j] + m1), NEG);              // fE[t][n-1] * p * m1  (NT:143)
			f.fE[j] = fmaxf(ne, NEG);
		}
	}
	return msum;
}

// the forward counterpart of lane_renorm_b
template <class RC>
DYN_DEV void lane_renorm_f(const RWarp<RC>& w, Fw<RC::CPL>& f)
{
	float lm = NEG;
#pragma unroll
	for (int j = 0; j < RC::CPL; ++j) lm = lin::max3f(lm, f.fM[j], f.fE[j]);
	if (lm > DEADT)
	{
		const float d = floorf(lm);
#pragma unroll
		for (int j = 0; j < RC::CPL; ++j)
		{
			f.fM[j] -= d;
			f.fE[j] -= d;
		}
		f.OF += (int)d;
	}
	f.sL = (float)(__shfl_sync(FULL, f.OF, (w.lane + 31) & 31) - f.OF);
}

template <class RC, int MODE, bool STEP, bool MASS = true>
DYN_DEV float fwd_row(const RWarp<RC>& w, Fw<RC::CPL>& f, RowSink& rs, TrainAcc<RC::CPL>& ta, uint32_t* hdr_row, float* recs,
	float thr, float x, float xprev, const float (&bc)[RC::CPL], const float (&bn)[RC::CPL], float m1, float e2, uint32_t cell = 0u)
{
	constexpr int C = RC::CPL;
	constexpr bool VIT = (MODE == 1 || MODE == 3);
	const int lane = w.lane;
	// the values the right lane needs are those of the previous row: send them first, consume them last
	constexpr int H = C / 2;
	const float vlraw = VIT ? __shfl_sync(FULL, f.VE[C - 1], (lane + 31) & 31) : 0.0f;
	const float flraw = STEP ? __shfl_sync(FULL, f.fE[C - 1], (lane + 31) & 31) : 0.0f;
	const float2 x2 = make_float2(x, x);
	float2 p[H];
	float PM[C], PE[C];
	float msum = 0.0f;
#pragma unroll
	for (int h = 0; h < H; ++h)
	{
		p[h] = STEP ? w.emis_pair(h, x2) : make_float2(0.0f, 0.0f);
		const float2 pe = mul2(make_float2(f.fE[2 * h], f.fE[2 * h + 1]), make_float2(bc[2 * h], bc[2 * h + 1]));
		// bM[t][n] = bE[t+1][n] * p(t,n) (NT:200); the last row has no match state
		const float2 pm = STEP ? mul2(make_float2(f.fM[2 * h], f.fM[2 * h + 1]), mul2(make_float2(bn[2 * h], bn[2 * h + 1]), p[h]))
		                       : make_float2(0.0f, 0.0f);
		PE[2 * h] = pe.x;
		PE[2 * h + 1] = pe.y;
		PM[2 * h] = pm.x;
		PM[2 * h + 1] = pm.y;
		if (MASS)
		{
			msum += pm.x + pe.x;
			msum += pm.y + pe.y;
		}
	}
	if (VIT)
	{
		// posterior-Viterbi fill (NT:357-362) as a max-product, in place from the highest slot down;
		// decision bit set <=> the E state of this cell is entered from E (the test of NT:448 at fill time)
		const float vl = vlraw * f.sV;
		unsigned bits[C];
		float lmax = 0.0f;
		float vmx[C], left[C];
#pragma unroll
		for (int j = C - 1; j >= 0; --j)
		{
			vmx[j] = fmaxf(f.VM[j], f.VE[j]);
			bits[j] = __ballot_sync(FULL, f.VM[j] < f.VE[j]);
			left[j] = (j > 0) ? f.VE[j - 1] : vl;
			if (MODE == 1) lmax = lin::max3f(lmax, PM[j], PE[j]);
		}
#pragma unroll
		for (int h = 0; h < H; ++h)
		{
			const float2 vm = mul2(make_float2(left[2 * h], left[2 * h + 1]), make_float2(PM[2 * h], PM[2 * h + 1]));
			const float2 ve = mul2(make_float2(vmx[2 * h], vmx[2 * h + 1]), make_float2(PE[2 * h], PE[2 * h + 1]));
			f.VM[2 * h] = vm.x;
			f.VM[2 * h + 1] = vm.y;
			f.VE[2 * h] = ve.x;
			f.VE[2 * h + 1] = ve.y;
		}
		if (MODE == 3)
		{
			if (lane == 0)
			{
				if (C == 2) *reinterpret_cast<uint2*>(hdr_row) = make_uint2(bits[0], bits[C - 1]);
				else
				{
#pragma unroll
					for (int j = 0; j < C; ++j) hdr_row[j] = bits[j];
				}
			}
		}
		else
		{
		// sparse posterior records: one per lane that holds a posterior above the threshold (NaN counts as hot)
		const bool hot = !(lmax <= thr);
		const unsigned hm = __ballot_sync(FULL, hot);
		const uint32_t pos = rs.n + __popc(hm & ((1u << lane) - 1u));
		if (hot)  // room for a whole group of rows is checked once per group (forward_pass)
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
			float4* dst = reinterpret_cast<float4*>(recs + (size_t)pos * RC::RECF);
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
			uint4* dst = reinterpret_cast<uint4*>(hdr_row);
#pragma unroll
			for (int q = 0; q < RC::HDRW / 4; ++q) dst[q] = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
		}
		rs.n += __popc(hm);
		}
	}
	else if (MODE == 4)
	{
		// the posterior of the path cell: the lane that holds its ring slot writes it
		const int q = (int)((cell & 0x7fffffffu) % (uint32_t)RC::SLOTS);
		const int ql = q / C, jq = q - ql * C;
		const bool isM = (cell >> 31) != 0u;
		float pv = 0.0f;
#pragma unroll
		for (int j = 0; j < C; ++j)
			if (j == jq) pv = isM ? PM[j] : PE[j];
		if (lane == ql && cell != 0xffffffffu) *reinterpret_cast<float*>(hdr_row) = pv;
	}
	else
	{
		// training statistics (NT:494-514): gamma = pM + pE weighs sample x[t-1] for the kmer of the cell's column
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			const float g = PM[j] + PE[j];
			const float dx = xprev - ta.mu[j];
			const float gd = g * dx;
			ta.gw[j] += g;
			ta.gx[j] += gd;
			ta.gxx[j] = fmaf(gd, dx, ta.gxx[j]);
			ta.sM += PM[j];
			ta.sE += PE[j];
		}
	}
	if (STEP)
	{
		const float fl = flraw * f.sL;
		const float2 m2 = make_float2(m1, m1), e22 = make_float2(e2, e2);
		float left[C];
#pragma unroll
		for (int j = 0; j < C; ++j) left[j] = (j > 0) ? f.fE[j - 1] : fl;
#pragma unroll
		for (int h = 0; h < H; ++h)
		{
			const float2 fe = make_float2(f.fE[2 * h], f.fE[2 * h + 1]);
			const float2 ne = mul2(fma2(fe, e22, make_float2(f.fM[2 * h], f.fM[2 * h + 1])), p[h]);  // (fM + fE*e2) * p    (NT:146-150, e1 = 1)
			const float2 nm = mul2(make_float2(left[2 * h], left[2 * h + 1]), mul2(p[h], m2));       // fE[t][n-1] * p * m1  (NT:143)
			f.fM[2 * h] = nm.x;
			f.fM[2 * h + 1] = nm.y;
			f.fE[2 * h] = ne.x;
			f.fE[2 * h + 1] = ne.y;
		}
	}
	return msum;
}

// pass 2: recomputation + forward + posterior (+ posterior-Viterbi fill | training statistics), group by group.
// Returns log2 Zf - log2 Zb (NaN on a fault).
template <class RC, int MODE, bool TL>
DYN_DEV double forward_pass(RWarp<RC>& w, const SlotScratch& sc, const BatchArgs& args, uint64_t pc_off,
	unsigned char* smem_raw, double Z2, uint32_t& nrec_out, bool& overflow, double& xi_m, double& xi_e, int& fault_out)
{
	constexpr int C = RC::CPL;
	constexpr int ROWF = RC::ROWF;
	constexpr int HW = RC::HW;
	const float m1 = RC::LOGD ? args.m1 : args.m1_lin, e2 = RC::LOGD ? args.e2 : args.e2_lin;
	const float thr = args.thr_rib;
	const int G = args.rib_guard;
	const int lane = w.lane;
	const int S = (int)w.S;
	constexpr int GR = RC::GR;
	const int gl = (S - 1) / GR;
	// row k, slot pair h (slots 2h, 2h+1) of this lane: rows2[(k * H2 + h) * 32] — one 64-bit shared-memory access per pair
	constexpr int H2 = C / 2;
	float2* const rows2 = reinterpret_cast<float2*>(smem_raw) + lane;
	// log2 domain: the rows are stored as they are, next to the lane offset in force for each (it changes row by row)
	int* const rows_ob = reinterpret_cast<int*>(smem_raw + (size_t)(RC::GR + 1) * ROWF * 4) + lane;
	auto put_row = [&](int k, const Bw<C>& bb, float kp) {
		if (RC::LOGD)
		{
#pragma unroll
			for (int h = 0; h < H2; ++h) rows2[(k * H2 + h) * 32] = make_float2(bb.bE[2 * h], bb.bE[2 * h + 1]);
			rows_ob[k * 32] = bb.OB;
			return;
		}
		const float2 k2 = make_float2(kp, kp);
#pragma unroll
		for (int h = 0; h < H2; ++h) rows2[(k * H2 + h) * 32] = mul2(make_float2(bb.bE[2 * h], bb.bE[2 * h + 1]), k2);
	};
	auto get_row = [&](int k, float (&dst)[C]) {
#pragma unroll
		for (int h = 0; h < H2; ++h)
		{
			const float2 v = rows2[(k * H2 + h) * 32];
			dst[2 * h] = v.x;
			dst[2 * h + 1] = v.y;
		}
	};
	float* const recs = static_cast<float*>(sc.recs);
	Fw<C> f;
	Bw<C> b;
	TrainAcc<C> ta;
	RowSink rs;
	rs.n = 0;
	rs.cap = (uint32_t)args.rec_cap;
	rs.overflow = false;
	// posterior = sf * sb * 2^(OF + OB - Z2) = sf * sb * c0 * 2^(OF + OB - Z2i); c0 absorbs the closed-loop correction
	const double Z2f = floor(Z2);
	const int Z2i = (int)Z2f;
	float c0 = (float)exp2(Z2f - Z2);
	int fault = 0;

	int mid = (int)sc.sched[0].x;
	w.load_window(mid);
#pragma unroll
	for (int j = 0; j < C; ++j)
	{
		f.fM[j] = f.fE[j] = f.VM[j] = f.VE[j] = zero_v<RC>();
		ta.gw[j] = ta.gx[j] = ta.gxx[j] = 0.0f;
		ta.mu[j] = 0.0f;
	}
	ta.sM = ta.sE = 0.0f;
	ta.dM = 0.0;
	ta.dE = (lane == 0) ? -1.0 : 0.0;  // the uniform row body also counts row 0: posterior(0,0) = 1, not part of NT:494-514
	if (MODE == 2)
	{
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			const int n = RWarp<RC>::col_of_slot(lane * C + j, mid - HW);
			ta.mu[j] = (n >= 0 && n <= min(mid + HW, (int)w.N - 1)) ? w.pc[n].pad : 0.0f;
		}
	}
	// row 0: fE[0][0] = 1 (NT:120).  The posterior-Viterbi state starts one row early, VE[-1][0] = 1, so that the uniform
	// row body yields VE[0][0] = 1 * posterior(0,0) = 1 (NT:336) at row 0
	f.OF = (lane == 0) ? 0 : -RDC;
	f.OV = (lane == 0) ? 0 : -RDCV;
	if (lane == 0)
	{
		f.fE[0] = one_v<RC>();
		f.VE[0] = one_v<RC>();
	}
	f.sL = nb_factor<RC>(__shfl_sync(FULL, f.OF, (lane + 31) & 31) - f.OF);
	f.sV = nb_factor<RC>(__shfl_sync(FULL, f.OV, (lane + 31) & 31) - f.OV);

	// checkpoint / samples / next window centre of the group ahead are requested one group early
	float ckM[C], ckE[C];
	int ckO;
	float x8;
	uint32_t cg8 = 0xffffffffu;  // MODE 4: path cells of rows 8g .. 8g+8 (lanes 0 .. 8)
	int mid_next;
	constexpr int HDRS = (MODE == 3) ? C : RC::HDRW;  // words per row header
	int* const ring_ob = reinterpret_cast<int*>(sc.ring + (size_t)SG * RC::CKF);
	// two-level checkpoints: replay the backward pass over the super-group that starts at group g0 and park the group
	// checkpoints in the ring (the emission window travels up to the super-group's last group and back down to g0's)
	auto replay = [&](int g0) {
		const int gs = min(g0 + SG - 1, gl);
		const float* cf = sc.ckpt + (size_t)(gs / SG) * RC::CKF;
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			b.bM[j] = cf[j * 32 + lane];
			b.bE[j] = cf[(C + j) * 32 + lane];
		}
		b.OB = reinterpret_cast<const int*>(sc.ckpt_ob)[(size_t)(gs / SG) * 32 + lane];
		b.sR = nb_factor<RC>(__shfl_sync(FULL, b.OB, (lane + 1) & 31) - b.OB);
		int mr = (int)sc.sched[gs].x;
		if (gs != g0) w.load_window(mr);
		ckpt_put<RC>(sc.ring, ring_ob, (uint32_t)(gs % SG), lane, b);
#pragma unroll 1
		for (int gp = gs; gp > g0; --gp)
		{
			const int i = GR * gp + (lane & (GR - 1));
			const float xr = (i < S) ? w.sig[i] : 0.0f;
			const int target = (int)sc.sched[gp - 1].x;
			const bool clipr = group_needs_clip<RC>(w, mr, gp);
			if (!RC::LOGD && gp < gl && !clipr)
			{
#pragma unroll
				for (int k = GR - 1; k >= 0; --k) bwd_row<RC>(w, b, __shfl_sync(FULL, xr, k), m1, e2);
			}
			else
			{
#pragma unroll 1
				for (int k = ((gp < gl) ? GR : S - GR * gl) - 1; k >= 0; --k)
				{
					bwd_row<RC>(w, b, __shfl_sync(FULL, xr, k), m1, e2);
					if (clipr) clip_row<RC>(w, b.bM, b.bE, mr, (uint32_t)GR * (uint32_t)gp + (uint32_t)k);
					if (RC::LOGD) lane_renorm_b<RC>(w, b);
				}
			}
			int cand, kmax;
			bwd_stats<RC>(b, cand, kmax);
			bwd_boundary<RC>(w, b, cand, kmax, mr, target);
			ckpt_put<RC>(sc.ring, ring_ob, (uint32_t)((gp - 1) % SG), lane, b);
		}
		__syncwarp();
	};
	auto prefetch = [&](int g) {
		const float* cf = TL ? sc.ring + (size_t)(g % SG) * RC::CKF : sc.ckpt + (size_t)g * RC::CKF;
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			ckM[j] = cf[j * 32 + lane];
			ckE[j] = cf[(C + j) * 32 + lane];
		}
		ckO = TL ? ring_ob[(size_t)(g % SG) * 32 + lane] : reinterpret_cast<const int*>(sc.ckpt_ob)[(size_t)g * 32 + lane];
		const int i = GR * g + (lane & (GR - 1));
		x8 = (i < S) ? w.sig[i] : 0.0f;
		if (MODE == 4)
		{
			// (read before this group's posteriors overwrite the same words: the stores depend on these values)
			const int r = GR * g + lane;
			cg8 = (lane <= GR && r <= S) ? __float_as_uint(sc.pp[r]) : 0xffffffffu;
		}
		mid_next = (g < gl) ? (int)sc.sched[g + 1].x : 0;
	};
	if (TL) replay(0);
	prefetch(0);
	float xprev = 0.0f;
	float macc = 0.0f;

	for (int g = 0; g <= gl; ++g)
	{
		const int nr = (g < gl) ? GR : S - GR * gl;  // samples of this group
		if (MODE == 1 && rs.n + (uint32_t)(GR + 1) * 32u > rs.cap)
		{
			// not enough room for the records this group can produce at most (pathological record density)
			rs.overflow = true;
			fault = 10;
			break;
		}
		const float xg = x8;
		const uint32_t cgq = cg8;
		const int midn = mid_next;
		// ---- step a: recompute the backward rows 8g+nr .. 8g of this group into shared memory, scaled by the group's
		// posterior factor kap = 2^(OF + OB - Z2) (forward offsets are fixed inside a group)
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			b.bM[j] = ckM[j];
			b.bE[j] = ckE[j];
		}
		b.OB = ckO;
		b.sR = nb_factor<RC>(__shfl_sync(FULL, b.OB, (lane + 1) & 31) - b.OB);
		const float kap = RC::LOGD ? 0.0f : lin::kappa(f.OF, b.OB, Z2i, c0);
		const float c0l = RC::LOGD ? lg2(c0) : 0.0f;  // log2 domain: the closed-loop correction as a summand
		const bool clip = group_needs_clip<RC>(w, mid, g);
		__syncwarp();  // (as in pass 1)
		if (!RC::LOGD && g < gl && !clip)
		{
			put_row(GR, b, kap);
#pragma unroll
			for (int k = GR - 1; k >= 0; --k)
			{
				bwd_row<RC>(w, b, __shfl_sync(FULL, xg, k), m1, e2);
				put_row(k, b, kap);
			}
		}
		else
		{
			put_row(nr, b, kap);
#pragma unroll 1
			for (int k = nr - 1; k >= 0; --k)
			{
				bwd_row<RC>(w, b, __shfl_sync(FULL, xg, k), m1, e2);
				if (clip) clip_row<RC>(w, b.bM, b.bE, mid, (uint32_t)GR * (uint32_t)g + (uint32_t)k);
				if (RC::LOGD) lane_renorm_b<RC>(w, b);
				put_row(k, b, kap);
			}
		}
		__syncwarp();
		if (g < gl && !(TL && (g + 1) % SG == 0)) prefetch(g + 1);

		// ---- step b: forward rows 8g .. 8g+nr-1 -------------------------------------------------------------
		// where row 8g+k's header goes (MODE 4: its path posterior)
		uint32_t* const hdr_g = (MODE == 4) ? reinterpret_cast<uint32_t*>(sc.pp) + (size_t)(GR * g) : sc.hdr + (size_t)(GR * g) * HDRS;
		constexpr int HSTR = (MODE == 4) ? 1 : HDRS;
		float bc[C], bn[C];
		get_row(0, bc);
		int obc = RC::LOGD ? rows_ob[0] : 0, obn = 0;
		if (!RC::LOGD && g < gl && !clip)
		{
#pragma unroll
			for (int k = 0; k < GR; ++k)
			{
				get_row(k + 1, bn);
				const float x = __shfl_sync(FULL, xg, k);
				// the posterior mass is measured on the group's FIRST and LAST row: offsets and the posterior factor are fixed
				// inside a group and the drift of the stored values is monotone in the row (backward values shrink towards the
				// group's first row, forward values towards its last), so a lane that lost its values, or whose factor left
				// the float range (the clamp of lin::kappa), shows at one of the two ends.  (With 8-row groups the last row was
				// enough; at 16 rows a 2.5x-noise read lost 2^-128 over one group on the backward side only, and one of 24 such
				// reads came back with a posterior of 2^-10 instead of 1 and no fault — gpu_soak, third session.)
				const uint32_t cell = (MODE == 4) ? __shfl_sync(FULL, cgq, k) : 0u;
				if (k == GR - 1 || k == 0) macc += fwd_row<RC, MODE, true, true>(w, f, rs, ta, hdr_g + k * HSTR, recs, thr, x, xprev, bc, bn, m1, e2, cell);
				else fwd_row<RC, MODE, true, false>(w, f, rs, ta, hdr_g + k * HSTR, recs, thr, x, xprev, bc, bn, m1, e2, cell);
				xprev = x;
#pragma unroll
				for (int j = 0; j < C; ++j) bc[j] = bn[j];
			}
		}
		else
		{
#pragma unroll 1
			for (int k = 0; k < nr; ++k)
			{
				get_row(k + 1, bn);
				const float x = __shfl_sync(FULL, xg, k);
				const uint32_t cell = (MODE == 4) ? __shfl_sync(FULL, cgq, k) : 0u;
				if constexpr (RC::LOGD)
				{
					obn = rows_ob[(k + 1) * 32];
					const float kc = (float)(f.OF + obc - Z2i) + c0l, kn = (float)(f.OF + obn - Z2i) + c0l;
					macc += fwd_row_log<RC, MODE, true>(w, f, ta, hdr_g + k * HSTR, x, xprev, bc, bn, kc, kn, m1, e2, cell);
					obc = obn;
				}
				else macc += fwd_row<RC, MODE, true>(w, f, rs, ta, hdr_g + k * HSTR, recs, thr, x, xprev, bc, bn, m1, e2, cell);
				if (clip) clip_row<RC>(w, f.fM, f.fE, mid, (uint32_t)GR * (uint32_t)g + (uint32_t)k + 1u);  // f now holds row t+1
				if constexpr (RC::LOGD) lane_renorm_f<RC>(w, f);
				xprev = x;
#pragma unroll
				for (int j = 0; j < C; ++j) bc[j] = bn[j];
			}
			// last row T-1: posteriors, Viterbi, bits, records; no forward step, no match posterior
			if (g == gl)
			{
				const uint32_t cell = (MODE == 4) ? __shfl_sync(FULL, cgq, nr) : 0u;
				if constexpr (RC::LOGD)
				{
					const float kc = (float)(f.OF + obc - Z2i) + c0l;
					macc += fwd_row_log<RC, MODE, false>(w, f, ta, hdr_g + nr * HSTR, 0.0f, xprev, bc, bn, kc, kc, m1, e2, cell);
				}
				else macc += fwd_row<RC, MODE, false>(w, f, rs, ta, hdr_g + nr * HSTR, recs, thr, 0.0f, xprev, bc, bn, m1, e2, cell);
			}
		}
		__syncwarp();

		// ---- group boundary: state row 8g+8 ------------------------------------------------------------------
		// closed loop: the posterior mass of every row is 1; a row that lost mass is a fault, the mean deviation (slow
		// common-mode FP32 drift) is folded into the posterior factor of the next group
		const float cnt = (!RC::LOGD && g < gl && !clip) ? 2.0f : ((g < gl) ? (float)nr : (float)(nr + 1));  // rows whose mass was summed
		const float mass = warp_sum(macc, lane);
		macc = 0.0f;
		if (!(fabsf(mass - cnt) <= RIB_MASS_TOL) && !fault) { fault = 5; RIB_DBG("p2 g=%d mass=%g of %g\n", g, mass, cnt); }
		if (MODE == 2)
		{
			ta.dM += (double)ta.sM;
			ta.dE += (double)ta.sE;
			ta.sM = ta.sE = 0.0f;
		}
		if (g == gl) break;
#ifndef DYN_HOST_EMU
		c0 *= __fdividef(cnt, mass);  // (mass is 1 +- 1e-4: the fast reciprocal is exact to 2 ulp, and both sweeps use it)
#else
		c0 *= cnt / mass;
#endif
		// forward values: renormalise (own maximum -> [1, 2), at most RDC below the row's largest lane), window guard,
		// range guard (largest forward x largest backward value of the row against Z)
		{
			float lm = zero_v<RC>(), bm = zero_v<RC>();
#pragma unroll
			for (int j = 0; j < C; ++j)
			{
				lm = lin::max3f(lm, f.fM[j], f.fE[j]);
				bm = fmaxf(bm, bc[j]);  // row 8g+8, times kap = b / Z * 2^OF
			}
			const int cand = alive<RC>(lm) ? f.OF + vexp<RC>(lm) : NONE;
			const int kmax = warp_max_int(cand);
			// (the range guard is about what a flushed cell can have carried: nothing is flushed in the log2 domain)
			const int kb = RC::LOGD ? 0 : warp_max_int(is_alive(bm) ? fexp(bm) - f.OF : NONE);
			if ((kmax == NONE || kb == NONE || (!RC::LOGD && kmax + kb > lin::LIN_GUARD_BITS)) && !fault) { fault = 6; RIB_DBG("p2 g=%d range kmax=%d kb=%d\n", g, kmax, kb); }
			int first, last;
			mass_extent<RC>(cand, kmax, mid, G, first, last);
			if ((first == 0 || last == 31) && !fault) { fault = 7; RIB_DBG("p2 g=%d mid=%d edge first=%d last=%d\n", g, mid, first, last); }
			const int nO = max(cand, kmax - RDC);
			const float scl = lin::pow2i(f.OF - nO);
#pragma unroll
			for (int j = 0; j < C; ++j)
			{
				f.fM[j] = rescale<RC>(f.fM[j], f.OF - nO, scl);
				f.fE[j] = rescale<RC>(f.fE[j], f.OF - nO, scl);
			}
			f.OF = nO;
			f.sL = nb_factor<RC>(__shfl_sync(FULL, nO, (lane + 31) & 31) - nO);
		}
		if (MODE == 1 || MODE == 3)
		{
			float lm = zero_v<RC>();
#pragma unroll
			for (int j = 0; j < C; ++j) lm = lin::max3f(lm, f.VM[j], f.VE[j]);
			const int cand = alive<RC>(lm) ? f.OV + vexp<RC>(lm) - lin::E0V : NONE;
			const int kmax = warp_max_int(cand);
			// every posterior-Viterbi score underflowed (or is NaN): the decision bits from here on would be meaningless
			if (kmax == NONE && !fault) { fault = 8; RIB_DBG("p2 g=%d viterbi dead\n", g); }
			const int nO = max(cand, kmax - RDCV);
			const float scl = lin::pow2i(f.OV - nO);
#pragma unroll
			for (int j = 0; j < C; ++j)
			{
				f.VM[j] = rescale<RC>(f.VM[j], f.OV - nO, scl);
				f.VE[j] = rescale<RC>(f.VE[j], f.OV - nO, scl);
			}
			f.OV = nO;
			f.sV = nb_factor<RC>(__shfl_sync(FULL, nO, (lane + 31) & 31) - nO);
		}
		if (fault) break;
		// move to the window of the next group
#pragma unroll 1
		while (mid < midn) slide_up<RC, MODE>(w, f, ta, args, pc_off, mid);
		if (TL && (g + 1) % SG == 0)
		{
			// next super-group: its checkpoints first (the replay leaves the emission window where it is now)
			replay(g + 1);
			prefetch(g + 1);
		}
	}
	// Zf = fE[T-1][N-1] (NT:285)
	float v = zero_v<RC>();
	with_slot<C>(lane, pmod((int)w.N - 1, RC::SLOTS), GetOne<C>{f.fE, v});
	const int ql = pmod((int)w.N - 1, RC::SLOTS) / C;
	double dz = (RC::LOGD ? (double)v : log2((double)v)) + (double)f.OF - Z2;
	dz = shfl_f64(dz, ql);
	RIB_DBG("p2 end dz=%g fault=%d\n", dz, fault);
	fault = warp_max_int(fault);
	fault_out = fault;
	if (fault) dz = NAN;
	if (MODE == 2)
	{
		// columns still inside the window
#pragma unroll
		for (int j = 0; j < C; ++j)
		{
			const int n = RWarp<RC>::col_of_slot(lane * C + j, mid - HW);
			if (n >= 1 && n <= min(mid + HW, (int)w.N - 1)) flush_slot<RC>(args, pc_off, ta, j, n);
		}
		double sm_ = ta.dM, se_ = ta.dE;
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
// GA (records-free layout, MODE 3): the row header is the C decision words; the path cell of every row (column | match
// state << 31) is left in pp[] for the gather sweep, which replaces it by its posterior; the medians follow that sweep.
template <class RC, bool GA = false>
DYN_DEV bool traceback_pass(RWarp<RC>& w, const SlotScratch& sc, const BatchArgs& args, const ReadDesc& rd)
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
		if (GA)
		{
			h[0] = h[1] = 0u;
#pragma unroll
			for (int jj = 0; jj < C; ++jj) h[2 + jj] = (row < T && row >= 1) ? sc.hdr[(size_t)row * C + jj] : 0u;
		}
		else
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
		if (GA)
		{
			if (mycell != 0xffffffffu) sc.pp[row] = __uint_as_float(mycell);
		}
		else if (mycell != 0xffffffffu)
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
	if (!GA) segment_medians_impl(w.lane, w.T, w.N, sc, args, rd);
	return true;
}

// one read, all passes.  A read the ribbon cannot represent leaves with ST_LIN_FAULT (host: full-band kernels).
template <class RC, int MODE, bool TL>
DYN_DEV void ribbon_read(const BatchArgs& args, const ReadDesc& rd, uint32_t ridx, const SlotScratch& sc,
	unsigned char* smem_raw, int lane)
{
	RWarp<RC> w;
	w.lane = lane;
	w.S = rd.S;
	w.T = rd.S + 1;
	w.N = rd.N;
	w.bw_ref = (int)rd.bw;
	w.ratio = rd.ratio;
	w.ratio_f = (float)rd.ratio;
	w.sig = args.signal + rd.sig_off;
	w.pc = args.pc + rd.pc_off;

	ReadOut out;
	out.Z = 0.0;
	out.dZ = 0.0;
	out.nrec = 0;
	out.status = ST_OK;
	out.xi_m = 0.0;
	out.xi_e = 0.0;

	// why a read is handed to the full-band kernels (ReadOut.nrec of a faulted read; dyn_last_ribbon counts them):
	//  1 nothing alive  2 backward mass at a window edge  3 nothing alive in a group clipped to the reference band (the
	//  alignment leaves the band)  4 window misses row 0
	//  5 posterior mass of a group != 1  6 range guard  7 forward mass at a window edge  8 posterior-Viterbi scores died
	//  9 Zf != Zb  10 record buffer full  11 traceback incomplete  12 Zb not finite
	int fault = 0;
	const double Z2 = (MODE == 0) ? rib::backward_pass<RC, false, TL>(w, sc, args, fault) : rib::backward_pass<RC, true, TL>(w, sc, args, fault);
	out.Z = Z2 * LN2;
	fault = warp_max_int(fault);
	if (!fault && !(Z2 > -1.0e30 && Z2 < 1.0e30)) fault = 12;
	if (!fault && MODE != 0)
	{
		__threadfence_block();
		__syncwarp();
		uint32_t nrec = 0;
		bool overflow = false;
		const double dz2 = rib::forward_pass<RC, MODE, TL>(w, sc, args, rd.pc_off, smem_raw, Z2, nrec, overflow, out.xi_m, out.xi_e, fault);
		out.nrec = nrec;
		out.dZ = dz2 * LN2;
		if (!fault && !(fabs(dz2) <= lin::LIN_Z_TOL)) fault = 9;
		if (!fault && MODE == 1)
		{
			__threadfence_block();
			__syncwarp();
			if (!rib::traceback_pass<RC>(w, sc, args, rd)) fault = 11;
		}
		if (!fault && MODE == 3)
		{
			// records-free layout (long reads): traceback over the decision bits, then a second forward sweep that evaluates
			// the posterior of every path cell (bit-identical to what MODE 1 would have recorded), then the medians
			__threadfence_block();
			__syncwarp();
			if (!rib::traceback_pass<RC, true>(w, sc, args, rd)) fault = 11;
			if (!fault)
			{
				__threadfence_block();
				__syncwarp();
				double xm, xe;
				const double dz3 = rib::forward_pass<RC, 4, TL>(w, sc, args, rd.pc_off, smem_raw, Z2, nrec, overflow, xm, xe, fault);
				if (!fault && !(fabs(dz3) <= lin::LIN_Z_TOL)) fault = 9;
				if (!fault)
				{
					__threadfence_block();
					__syncwarp();
					segment_medians_impl(w.lane, w.T, w.N, sc, args, rd);
				}
			}
		}
	}
	if (fault)
	{
		out.status = ST_LIN_FAULT;
		out.nrec = (uint32_t)fault;
	}
	if (lane == 0) args.out[ridx] = out;
}

} // namespace rib
} // namespace dyn
