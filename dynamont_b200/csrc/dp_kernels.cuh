// Warp-per-read banded forward/backward/posterior-decoding kernels for Dynamont's basic ("NT") mode.
//
// What the reference computes (NT_aligner_api.cpp:110-456, restated in SURVEY.md Appendix A) and how it is
// mapped here:
//
//   * one warp owns one read; the band row (<= 2*bw+1 lattice columns) lives in registers, CPL consecutive
//     columns per lane ("ring slots": column n -> slot n mod 32*CPL, lane = slot / CPL).  The band slides by
//     at most one column per row, which retires one slot and activates another; nothing else moves.
//   * the n-1 / n+1 neighbour of a lane's first / last column comes from one warp shuffle per row.
//   * all arithmetic is FP32 in the log2 domain; every lane carries its own double offset (block floating
//     point), renormalised from the lane-local maximum every RN rows.  No cross-lane reduction is needed.
//   * pass 1 (backward, t descending) keeps only a checkpoint every CK rows (bM, bE slots + lane offsets).
//   * pass 2 (t ascending) recomputes backward rows block-wise into shared memory, runs the forward
//     recurrence normalised by the backward offsets (so that fM+bM and fE+bE ARE the log2 posteriors),
//     feeds them into the posterior-Viterbi fill (NT:338-363), stores 1 decision bit per cell and the few
//     cells per row whose posterior is not negligible (sparse records).
//   * pass 3 walks the decision bits backwards (NT:383-456), looks the path cells' posteriors up in the
//     sparse records (normalising every row by its own recorded mass) and takes the per-segment median.
//
// Nothing lattice-sized is ever written to HBM except 1 bit per cell.
#pragma once

#include "dp_common.cuh"

namespace dyn
{

template <int CPL_, int CK_, int RN_, int RV_, bool UNI_ = false>
struct Cfg
{
	// UNI: every kmer of the pore model has the same standard deviation (both shipped 5-mer models and the synthetic
	// 9-mer models): the emission constants a and c are warp-uniform scalars, only b = mu * a is per column; an inactive
	// ring slot is gated through b (z = x*a - 1e18 -> 2^-1e36 = 0) instead of c.  26 registers fewer per thread.
	static constexpr bool UNI = UNI_;
	static constexpr int CPL = CPL_;         // lattice columns per lane
	static constexpr int SLOTS = 32 * CPL_;  // ring capacity; needs 2*bw + 2 <= SLOTS
	static constexpr int CK = CK_;           // checkpoint spacing (rows); divides 32, multiple of RN
	static constexpr int RN = RN_;           // renormalisation period of the backward pass (rows), power of two
	static constexpr int RV = RV_;           // renormalisation period of the posterior-Viterbi scores, power of two
	static constexpr int NRN = CK_ / RN_ + 1;
	static constexpr int CKF = 2 * CPL_ * 32;  // floats per checkpoint
	static constexpr int ROWF = CPL_ * 32;     // floats per shared-memory row
	static constexpr size_t SMEM_BYTES = (size_t)(CK_ + 2) * ROWF * 4 + (size_t)NRN * 32 * (8 + 4);  // rows t_lo .. t_lo+CK, + 1 row the fast forward row may prefetch past the block
	static_assert(CK_ % RN_ == 0 && 32 % CK_ == 0, "CK must divide 32 and be a multiple of RN");
	static_assert((RN_ & (RN_ - 1)) == 0 && (RV_ & (RV_ - 1)) == 0, "RN, RV: powers of two");
};

// Scratch memory of one resident warp ("slot"), sized by the host for the longest read of the batch.
struct SlotScratch
{
	float* ckpt;        // [nck][CKF]
	double* ckpt_ob;    // [nck][32]
	uint16_t* bits;     // [T][32]   decision bits: bit j of word (t, lane) set <=> cell in slot lane*CPL+j came from E
	uint32_t* rowptr;   // [T+1]     first sparse record of row t
	void* recs;         // [rec_cap] LaneRec<CPL>
	uint32_t* pn;       // [T]       path column of row t (bit 31: match state)
	float* pp;          // [T]       posterior of the path cell of row t
	// ribbon kernels (dp_ribbon.cuh)
	uint2* sched;       // [T/32+2]  window schedule: {window centre of row 32c, slide bits of rows 32c .. 32c+31}
	uint32_t* hdr;      // [T+1][HDRW] row headers: first record, hot-lane mask, decision words
	float* ring;        // two-level checkpoints: the 8 group checkpoints of the current super-group (+ their lane offsets)
};

struct BatchArgs
{
	const ReadDesc* reads;
	const uint32_t* order;      // processing order (longest first)
	uint32_t n_reads;
	uint32_t* queue;            // atomic work counter
	const float* signal;
	const PosConst* pc;
	const SlotScratch* slots;   // [gridDim.x]
	uint64_t rec_cap;           // sparse records per slot
	ReadOut* out;               // [n_reads]
	uint32_t* out_sigpos;       // [sum Kc]
	double* out_prob;           // [sum Kc]
	float m1, e2;               // log2 transition scores (e1 = log 1 = 0 is omitted, NT:33,85)
	float thr2;                 // sparse-record threshold on max(log2 pM, log2 pE)
	float m1_lin, e2_lin;       // the same transitions as plain probabilities (linear-domain kernels)
	float thr_lin;              // 2^thr2
	int mode;                   // 0: Z only (backward pass), 1: full alignment, 2: training statistics
	float uni_a, uni_c;         // Cfg::UNI kernels: the model-wide emission constants (every kmer has the same sigma)
	int fwd_fast;               // linear-domain pass 2: branch-free row body for groups of rows without a band slide
	// ribbon kernels (dp_ribbon.cuh)
	uint32_t n_slots;           // resident warps (scratch slots)
	float thr_rib;              // a lane is recorded when one of its posteriors exceeds this
	int rib_guard;              // the window's edge lanes must stay this many bits below the row maximum
	// training (mode 2)
	double* stat_w;             // [K] pooled sum of gamma              (NT:510)
	double* stat_x;             // [K] pooled sum of gamma * x          (NT:511)
	double* stat_xx;            // [K] pooled sum of gamma * x^2        (NT:512)
	const int32_t* kmers;       // [sum N] kmer id of column n (kmer[n-1]; entry 0 unused), parallel to pc
	double* read_w;             // per-read per-column statistics [sum N]
	double* read_x;
	double* read_xx;
};

// ------------------------------------------------------------------------------------------------------
// ring-slot helpers
// ------------------------------------------------------------------------------------------------------
DYN_DEV int pmod(int x, int m)
{
	int r = x % m;
	return r < 0 ? r + m : r;
}

template <int CPL>
struct Emis
{
	float a[CPL], b[CPL], c[CPL];
};

#define DYN_SLOT_CASE(J) \
	case J:              \
		if (J < CPL && lane == ql) { body(J); } \
		break;

// calls body(j) in the lane that owns ring slot q, with j a compile-time constant (registers stay registers)
template <int CPL, typename F>
DYN_DEV void with_slot(int lane, int q, F body)
{
	const int ql = q / CPL;
	const int j = q - ql * CPL;
	switch (j)
	{
		DYN_SLOT_CASE(0) DYN_SLOT_CASE(1) DYN_SLOT_CASE(2) DYN_SLOT_CASE(3)
		DYN_SLOT_CASE(4) DYN_SLOT_CASE(5) DYN_SLOT_CASE(6) DYN_SLOT_CASE(7)
		DYN_SLOT_CASE(8) DYN_SLOT_CASE(9) DYN_SLOT_CASE(10) DYN_SLOT_CASE(11)
		DYN_SLOT_CASE(12) DYN_SLOT_CASE(13) DYN_SLOT_CASE(14) DYN_SLOT_CASE(15)
		DYN_SLOT_CASE(16) DYN_SLOT_CASE(17) DYN_SLOT_CASE(18) DYN_SLOT_CASE(19)
		DYN_SLOT_CASE(20) DYN_SLOT_CASE(21) DYN_SLOT_CASE(22) DYN_SLOT_CASE(23)
		DYN_SLOT_CASE(24) DYN_SLOT_CASE(25) DYN_SLOT_CASE(26) DYN_SLOT_CASE(27)
		DYN_SLOT_CASE(28) DYN_SLOT_CASE(29) DYN_SLOT_CASE(30) DYN_SLOT_CASE(31)
	default:
		break;
	}
}
#undef DYN_SLOT_CASE

// slot functors (compile-time slot index keeps the arrays in registers)
template <int CPL>
struct SetEmis
{
	Emis<CPL>& e;
	float va, vb, vc;
	DYN_DEV void operator()(int j) const
	{
#pragma unroll
		for (int jj = 0; jj < CPL; ++jj)
			if (jj == j)
			{
				e.a[jj] = va;
				e.b[jj] = vb;
				e.c[jj] = vc;
			}
	}
};

// set the emission constants of a slot and force its two state values
template <int CPL>
struct SetEmisState
{
	Emis<CPL>& e;
	float va, vb, vc;
	float (&p)[CPL];
	float (&q)[CPL];
	float v;
	DYN_DEV void operator()(int j) const
	{
#pragma unroll
		for (int jj = 0; jj < CPL; ++jj)
			if (jj == j)
			{
				e.a[jj] = va;
				e.b[jj] = vb;
				e.c[jj] = vc;
				p[jj] = v;
				q[jj] = v;
			}
	}
};

template <int CPL>
struct SetOne
{
	float (&p)[CPL];
	float v;
	DYN_DEV void operator()(int j) const
	{
#pragma unroll
		for (int jj = 0; jj < CPL; ++jj)
			if (jj == j) p[jj] = v;
	}
};

template <int CPL>
struct GetOne
{
	const float (&p)[CPL];
	float& out;
	DYN_DEV void operator()(int j) const
	{
#pragma unroll
		for (int jj = 0; jj < CPL; ++jj)
			if (jj == j) out = p[jj];
	}
};

// The warp-uniform view of one read.
template <class CFG>
struct Warp
{
	static constexpr int CPL = CFG::CPL;
	static constexpr int SLOTS = CFG::SLOTS;

	int lane;
	uint32_t T, N, S;
	int bw;
	double ratio;
	const float* sig;
	const PosConst* pc;
	float m1, e2;
	float ua, uc;  // UNI: the model-wide emission constants a, c
	Emis<CPL> em;

	DYN_DEV bool valid_col(int n) const { return n >= 0 && n < (int)N; }
	// log2 emission of ring slot j (compile-time j) for sample x
	DYN_DEV float emis(int j, float x) const
	{
		return CFG::UNI ? emis2(x, ua, em.b[j], uc) : emis2(x, em.a[j], em.b[j], em.c[j]);
	}
	// b constant of a column (UNI: a column that scores no kmer is gated through b)
	DYN_DEV static float bsel(const PosConst& v) { return (CFG::UNI && v.c < 0.5f * CNEG) ? BIGB : v.b; }
	DYN_DEV static float b_off() { return CFG::UNI ? BIGB : 0.0f; }

	DYN_DEV void activate(int n)
	{
		if (!valid_col(n)) return;
		const PosConst v = pc[n];  // uniform address: one broadcast transaction
		with_slot<CPL>(lane, pmod(n, SLOTS), SetEmis<CPL>{em, v.a, bsel(v), v.c});
	}
	DYN_DEV void deactivate(int n)
	{
		if (!valid_col(n)) return;
		with_slot<CPL>(lane, pmod(n, SLOTS), SetEmis<CPL>{em, 0.0f, b_off(), CNEG});
	}
	// lattice column held by ring slot q when the window starts at column n0 (may be negative)
	DYN_DEV int col_of_slot(int q, int n0) const { return n0 + pmod(q - n0, SLOTS); }

	// load the emission constants of every column of band(row with centre mid); everything else inactive
	DYN_DEV void load_window(int mid)
	{
		const int n0 = mid - bw;
		const int nlast = min(mid + bw, (int)N - 1);
#pragma unroll
		for (int j = 0; j < CPL; ++j)
		{
			const int n = col_of_slot(lane * CPL + j, n0);
			if (n >= 0 && n <= nlast)
			{
				const PosConst v = pc[n];
				em.a[j] = v.a;
				em.b[j] = bsel(v);
				em.c[j] = v.c;
			}
			else
			{
				em.a[j] = 0.0f;
				em.b[j] = b_off();
				em.c[j] = CNEG;
			}
		}
	}
	// move the emission window from centre mid_from up to centre mid_to (mid_to >= mid_from)
	DYN_DEV void slide_window_up(int mid_from, int mid_to)
	{
#pragma unroll 1
		for (int m = mid_from + 1; m <= mid_to; ++m)
		{
			deactivate(m - 1 - bw);
			activate(m + bw);
		}
	}
};

// per 32-row chunk: the lane's sample and the band-slide mask (bit i <=> centre(base+i) != centre(base+i+1))
struct Chunk
{
	float xv;
	unsigned smask;
};

template <class CFG>
DYN_DEV Chunk chunk_load(const Warp<CFG>& w, uint32_t base)
{
	Chunk c;
	const uint32_t r = base + w.lane;
	const uint32_t c0 = band_mid(r, w.ratio), c1 = band_mid(r + 1, w.ratio);
	c.smask = __ballot_sync(FULL, c0 != c1);
	c.xv = (r < w.S) ? w.sig[r] : 0.0f;
	return c;
}

// ------------------------------------------------------------------------------------------------------
// backward recurrence (NT_aligner_api.cpp:158-207)
// ------------------------------------------------------------------------------------------------------
template <int CPL>
struct Bwd
{
	float bM[CPL], bE[CPL];
	double OB;  // this lane's values are (true log2 value) - OB
	float dR;   // OB(right neighbour lane) - OB(this lane)
};

template <class CFG>
DYN_DEV void bwd_row(Warp<CFG>& w, Bwd<CFG::CPL>& b, float x)
{
	constexpr int CPL = CFG::CPL;
	float s[CPL], A[CPL];
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		s[j] = w.emis(j, x);
		A[j] = b.bM[j] + (s[j] + w.m1);  // bM[t+1][n] + score(x[t], kmer[n-1]) + m1, consumed by column n-1
	}
	const float Ar = __shfl_sync(FULL, A[0], (w.lane + 1) & 31) + b.dR;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		const float ext1 = (j + 1 < CPL) ? A[j + 1] : Ar;
		const float nm = b.bE[j] + s[j];      // bM[t][n] = bE[t+1][n] + score      (NT:200)
		b.bE[j] = logplus2(ext1, nm + w.e2);  //                                    (NT:194,201,204)
		b.bM[j] = nm;
	}
}

// lane-local renormalisation; returns the increment that was subtracted (0 for a lane without finite cells)
template <class CFG>
DYN_DEV float bwd_renorm(Warp<CFG>& w, Bwd<CFG::CPL>& b)
{
	constexpr int CPL = CFG::CPL;
	float lm = b.bE[0];
#pragma unroll
	for (int j = 1; j < CPL; ++j) lm = fmaxf(lm, b.bE[j]);
	const bool dead = lm < DEADT;
	const float inc = dead ? 0.0f : lm;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		b.bM[j] -= inc;
		b.bE[j] -= inc;
	}
	b.OB += (double)inc;
	// a lane without finite cells adopts its right neighbour's offset, so that the first values that
	// flow into it later are represented at a sane scale
	double obr = shfl_f64(b.OB, (w.lane + 1) & 31);
	if (dead) b.OB = obr;
	obr = shfl_f64(b.OB, (w.lane + 1) & 31);
	b.dR = (float)(obr - b.OB);
	return inc;
}

// one backward step: row t from row t+1, including the band slide between the two rows
template <class CFG>
DYN_DEV void bwd_step(Warp<CFG>& w, Bwd<CFG::CPL>& b, float x, bool slide, int& mid)
{
	constexpr int CPL = CFG::CPL;
	// on entry mid = centre of row t+1
	if (slide)
	{
		// column mid_t - bw enters the band.  While it was the column just below the band its slot kept
		// computing bE from its (in-band) right neighbour — the M-transition term carries the neighbour's
		// emission, not its own, so an inactive slot does not gate it.  (t+1, n) is out of band: force -inf.
		const int nb = mid - 1 - w.bw;
		if (nb >= 0)
		{
			const PosConst v = w.pc[nb];
			with_slot<CPL>(w.lane, pmod(nb, CFG::SLOTS), SetEmisState<CPL>{w.em, v.a, w.bsel(v), v.c, b.bM, b.bE, NEG});
		}
	}
	bwd_row<CFG>(w, b, x);
	if (slide)
	{
		const int ntop = mid + w.bw;  // column that was in band(t+1) but is not in band(t)
		if (ntop < (int)w.N)
			with_slot<CPL>(w.lane, pmod(ntop, CFG::SLOTS), SetEmisState<CPL>{w.em, 0.0f, w.b_off(), CNEG, b.bM, b.bE, NEG});
		--mid;
	}
}

template <class CFG>
DYN_DEV void bwd_init_terminal(Warp<CFG>& w, Bwd<CFG::CPL>& b)
{
	constexpr int CPL = CFG::CPL;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		b.bM[j] = NEG;
		b.bE[j] = NEG;
	}
	with_slot<CPL>(w.lane, pmod((int)w.N - 1, CFG::SLOTS), SetOne<CPL>{b.bE, 0.0f});  // bE[T-1][N-1] = 0 (NT:170)
	b.OB = 0.0;
	b.dR = 0.0f;
}

template <class CFG>
DYN_DEV void ckpt_store(const SlotScratch& sc, uint32_t idx, int lane, const Bwd<CFG::CPL>& b)
{
	constexpr int CPL = CFG::CPL;
	float* f = sc.ckpt + (size_t)idx * CFG::CKF;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		f[j * 32 + lane] = b.bM[j];
		f[(CPL + j) * 32 + lane] = b.bE[j];
	}
	sc.ckpt_ob[(size_t)idx * 32 + lane] = b.OB;
}

template <class CFG>
DYN_DEV void ckpt_load(const SlotScratch& sc, uint32_t idx, int lane, Bwd<CFG::CPL>& b)
{
	constexpr int CPL = CFG::CPL;
	const float* f = sc.ckpt + (size_t)idx * CFG::CKF;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		b.bM[j] = f[j * 32 + lane];
		b.bE[j] = f[(CPL + j) * 32 + lane];
	}
	b.OB = sc.ckpt_ob[(size_t)idx * 32 + lane];
	const double obr = shfl_f64(b.OB, (lane + 1) & 31);
	b.dR = (float)(obr - b.OB);
}

#ifdef DYN_DEBUG_ROWS
static double* g_dbg_rows = nullptr;  // emulator-only debugging aid: [T][N][2] true log2 values of bM, bE
#endif

// ------------------------------------------------------------------------------------------------------
// pass 1: backward over the whole read.  Returns log2 Zb (double) in every lane; stores checkpoints.
// ------------------------------------------------------------------------------------------------------
template <class CFG, bool STORE>
DYN_DEV double backward_pass(Warp<CFG>& w, const SlotScratch& sc)
{
	constexpr int CPL = CFG::CPL;
	Bwd<CPL> b;
	int mid = (int)band_mid(w.T - 1, w.ratio);
	w.load_window(mid);
	bwd_init_terminal<CFG>(w, b);
	if (STORE && ((w.T - 1) & (CFG::CK - 1)) == 0) ckpt_store<CFG>(sc, (w.T - 1) / CFG::CK, w.lane, b);

	int t = (int)w.T - 2;
	// the next chunk's samples are requested one chunk ahead so their latency hides behind 32 rows of work
	Chunk nxt = chunk_load<CFG>(w, (uint32_t)t & ~31u);
	while (t >= 0)
	{
		const uint32_t base = (uint32_t)t & ~31u;
		const Chunk cur = nxt;
		if (base >= 32) nxt = chunk_load<CFG>(w, base - 32);
		float x = __shfl_sync(FULL, cur.xv, t - (int)base);
#pragma unroll 1
		for (int i = t - (int)base; i >= 0; --i)
		{
			const uint32_t tt = base + i;
			const float xn = __shfl_sync(FULL, cur.xv, (i - 1) & 31);  // next row's sample, off the critical path
			bwd_step<CFG>(w, b, x, (cur.smask >> i) & 1u, mid);
			x = xn;
			if ((tt & (CFG::RN - 1)) == 0)
			{
				bwd_renorm<CFG>(w, b);
				if (STORE && (tt & (CFG::CK - 1)) == 0) ckpt_store<CFG>(sc, tt / CFG::CK, w.lane, b);
			}
#ifdef DYN_DEBUG_ROWS
			if (g_dbg_rows)
			{
#pragma unroll
				for (int j = 0; j < CPL; ++j)
				{
					const int n = w.col_of_slot(w.lane * CPL + j, mid - w.bw);
					if (n >= 0 && n < (int)w.N && n <= mid + w.bw)
					{
						g_dbg_rows[((size_t)tt * w.N + n) * 2] = (double)b.bM[j] + b.OB;
						g_dbg_rows[((size_t)tt * w.N + n) * 2 + 1] = (double)b.bE[j] + b.OB;
					}
				}
			}
#endif
		}
		t = (int)base - 1;
	}
	// Zb = bE[0][0] (NT:286): column 0 is ring slot 0 = lane 0, j 0
	const double z = (double)b.bE[0] + b.OB;
	return shfl_f64(z, 0);
}

// ------------------------------------------------------------------------------------------------------
// pass 2 state
// ------------------------------------------------------------------------------------------------------
template <int CPL>
struct Fwd
{
	float fM[CPL], fE[CPL];
	float VM[CPL], VE[CPL];
	double OF;  // forward values are (true log2 value) - OF, with OF tracking Z2 - OB so that f + b = log2 posterior
	double OV;  // Viterbi values are (true value) - OV
	float dL;   // OF(left lane) - OF(this lane)
	float dVL;  // OV(left lane) - OV(this lane)
};

template <class CFG>
struct Smem
{
	float* bE;     // [(CK+1)][CPL][32]
	double* OB;    // [NRN][32]   lane offsets of the backward pass at rows t_lo + i*RN
	float* inc;    // [NRN][32]   increment subtracted at those rows
	DYN_DEV explicit Smem(unsigned char* p)
	{
		OB = reinterpret_cast<double*>(p);
		bE = reinterpret_cast<float*>(p + (size_t)CFG::NRN * 32 * 8);
		inc = bE + (size_t)(CFG::CK + 1) * CFG::ROWF;
	}
};

// re-express the forward values of this lane relative to the offset target (a double); tracks the offset
// actually applied so that rounding of the shift never accumulates
template <class CFG>
DYN_DEV void fwd_shift_to(Warp<CFG>& w, Fwd<CFG::CPL>& f, double target)
{
	constexpr int CPL = CFG::CPL;
	const float sh = (float)(f.OF - target);
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		f.fM[j] += sh;
		f.fE[j] += sh;
	}
	f.OF -= (double)sh;
	const double ofl = shfl_f64(f.OF, (w.lane + 31) & 31);
	f.dL = (float)(ofl - f.OF);
}

template <class CFG>
DYN_DEV void vit_renorm(Warp<CFG>& w, Fwd<CFG::CPL>& f)
{
	constexpr int CPL = CFG::CPL;
	float lm = fmaxf(f.VM[0], f.VE[0]);
#pragma unroll
	for (int j = 1; j < CPL; ++j) lm = fmaxf(lm, fmaxf(f.VM[j], f.VE[j]));
	const bool dead = lm < DEADT;
	const float inc = dead ? 0.0f : lm;
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		f.VM[j] -= inc;
		f.VE[j] -= inc;
	}
	f.OV += (double)inc;
	double ovl = shfl_f64(f.OV, (w.lane + 31) & 31);
	if (dead) f.OV = ovl;
	ovl = shfl_f64(f.OV, (w.lane + 31) & 31);
	f.dVL = (float)(ovl - f.OV);
}

struct RecSink
{
	void* recs;
	uint64_t cap;
	uint32_t n;
	bool overflow;
};

// One row of pass 2.  On entry f holds the forward values of row t and the Viterbi values of row t-1.
//   DO_V:    posteriors of row t, posterior-Viterbi update (NT:357-362), decision bits, sparse records
//   DO_STEP: forward recurrence to row t+1 (NT:141-150) incl. the band slide between t and t+1
// bc/bn: this lane's values of the recomputed backward rows bE(t) / bE(t+1), already in registers; pf (or NULL):
// shared-memory row bE(t+1), from which bc/bn of the NEXT row are fetched before the MUFU-heavy forward step so
// that their latency is hidden.
template <class CFG, bool DO_V, bool DO_STEP>
DYN_DEV void fwd_row(Warp<CFG>& w, Fwd<CFG::CPL>& f, const SlotScratch& sc, RecSink& rs, float thr2, uint32_t t,
	float x, bool slide, int& mid_f, float (&bc)[CFG::CPL], float (&bn)[CFG::CPL], const float* pf, bool rn_row, float inc_t)
{
	constexpr int CPL = CFG::CPL;
	const int lane = w.lane;
	if (DO_STEP && slide) w.activate(mid_f + 1 + w.bw);  // column entering band(t+1)

	float s[CPL], LPM[CPL], LPE[CPL];
#pragma unroll
	for (int j = 0; j < CPL; ++j)
	{
		s[j] = w.emis(j, x);
		if (DO_V)
		{
			LPE[j] = f.fE[j] + bc[j];
			// bM[t][n] = bE[t+1][n] + score(x[t], kmer[n-1]) (NT:200); the last row has no match state (-inf)
			LPM[j] = DO_STEP ? f.fM[j] + (bn[j] + s[j]) : NEG;
		}
	}
	if (DO_V)
	{
		if (DO_STEP && rn_row)
		{
			// the backward pass renormalised row t after forming bM[t]: apply the same increment
#pragma unroll
			for (int j = 0; j < CPL; ++j) LPM[j] -= inc_t;
		}
		if (DO_STEP && slide && mid_f - w.bw >= 0)
		{
			// column lo_t leaves the band at row t+1, so bE[t+1][lo_t] is out of band (-inf in the reference) and
			// with it bM[t][lo_t]; the recomputed row holds the ungated neighbour term there (see bwd_step)
			with_slot<CPL>(lane, pmod(mid_f - w.bw, CFG::SLOTS), SetOne<CPL>{LPM, NEG});
		}
		// posterior-Viterbi fill, in place from the highest slot down so that slot j-1 still holds row t-1;
		// decision bit = sign(VM - VE): set <=> the E state of this cell is entered from E (test of NT:448)
		const float vl = __shfl_sync(FULL, f.VE[CPL - 1], (lane + 31) & 31) + f.dVL;
		unsigned acc = 0;
		float lmax = NEG;
#pragma unroll
		for (int j = CPL - 1; j >= 0; --j)
		{
			const float vmx = fmaxf(f.VM[j], f.VE[j]);
			acc = __funnelshift_l(__float_as_uint(f.VM[j] - f.VE[j]), acc, 1);
			const float left = (j > 0) ? f.VE[j - 1] : vl;
			f.VM[j] = left + LPM[j];
			f.VE[j] = vmx + LPE[j];
			lmax = fmaxf(lmax, fmaxf(LPM[j], LPE[j]));
		}
		if ((t & (CFG::RV - 1)) == 0) vit_renorm<CFG>(w, f);
		sc.bits[(size_t)t * 32 + lane] = (uint16_t)acc;

		// sparse posterior records: every lane that holds a cell with a non-negligible posterior dumps its CPL
		// (match, extend) pairs with vector stores — typically one lane per row
		if (lane == 0) sc.rowptr[t] = rs.n;
		const bool hot = lmax > thr2;
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
					tmp[j] = LPM[j];
					tmp[CPL + j] = LPE[j];
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
		const float fl = __shfl_sync(FULL, f.fE[CPL - 1], (lane + 31) & 31) + f.dL;
		if (pf)
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
			const float ne = logplus2(f.fM[j], f.fE[j] + w.e2) + s[j];
			f.fM[j] = left + (s[j] + w.m1);
			f.fE[j] = ne;
		}
		if (slide)
		{
			const int nold = mid_f - w.bw;  // column of band(t) that is not in band(t+1)
			if (nold >= 0)
				with_slot<CPL>(lane, pmod(nold, CFG::SLOTS), SetEmisState<CPL>{w.em, 0.0f, w.b_off(), CNEG, f.fM, f.fE, NEG});
			++mid_f;
		}
	}
}

// ------------------------------------------------------------------------------------------------------
// pass 2: forward + posterior + posterior-Viterbi fill.  Returns (Zf - Zb) in log2 units.
// ------------------------------------------------------------------------------------------------------
template <class CFG>
DYN_DEV float forward_posterior_pass(Warp<CFG>& w, const SlotScratch& sc, const BatchArgs& args, unsigned char* smem_raw,
	double Z2, uint32_t& nrec_out, bool& overflow)
{
	constexpr int CPL = CFG::CPL;
	constexpr int CK = CFG::CK;
	constexpr int RN = CFG::RN;
	constexpr int ROWF = CFG::ROWF;
	Smem<CFG> sm(smem_raw);
	Fwd<CPL> f;
	Bwd<CPL> b;
	const int lane = w.lane;
	const uint32_t T = w.T;
	RecSink rs;
	rs.recs = sc.recs;
	rs.cap = args.rec_cap;
	rs.n = 0;
	rs.overflow = false;

	int mid_f = 0;  // band centre of the forward row
	float bc[CPL], bn[CPL];  // backward rows bE(t), bE(t+1) of the current forward row
#pragma unroll
	for (int j = 0; j < CPL; ++j) bc[j] = bn[j] = NEG;
	w.load_window(0);
	const uint32_t kb = (T - 1) / CK;
	// samples/slide masks per 32-row chunk: a block (CK | 32) lies inside one chunk; both the backward
	// recomputation and the forward rows of the block use it; the next chunk is requested a chunk ahead
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
		const bool from_ckpt = (t_hi <= T - 1);
		const uint32_t src_row = from_ckpt ? t_hi : T - 1;
		int mid_b = (int)band_mid(src_row, w.ratio);
		w.slide_window_up(mid_f, mid_b);
		if (from_ckpt) ckpt_load<CFG>(sc, k + 1, lane, b);
		else bwd_init_terminal<CFG>(w, b);
		{
			float* dst = sm.bE + (size_t)(src_row - t_lo) * ROWF;
#pragma unroll
			for (int j = 0; j < CPL; ++j) dst[j * 32 + lane] = b.bE[j];
			// offsets in force for the rows above the last renormalisation row of this block
			sm.OB[((src_row - t_lo + RN - 1) / RN) * 32 + lane] = b.OB;
		}
		// the slide between src_row-1 and src_row may belong to the next chunk when src_row is 32-aligned
		float xb = __shfl_sync(FULL, cur.xv, ((int)src_row - 1) & 31);
#pragma unroll 1
		for (int tt = (int)src_row - 1; tt >= (int)t_lo; --tt)
		{
			const int i = tt & 31;
			const float x = xb;
			xb = __shfl_sync(FULL, cur.xv, (i - 1) & 31);
			bwd_step<CFG>(w, b, x, (cur.smask >> i) & 1u, mid_b);
			float* dst = sm.bE + (size_t)(tt - (int)t_lo) * ROWF;
			if ((tt & (RN - 1)) == 0)
			{
				const float inc = bwd_renorm<CFG>(w, b);
				sm.OB[((tt - (int)t_lo) / RN) * 32 + lane] = b.OB;
				sm.inc[((tt - (int)t_lo) / RN) * 32 + lane] = inc;
			}
#pragma unroll
			for (int j = 0; j < CPL; ++j) dst[j * 32 + lane] = b.bE[j];
		}
		__syncwarp();
		// the emission window is band(t_lo) again (mid_b == mid_f)

		// ---- step b: forward rows t_lo .. min(t_hi, T) - 1 ----------------------------------------------
		uint32_t t = t_lo;
		if (k == 0)
		{
			// row 0: fE[0][0] = 0 (NT:120) expressed relative to OF = Z2 - OB(0); VE[0][0] = 0 (NT:336)
#pragma unroll
			for (int j = 0; j < CPL; ++j)
			{
				f.fM[j] = NEG;
				f.fE[j] = NEG;
				f.VM[j] = NEG;
				f.VE[j] = NEG;
			}
			f.OF = Z2 - sm.OB[lane];
			f.OV = 0.0;
			f.dVL = 0.0f;
			if (lane == 0)
			{
				f.fE[0] = -sm.bE[0];  // exact: Z2 - OB(0)[lane 0] = bE[0][0]
				f.VE[0] = 0.0f;
			}
			const double ofl = shfl_f64(f.OF, (lane + 31) & 31);
			f.dL = (float)(ofl - f.OF);
			// row 0 only steps the forward recurrence (T >= 2, so row 0 is never the last row)
			fwd_row<CFG, false, true>(w, f, sc, rs, args.thr2, 0, __shfl_sync(FULL, cur.xv, 0), cur.smask & 1u, mid_f,
				bc, bn, nullptr, false, 0.0f);
			fwd_shift_to<CFG>(w, f, Z2 - sm.OB[32 + lane]);  // row 0 is a renormalisation row
			t = 1;
		}
		const uint32_t t_end = min(t_hi, T - 1);  // regular rows: 1 <= t <= T-2
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
		float x = __shfl_sync(FULL, cur.xv, t & 31);
#pragma unroll 1
		for (; t < t_end; ++t)
		{
			const int i = t & 31;
			const uint32_t r = t - t_lo;
			const bool rn_row = (t & (RN - 1)) == 0;
			const float inc_t = rn_row ? sm.inc[(r / RN) * 32 + lane] : 0.0f;
			const float xn = __shfl_sync(FULL, cur.xv, (i + 1) & 31);
			fwd_row<CFG, true, true>(w, f, sc, rs, args.thr2, t, x, (cur.smask >> i) & 1u, mid_f, bc, bn,
				(t + 1 < t_end) ? sm.bE + (size_t)(r + 1) * ROWF : nullptr, rn_row, inc_t);
			if (rn_row) fwd_shift_to<CFG>(w, f, Z2 - sm.OB[(r / RN + 1) * 32 + lane]);
			x = xn;
		}
		__syncwarp();
	}
	// last row T-1: posteriors, Viterbi, bits, records; no forward step, no match posterior
	{
		const float* row = sm.bE + (size_t)((T - 1) - kb * CK) * ROWF;
#pragma unroll
		for (int j = 0; j < CPL; ++j) bc[j] = row[j * 32 + lane];
		fwd_row<CFG, true, false>(w, f, sc, rs, args.thr2, T - 1, 0.0f, false, mid_f, bc, bn, nullptr, false, 0.0f);
	}
	// (Zf - Zb) in log2 units = fE[T-1][N-1] + bE[T-1][N-1] with bE = 0 there (NT:170,285-286)
	float v = 0.0f;
	with_slot<CPL>(lane, pmod((int)w.N - 1, CFG::SLOTS), GetOne<CPL>{f.fE, v});
	const float lpe_end = __shfl_sync(FULL, v, pmod((int)w.N - 1, CFG::SLOTS) / CPL);
	if (lane == 0) sc.rowptr[T] = rs.n;
	nrec_out = rs.n;
	overflow = rs.overflow;
	return lpe_end;
}

// ------------------------------------------------------------------------------------------------------
// pass 3: traceback (NT:383-456), path posteriors, per-segment medians (aligner.cpp:247-263)
// ------------------------------------------------------------------------------------------------------
// median of d <= 32 values (one per lane; lanes >= d hold a value above every real one): sort them across the lanes, lane
// k then owns the k-th smallest.  Returns (k-th smallest, (k-1)-th smallest) broadcast to all lanes.
DYN_DEV void rank_select(float v, uint32_t d, uint32_t k, int lane, float& kth, float& prev)
{
	// bitonic sorting network over the 32 lanes, ascending: 15 compare-exchange steps (one shuffle + one predicated
	// min/max each) instead of the all-pairs ranking of the first version (32 shuffles + 32 compares: 236 instructions per
	// segment against ~80).  Lanes >= d hold a value above every real one and end up on top.
	(void)d;
#pragma unroll
	for (int k2 = 2; k2 <= 32; k2 <<= 1)
	{
#pragma unroll
		for (int j = k2 >> 1; j > 0; j >>= 1)
		{
			const float o = __shfl_sync(FULL, v, lane ^ j);
			const bool keep_min = ((lane & k2) == 0) == ((lane & j) == 0);
			v = keep_min ? fminf(v, o) : fmaxf(v, o);
		}
	}
	kth = __shfl_sync(FULL, v, (int)k);
	prev = (k > 0) ? __shfl_sync(FULL, v, (int)k - 1) : kth;
}

// the same for d <= 32*NV non-negative values held in registers (value i in lane i % 32, slot i / 32; unused slots hold
// a value above every real one): bisection on the bit pattern, one vote per slot and step.  prev = (k-1)-th smallest.
DYN_DEV uint32_t warp_min_u32(uint32_t v, int lane)
{
#ifndef DYN_HOST_EMU
	(void)lane;
	return __reduce_min_sync(FULL, v);
#else
	for (int o = 16; o; o >>= 1) v = min(v, (uint32_t)__shfl_sync(FULL, (int)v, (lane + o) & 31));
	return v;
#endif
}
DYN_DEV uint32_t warp_max_u32(uint32_t v, int lane)
{
#ifndef DYN_HOST_EMU
	(void)lane;
	return __reduce_max_sync(FULL, v);
#else
	for (int o = 16; o; o >>= 1) v = max(v, (uint32_t)__shfl_sync(FULL, (int)v, (lane + o) & 31));
	return v;
#endif
}

template <int NV>
DYN_DEV void reg_select(const float (&v)[NV], uint32_t k, int lane, float& kth, float& prev)
{
	// bisection on the bit pattern, started from the smallest / largest real value (unused slots hold PAD = 3e38): the
	// path posteriors of a dwell mostly lie within a few percent of each other, which halves the number of steps
	uint32_t mn = 0x7f800000u, mx = 0u;
#pragma unroll
	for (int q = 0; q < NV; ++q)
	{
		const uint32_t bits = __float_as_uint(v[q]);
		mn = min(mn, bits);
		mx = max(mx, bits < 0x7f000000u ? bits : 0u);
	}
	uint32_t lo = warp_min_u32(mn, lane), hi = max(warp_max_u32(mx, lane), lo);
	while (lo < hi)
	{
		const uint32_t midv = lo + (hi - lo) / 2;
		uint32_t cnt = 0;
#pragma unroll
		for (int q = 0; q < NV; ++q) cnt += __popc(__ballot_sync(FULL, __float_as_uint(v[q]) <= midv));
		if (cnt >= k + 1) hi = midv;
		else lo = midv + 1;
	}
	kth = __uint_as_float(lo);
	// sorted s[k] = kth.  s[k-1] = kth when fewer than k values lie strictly below it, else the largest of those
	uint32_t below = 0, mxb = 0;
#pragma unroll
	for (int q = 0; q < NV; ++q)
	{
		const uint32_t bits = __float_as_uint(v[q]);
		const bool lt = bits < lo;
		below += __popc(__ballot_sync(FULL, lt));
		mxb = max(mxb, lt ? bits : 0u);
	}
	mxb = warp_max_u32(mxb, lane);
	prev = (below == k && k > 0) ? __uint_as_float(mxb) : kth;
}

// k-th smallest (0-based) of d non-negative floats at v[0..d) (global memory), all lanes cooperating: binary
// search on the bit pattern with a warp-wide count per step
DYN_DEV float coop_select(const float* v, uint32_t d, uint32_t kth, int lane)
{
	uint32_t lo = 0u, hi = 0x7f800000u;
	while (lo < hi)
	{
		const uint32_t midv = lo + (hi - lo) / 2;
		uint32_t cnt = 0;
		for (uint32_t i = lane; i < d; i += 32) cnt += (__float_as_uint(v[i]) <= midv) ? 1u : 0u;
		cnt = __reduce_add_sync(FULL, cnt);
		if (cnt >= kth + 1) hi = midv;
		else lo = midv + 1;
	}
	return __uint_as_float(lo);
}

// phase 1 of pass 3: walk the decision bits, write the path (sc.pn) and the segment borders.  Returns false when
// the path is incomplete; t_first = first path row.  Shared by the log2-domain and the linear-domain kernels.
template <class CFG>
DYN_DEV bool trace_decisions(Warp<CFG>& w, const SlotScratch& sc, const BatchArgs& args, unsigned char* smem_raw,
	const ReadDesc& rd, uint32_t& t_first_out)
{
	constexpr int CPL = CFG::CPL;
	constexpr int SLOTS = CFG::SLOTS;
	const int lane = w.lane;
	const uint32_t T = w.T, N = w.N;
	uint32_t* border = args.out_sigpos + rd.out_off;  // Kc = N-1 entries
	uint16_t* sbits = reinterpret_cast<uint16_t*>(smem_raw);  // 32 rows x 32 lanes

	// walk the decision bits from (T-1, N-1) in state E (NT:398-452), one 32-row chunk at a time; within a
	// chunk every lane tests one row, so a whole run of E rows in one column is consumed per step
	int t = (int)T - 1, n = (int)N - 1;
	int inM = 0;
	bool done = false;
	// the decision words of the chunk below the current one are fetched while the current one is walked
	uint4 nb[4];
	{
		const uint32_t r = ((uint32_t)t & ~31u) + lane;
		const uint4* src = reinterpret_cast<const uint4*>(sc.bits + (size_t)r * 32);
#pragma unroll
		for (int q = 0; q < 4; ++q) nb[q] = (r < T) ? src[q] : make_uint4(0u, 0u, 0u, 0u);
	}
	while (!done)
	{
		const int cbase = t & ~31;
		{
			uint4* dst = reinterpret_cast<uint4*>(sbits + lane * 32);
#pragma unroll
			for (int q = 0; q < 4; ++q) dst[q] = nb[q];
		}
		__syncwarp();
		if (cbase >= 32)
		{
			const uint4* src = reinterpret_cast<const uint4*>(sc.bits + (size_t)(cbase - 32 + lane) * 32);
#pragma unroll
			for (int q = 0; q < 4; ++q) nb[q] = src[q];
		}
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
				if (lane == 0)
				{
					sc.pn[t] = (uint32_t)n | 0x80000000u;
					border[n - 1] = (uint32_t)t - 1;  // Segment.signalPosition (NT:424-430)
				}
				--t;
				--n;
				inM = 0;
				continue;
			}
			// extension run in column n: rows t, t-1, ... until the row whose E cell was entered from M (NT:443-451)
			const int q = n % SLOTS;
			const int ql = q / CPL, j = q - ql * CPL;
			const int row = cbase + lane;
			const bool valid = row <= t && row >= 1;
			const bool from_m = valid && (((sbits[lane * 32 + ql] >> j) & 1) == 0);
			const unsigned mk = __ballot_sync(FULL, from_m);
			int tstar;  // lowest row of the run inside this chunk
			if (mk)
			{
				tstar = cbase + (31 - __clz(mk));
				inM = 1;
			}
			else
				tstar = max(cbase, 1);
			if (row >= tstar && row <= t) sc.pn[row] = (uint32_t)n;
			t = tstar - 1;
		}
		__syncwarp();
	}
	// a complete path consumes every column: n == 0 and the first match sits at row 1
	t_first_out = (uint32_t)t + 1;  // first path row
	if (!((n == 0) && !inM)) return false;
	__threadfence_block();
	__syncwarp();
	return true;
}

DYN_DEV void segment_medians_impl(int lane, uint32_t T, uint32_t N, const SlotScratch& sc, const BatchArgs& args, const ReadDesc& rd);
template <class CFG>
DYN_DEV void segment_medians(Warp<CFG>& w, const SlotScratch& sc, const BatchArgs& args, const ReadDesc& rd)
{
	segment_medians_impl(w.lane, w.T, w.N, sc, args, rd);
}

template <class CFG>
DYN_DEV bool traceback_pass(Warp<CFG>& w, const SlotScratch& sc, const BatchArgs& args, unsigned char* smem_raw,
	const ReadDesc& rd)
{
	constexpr int CPL = CFG::CPL;
	constexpr int SLOTS = CFG::SLOTS;
	typedef LaneRec<CPL> Rec;
	const int lane = w.lane;
	const uint32_t T = w.T;
	uint32_t t_first = 0;
	if (!trace_decisions<CFG>(w, sc, args, smem_raw, rd, t_first)) return false;

	// phase 2: posterior of the path cell of every row, normalised by the row's recorded mass
	const Rec* recs = static_cast<const Rec*>(sc.recs);
	for (uint32_t r = t_first + lane; r < T; r += 32)
	{
		const uint32_t v = sc.pn[r];
		const uint32_t col = v & 0x7fffffffu;
		const bool isM = (v >> 31) != 0;
		const int q = (int)(col % SLOTS);
		const int ql = q / CPL, j = q - ql * CPL;
		const uint32_t r0 = sc.rowptr[r], r1 = sc.rowptr[r + 1];
		float mass = 0.0f, lp = 0.0f;
		bool found = false;
		for (uint32_t i = r0; i < r1; ++i)
		{
			const float* f = recs[i].v;
			for (int c = 0; c < 2 * CPL; ++c) mass += ex2(f[c]);
			if (__float_as_int(f[2 * CPL]) == ql)
			{
				found = true;
				lp = f[(isM ? 0 : CPL) + j];
			}
		}
		sc.pp[r] = (found && mass > 0.0f) ? exp2f(lp - log2f(mass)) : 0.0f;
	}
	__threadfence_block();
	__syncwarp();
	segment_medians<CFG>(w, sc, args, rd);
	return true;
}

// phase 3 of pass 3: per-segment median of the path posteriors sc.pp (NT:418-422, aligner.cpp:247-263)
DYN_DEV void segment_medians_impl(int lane, uint32_t T, uint32_t N, const SlotScratch& sc, const BatchArgs& args, const ReadDesc& rd)
{
	constexpr int NV = 4;           // register path: dwells up to 32 * NV samples
	constexpr float PAD = 3.0e38f;  // above every posterior
	const uint32_t* border = args.out_sigpos + rd.out_off;
	double* prob = args.out_prob + rd.out_off;
	const uint32_t Kc = N - 1;
	// segment s covers rows border[s] + 1 .. ext(s + 1), ext(s) = border[s] for s < Kc, T - 1 for s = Kc.
	// Borders are fetched 32 segments at a time and the path posteriors of the NEXT segment are requested before the
	// current one is ranked, so the dependent loads (border -> posteriors) are off the critical path.
	for (uint32_t s0 = 0; s0 < Kc; s0 += 32)
	{
		const uint32_t mine = (s0 + lane < Kc) ? border[s0 + lane] : (T - 1);
		const uint32_t last = (s0 + 32 < Kc) ? border[s0 + 32] : (T - 1);
		const uint32_t ns = min(32u, Kc - s0);
		// ext(s0 + j), j = 0 .. 32
		auto ext_at = [&](uint32_t j) -> uint32_t {
			const uint32_t e = __shfl_sync(FULL, mine, (int)(j & 31u));
			return (j < 32u) ? e : last;
		};
		float cur[NV], nxt[NV];
		uint32_t c_rs = ext_at(0) + 1, c_re = ext_at(1);
#pragma unroll
		for (int q = 0; q < NV; ++q)
		{
			const uint32_t idx = c_rs + q * 32 + lane;
			cur[q] = (idx <= c_re) ? sc.pp[idx] : PAD;
		}
		for (uint32_t j = 0; j < ns; ++j)
		{
			uint32_t n_rs = 0, n_re = 0;
#pragma unroll
			for (int q = 0; q < NV; ++q) nxt[q] = PAD;
			if (j + 1 < ns)
			{
				n_rs = ext_at(j + 1) + 1;
				n_re = ext_at(j + 2);
#pragma unroll
				for (int q = 0; q < NV; ++q)
				{
					const uint32_t idx = n_rs + q * 32 + lane;
					nxt[q] = (idx <= n_re) ? sc.pp[idx] : PAD;
				}
			}
			const uint32_t d = c_re - c_rs + 1;
			float up, dn;
			if (d <= 32) rank_select(cur[0], d, d / 2, lane, up, dn);
			else if (d <= 32 * NV) reg_select<NV>(cur, d / 2, lane, up, dn);
			else
			{
				const float* v = sc.pp + c_rs;
				up = coop_select(v, d, d / 2, lane);
				dn = (d & 1u) ? up : coop_select(v, d, d / 2 - 1, lane);
			}
			if (lane == 0) prob[s0 + j] = (d & 1u) ? (double)up : ((double)dn + (double)up) / 2.0;
			c_rs = n_rs;
			c_re = n_re;
#pragma unroll
			for (int q = 0; q < NV; ++q) cur[q] = nxt[q];
		}
	}
}

// ------------------------------------------------------------------------------------------------------
// training statistics (NT_aligner_api.cpp:494-514 and :641-725) from the sparse posterior records
// ------------------------------------------------------------------------------------------------------
// gamma(t,n) = pM + pE is accumulated per lattice column (one kmer per column) into read_w/x/xx; a later
// kernel folds columns into kmers.  Expected transition counts follow from the state posteriors:
//   #(E->M) = sum pM(t,n)            (every M(t+1,n) is entered from E(t,n-1))
//   #(E->E) = sum pE(t,n) - sum pM   (every E is entered from E or from the M directly before it)
template <class CFG>
DYN_DEV void train_stats_pass(Warp<CFG>& w, const SlotScratch& sc, const BatchArgs& args, const ReadDesc& rd,
	double& xi_m, double& xi_e)
{
	constexpr int CPL = CFG::CPL;
	typedef LaneRec<CPL> Rec;
	const Rec* recs = static_cast<const Rec*>(sc.recs);
	const int lane = w.lane;
	double sm_ = 0.0, se_ = 0.0;
	for (uint32_t r = 1 + lane; r < w.T; r += 32)
	{
		const uint32_t r0 = sc.rowptr[r], r1 = sc.rowptr[r + 1];
		double mass = 0.0;
		for (uint32_t i = r0; i < r1; ++i)
		{
			const float* f = recs[i].v;
			for (int c = 0; c < 2 * CPL; ++c) mass += (double)exp2f(f[c]);
		}
		if (!(mass > 0.0)) continue;
		const double inv = 1.0 / mass;
		const double xo = (double)w.sig[r - 1];
		const int n0 = (int)band_mid(r, w.ratio) - w.bw;
		for (uint32_t i = r0; i < r1; ++i)
		{
			const float* f = recs[i].v;
			const int rl = __float_as_int(f[2 * CPL]);
			for (int j = 0; j < CPL; ++j)
			{
				const double pm = (double)exp2f(f[j]) * inv, pe = (double)exp2f(f[CPL + j]) * inv;
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
}

// ------------------------------------------------------------------------------------------------------
// one read, all passes
// ------------------------------------------------------------------------------------------------------
template <class CFG, int MODE>
DYN_DEV void align_read(const BatchArgs& args, const ReadDesc& rd, uint32_t ridx, const SlotScratch& sc,
	unsigned char* smem_raw, int lane)
{
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

	ReadOut out;
	out.Z = 0.0;
	out.dZ = 0.0;
	out.nrec = 0;
	out.status = ST_OK;
	out.xi_m = 0.0;
	out.xi_e = 0.0;

	const double Z2 = (MODE == 0) ? backward_pass<CFG, false>(w, sc) : backward_pass<CFG, true>(w, sc);
	out.Z = Z2 * LN2;
	if (!(Z2 > (double)DEADT))
	{
		out.status = (MODE == 2) ? ST_TRAIN_FAILED : ST_ALIGN_FAILED;  // Zb is -inf (NT:289)
	}
	else if (MODE != 0)
	{
		uint32_t nrec = 0;
		bool overflow = false;
		const float dz2 = forward_posterior_pass<CFG>(w, sc, args, smem_raw, Z2, nrec, overflow);
		out.nrec = nrec;
		out.dZ = (double)dz2 * LN2;
		// the reference's consistency check (NT:288-291): |Zf - Zb| / (T*B) > 1e-8, B = 2*bw + 3.
		// In the reference's double arithmetic Zf and Zb agree to ~1e-9, so that test only ever fires when a
		// score is -inf (the band cut every path).  FP32 state carries |Zf - Zb| ~ 1e-4 .. 1e-3, which a narrow
		// band (small B) would trip, so an FP32 rounding allowance is added to the reference's tolerance.
		const double cells = (double)w.T * (double)(2 * w.bw + 3);
		if (!(dz2 > DEADT) || fabs(out.dZ) > 1e-8 * cells + 1e-6 * fabs(out.Z) + 1e-3)
			out.status = (MODE == 2) ? ST_TRAIN_FAILED : ST_ALIGN_FAILED;
		else if (overflow)
			out.status = ST_REC_OVERFLOW;
		else if (MODE == 1)
		{
			if (!traceback_pass<CFG>(w, sc, args, smem_raw, rd)) out.status = ST_INTERNAL;
		}
		else
		{
			__threadfence_block();
			__syncwarp();
			train_stats_pass<CFG>(w, sc, args, rd, out.xi_m, out.xi_e);
		}
	}
	if (lane == 0) args.out[ridx] = out;
}

} // namespace dyn
