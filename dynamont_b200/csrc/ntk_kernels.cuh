// Resquiggle ("NTK") mode, pre-pass stages: the dense sequence x signal (TN) and
// kmer x signal (TK) two-state HMMs, their posterior-mass row masks and the sorted key list of the sparse
// 5-state lattice (reference NTK_aligner_api.cpp:120-441).
//
// The dense pre-passes themselves live in ntk_prepass.cuh (FP64 linear domain with per-row scaling).  This file holds the
// shared helpers, the row-mask selection, the key list and the sparse 5-state stages.  The row masks are what the
// sparse stages consume, so they are the parity surface: membership is decided by a descending stable selection with a
// sequential log-sum-exp, with the semantics of columnArgsort + the insertion loop of preProcTN / preProcTK (:315-400).
#pragma once

#include "dp_common.cuh"

#ifndef DYN_HOST_EMU

namespace dyn
{
namespace ntk
{

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000ULL); }

// Aligner::logPlus (aligner.cpp:276-285)
__device__ __forceinline__ double log_plus(double x, double y)
{
	if (isinf(x)) return y;
	if (isinf(y)) return x;
	if (x < y)
	{
		const double t = x;
		x = y;
		y = t;
	}
	return x + log1p(exp(y - x));
}

// per-kmer emission constants prepared on the host with the reference's own libm calls (aligner.cpp:287-292)
struct KmerModel
{
	double mean, stdev, log_stdev;
};

struct Consts
{
	const KmerModel* model;  // [K]
	double half_log_2pi;     // 0.5 * log(2 * pi)
	double m, e;             // log ntMatch, log ntExtend (NTK:95-98)
};

// log_normal_pdf (aligner.cpp:287-292): -0.5 * z * z - log(stdev) - 0.5 * log(2 pi), evaluated left to right
__device__ __forceinline__ double score_kmer(const Consts& c, double x, uint32_t kmer)
{
	const KmerModel km = c.model[kmer];
	const double diff = x - km.mean;
	const double z = diff / km.stdev;
	return -0.5 * z * z - km.log_stdev - c.half_log_2pi;
}

// Row masks (NTK:343-353, 389-399): take the columns of row t in order of LP descending (stable: ties -> smaller
// index first, -inf last) until the sequential log-sum-exp of the taken values reaches SPARSETHRESHOLD (NTK:17: -0.02227639471 = log10(0.95), a mass of 0.978).
// One CTA per row.
//   compact path (the normal case): the columns within 40 nats of the row maximum — everything else adds < 1e-12 of
//   the mass of the largest column, and the selection stops at 0.978 of the row's mass — are gathered into shared memory
//   (<= RM_CAP of them), sorted once by (value descending, index ascending) with a bitonic network, and one thread runs
//   the reference's sequential log-sum-exp over the sorted list.  A 9-mer row (262 144 columns) is read twice instead of
//   once per selected column.
//   rounds path (fallback: more than RM_CAP candidates, an all -inf row, or the candidates do not reach the threshold):
//   every round finds the largest remaining value of the whole row with a block-wide arg-max.
constexpr int RM_CAP = 2048;

__device__ __forceinline__ bool rm_before(double va, uint32_t ia, double vb, uint32_t ib) { return va > vb || (va == vb && ia < ib); }

// rows_todo (may be null): only the rows with rows_todo[t] != 0 (the wide-row path's leftovers)
__global__ void __launch_bounds__(256) k_row_mask(const double* LP, uint32_t C, uint32_t words, uint32_t* mask, double threshold,
	const uint32_t* rows_todo = nullptr)
{
	__shared__ double s_val[RM_CAP];
	__shared__ uint32_t s_idx[RM_CAP];
	__shared__ double s_red[8];
	__shared__ int s_n, s_stop;
	const uint32_t t = blockIdx.x;
	if (rows_todo && rows_todo[t] == 0u) return;
	const double* row = LP + (size_t)t * C;
	uint32_t* m = mask + (size_t)t * words;
	for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) m[i] = 0u;
	if (threadIdx.x == 0) s_n = 0;
	// row maximum
	double mx = neg_inf();
	for (uint32_t i = threadIdx.x; i < C; i += blockDim.x) mx = fmax(mx, row[i]);
	for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
	if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
	__syncthreads();
	mx = s_red[0];
	for (int w = 1; w < 8; ++w) mx = fmax(mx, s_red[w]);
	bool compact = !isinf(mx);
	if (compact)
	{
		const double cut = mx - 40.0;
		for (uint32_t i = threadIdx.x; i < C; i += blockDim.x)
		{
			const double v = row[i];
			if (v >= cut)
			{
				const int pos = atomicAdd(&s_n, 1);
				if (pos < RM_CAP)
				{
					s_val[pos] = v;
					s_idx[pos] = i;
				}
			}
		}
		__syncthreads();
		compact = s_n <= RM_CAP;
	}
	if (compact)
	{
		const int n = s_n;
		int np2 = 1;
		while (np2 < n) np2 <<= 1;
		for (int i = n + threadIdx.x; i < np2; i += blockDim.x)
		{
			s_val[i] = neg_inf();
			s_idx[i] = 0xffffffffu;  // padding sorts last
		}
		__syncthreads();
		// bitonic sort, "before" = larger value, then smaller index
		for (int k = 2; k <= np2; k <<= 1)
			for (int j = k >> 1; j > 0; j >>= 1)
			{
				for (int i = threadIdx.x; i < np2; i += blockDim.x)
				{
					const int l = i ^ j;
					if (l > i)
					{
						const bool up = (i & k) == 0;
						const double va = s_val[i], vb = s_val[l];
						const uint32_t ia = s_idx[i], ib = s_idx[l];
						const bool swap = up ? rm_before(vb, ib, va, ia) : rm_before(va, ia, vb, ib);
						if (swap)
						{
							s_val[i] = vb; s_idx[i] = ib;
							s_val[l] = va; s_idx[l] = ia;
						}
					}
				}
				__syncthreads();
			}
		if (threadIdx.x == 0)
		{
			double sum = neg_inf();
			int reached = 0;
			for (int k = 0; k < n; ++k)
			{
				const uint32_t i = s_idx[k];
				m[i >> 5] |= 1u << (i & 31);
				sum = log_plus(sum, s_val[k]);
				if (sum >= threshold) { reached = 1; break; }
			}
			if (!reached && n < (int)C)
			{
				// the candidates do not carry the threshold mass: start over with the rounds below
				for (int k = 0; k < n; ++k) m[s_idx[k] >> 5] = 0u;
			}
			s_stop = reached || n >= (int)C;
		}
		__syncthreads();
		if (s_stop) return;
		__syncthreads();
	}
	// ---- rounds path
	double* r_val = s_val;      // [256]
	uint32_t* r_idx = s_idx;    // [256]
	double sum = neg_inf();
	uint32_t taken = 0;
	while (true)
	{
		// arg-max over the remaining columns; among equal values the smallest index wins; -inf is a legal value
		double bv = 0.0;
		uint32_t bi = 0xffffffffu;
		for (uint32_t i = threadIdx.x; i < C; i += blockDim.x)
		{
			if ((m[i >> 5] >> (i & 31)) & 1u) continue;
			const double v = row[i];
			if (bi == 0xffffffffu || v > bv) { bv = v; bi = i; }
		}
		r_val[threadIdx.x] = bv;
		r_idx[threadIdx.x] = bi;
		__syncthreads();
		for (int o = 128; o; o >>= 1)
		{
			if ((int)threadIdx.x < o)
			{
				const uint32_t oi = r_idx[threadIdx.x + o];
				const double ov = r_val[threadIdx.x + o];
				const uint32_t mi = r_idx[threadIdx.x];
				const double mv = r_val[threadIdx.x];
				if (oi != 0xffffffffu && (mi == 0xffffffffu || ov > mv || (ov == mv && oi < mi)))
				{
					r_val[threadIdx.x] = ov;
					r_idx[threadIdx.x] = oi;
				}
			}
			__syncthreads();
		}
		if (threadIdx.x == 0)
		{
			const uint32_t i = r_idx[0];
			m[i >> 5] |= 1u << (i & 31);
			sum = log_plus(sum, r_val[0]);
			++taken;
			s_stop = (sum >= threshold || taken >= C) ? 1 : 0;
		}
		__syncthreads();
		if (s_stop) break;
		__syncthreads();
	}
}

// ---- wide rows (9-mers: 262 144 columns): the same selection with the whole GPU on the two sweeps over the row ---------
// One CTA per row leaves a 2 MB row to 256 threads (264 ms for the 634 rows of a 60-base read); here the row maximum and
// the candidate compaction run over (row, chunk) CTAs, and only the sort + sequential log-sum-exp of the candidates
// (<= cap per row, RM_WIDE_CAP by default) is one CTA per row.  Rows the compact path cannot decide (more candidates than
// that, all -inf, threshold not reached) are flagged and go through k_row_mask (rows_todo).
constexpr uint32_t RM_CHUNK = 8192;
constexpr uint32_t RM_WIDE_CAP = 65536;

__device__ __forceinline__ unsigned long long rm_key(double v)
{
	const unsigned long long b = (unsigned long long)__double_as_longlong(v);
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);  // monotone in v (-inf lowest)
}
__device__ __forceinline__ double rm_unkey(unsigned long long k)
{
	return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}

struct WideMaskArgs
{
	const double* LP;
	uint32_t C, T, words;
	uint32_t* mask;               // [T][words], zeroed by the host
	double threshold;
	unsigned long long* rowmax;   // [T] rm_key of the row maximum, zeroed by the host
	uint32_t* cnt;                // [T] candidates found, zeroed by the host
	uint32_t cap;                 // candidates kept per row (a power of two)
	double* cval;                 // [T][cap]
	uint32_t* cidx;               // [T][cap]
	uint32_t* todo;               // [T] 1: the row needs k_row_mask
};

__global__ void __launch_bounds__(256) k_wide_rowmax(WideMaskArgs a)
{
	__shared__ double s_red[8];
	const uint32_t t = blockIdx.y, c0 = blockIdx.x * RM_CHUNK, c1 = min(c0 + RM_CHUNK, a.C);
	const double* row = a.LP + (size_t)t * a.C;
	double mx = neg_inf();
	for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) mx = fmax(mx, row[i]);
	for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
	if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		for (int w = 1; w < 8; ++w) mx = fmax(mx, s_red[w]);
		atomicMax(&a.rowmax[t], rm_key(mx));
	}
}

__global__ void __launch_bounds__(256) k_wide_candidates(WideMaskArgs a)
{
	const uint32_t t = blockIdx.y, c0 = blockIdx.x * RM_CHUNK, c1 = min(c0 + RM_CHUNK, a.C);
	const double* row = a.LP + (size_t)t * a.C;
	const double mx = rm_unkey(a.rowmax[t]);
	if (isinf(mx) || isnan(mx)) return;
	const double cut = mx - 40.0;
	for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x)
	{
		const double v = row[i];
		if (v >= cut)
		{
			const uint32_t pos = atomicAdd(&a.cnt[t], 1u);
			if (pos < a.cap)
			{
				a.cval[(size_t)t * a.cap + pos] = v;
				a.cidx[(size_t)t * a.cap + pos] = i;
			}
		}
	}
}

// one CTA per row: sort the row's candidates by (value descending, index ascending) — in shared memory when there are
// <= RM_CAP of them, in place in their (L2-resident) global list otherwise — then the reference's sequential log-sum-exp
__global__ void __launch_bounds__(1024) k_wide_select(WideMaskArgs a)
{
	__shared__ double s_val[RM_CAP];
	__shared__ uint32_t s_idx[RM_CAP];
	const uint32_t t = blockIdx.x;
	const uint32_t n = a.cnt[t];
	if (n == 0u || n > a.cap)
	{
		if (threadIdx.x == 0) a.todo[t] = 1u;
		return;
	}
	uint32_t np2 = 1;
	while (np2 < n) np2 <<= 1;
	double* val = a.cval + (size_t)t * a.cap;
	uint32_t* idx = a.cidx + (size_t)t * a.cap;
	if (np2 <= (uint32_t)RM_CAP)
	{
		for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x)
		{
			s_val[i] = (i < n) ? val[i] : neg_inf();
			s_idx[i] = (i < n) ? idx[i] : 0xffffffffu;  // padding sorts last
		}
		val = s_val;
		idx = s_idx;
	}
	else
	{
		// (cap is a power of two >= np2: the padding fits)
		for (uint32_t i = n + threadIdx.x; i < np2; i += blockDim.x)
		{
			val[i] = neg_inf();
			idx[i] = 0xffffffffu;
		}
	}
	__syncthreads();
	// bitonic sort, "before" = larger value, then smaller index (the order the atomics filled the list in does not matter)
	for (uint32_t k = 2; k <= np2; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1)
		{
			for (uint32_t i = threadIdx.x; i < np2; i += blockDim.x)
			{
				const uint32_t l = i ^ j;
				if (l > i)
				{
					const bool up = (i & k) == 0;
					const double va = val[i], vb = val[l];
					const uint32_t ia = idx[i], ib = idx[l];
					const bool swap = up ? rm_before(vb, ib, va, ia) : rm_before(va, ia, vb, ib);
					if (swap)
					{
						val[i] = vb; idx[i] = ib;
						val[l] = va; idx[l] = ia;
					}
				}
			}
			__syncthreads();
		}
	if (threadIdx.x == 0)
	{
		uint32_t* m = a.mask + (size_t)t * a.words;
		double sum = neg_inf();
		int reached = 0;
		for (uint32_t k = 0; k < n; ++k)
		{
			const uint32_t i = idx[k];
			m[i >> 5] |= 1u << (i & 31);
			sum = log_plus(sum, val[k]);
			if (sum >= a.threshold) { reached = 1; break; }
		}
		if (!reached && n < a.C)
		{
			// the candidates do not carry the threshold mass: k_row_mask starts over (it zeroes the row's mask itself)
			a.todo[t] = 1u;
		}
	}
}

// ---- sparse-lattice keys (preProcTNK, NTK:402-441): key(t,n,q) = t*N*K + n*K + q ----------------------------
// for n in tnMap[t] (n >= 1): kmer[n-1] and every q in tkMap[t]; key 0 iff 0 in tnMap[0].  Emitted in ascending order.
struct KeyArgs
{
	const uint32_t* tn;  // [T][wn]
	const uint32_t* tk;  // [T][wk]
	const int32_t* kmers;
	uint32_t T, N, K, wn, wk;
	uint64_t* count;     // [T]   (pass 0)   /  exclusive offsets (pass 1)
	uint64_t* keys;
};

template <bool FILL>
__global__ void k_keys(KeyArgs a)
{
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < a.T; t += gridDim.x * blockDim.x)
	{
		const uint32_t* tn = a.tn + (size_t)t * a.wn;
		const uint32_t* tk = a.tk + (size_t)t * a.wk;
		uint64_t pos = FILL ? a.count[t] : 0;
		const uint64_t tNK = (uint64_t)t * a.N * a.K;
		uint32_t ntk = 0;
		if (!FILL)
			for (uint32_t w = 0; w < a.wk; ++w) ntk += __popc(tk[w]);
		for (uint32_t wi = 0; wi < a.wn; ++wi)
		{
			uint32_t bits = tn[wi];
			while (bits)
			{
				const uint32_t n = wi * 32 + (__ffs(bits) - 1);
				bits &= bits - 1;
				if (n == 0)
				{
					if (t == 0)
					{
						if (FILL) a.keys[pos] = 0;
						++pos;
					}
					continue;
				}
				const uint32_t kn = (uint32_t)a.kmers[n - 1];
				const bool in_tk = (tk[kn >> 5] >> (kn & 31)) & 1u;
				if (!FILL)
				{
					pos += ntk + (in_tk ? 0 : 1);
					continue;
				}
				const uint64_t base = tNK + (uint64_t)n * a.K;
				bool put = in_tk;  // kmer[n-1] merged into the ascending scan of tkMap[t]
				for (uint32_t wk = 0; wk < a.wk; ++wk)
				{
					uint32_t kb = tk[wk];
					while (kb)
					{
						const uint32_t q = wk * 32 + (__ffs(kb) - 1);
						kb &= kb - 1;
						if (!put && kn < q)
						{
							a.keys[pos++] = base + kn;
							put = true;
						}
						a.keys[pos++] = base + q;
					}
				}
				if (!put) a.keys[pos++] = base + kn;
			}
		}
		if (!FILL) a.count[t] = pos;
	}
}

// ------------------------------------------------------------------------------------------------------
// sparse 5-state stages (rows B7-B10): logF / logB (NTK:443-607, with each result visible to later keys at once —
// the repair of SURVEY.md F2), Z check (:897-918), sparse logP (:159-177), MAP fill + traceback (:628-879).
// States A=0 P=1 S=2 E=3 I=4.  One warp per read: the keys of a lattice row are independent except for the I state
// (same row, column n-1 / n+1), which one lane then completes in key order; row t only needs rows t-1 / t+1.
// A key is found by binary search inside its row's slice of the sorted key list.
// ------------------------------------------------------------------------------------------------------
struct SparseArgs
{
	const double* signal;     // [S]
	const int32_t* kmers;     // [N-1]
	const uint64_t* keys;     // [nk] ascending
	const uint64_t* rowptr;   // [T+1] first key of row t
	uint64_t nk;
	uint32_t T, N, K, hp, k;
	Consts c;
	double tr[14];            // log a1,a2,p1,p2,p3,s1,s2,s3,e1,e2,e3,e4,i1,i2
	double *F, *B, *LP, *V;   // [nk][5]
	double* out_z;            // [2] Zf, Zb
	int32_t* out_status;      // 0 ok, 1 scores do not match, 2 traceback stuck (undefined in the reference)
	// segments, in traceback order (the host reverses): state ('M'/'P'), n-1+k/2, t-1, median probability, polish kmer id
	uint32_t* seg_n;
	char* seg_state;
	uint64_t* seg_seqpos;
	uint64_t* seg_sigpos;
	double* seg_prob;
	uint32_t* seg_kmer;
	double* prob_buf;         // [T + N] scratch for the probabilities of one segment
	int calc_prob;
};

enum { TA1 = 0, TA2, TP1, TP2, TP3, TS1, TS2, TS3, TE1, TE2, TE3, TE4, TI1, TI2 };

// The keys of up to three lattice rows, as 32-bit in-row codes n*K + q, staged in shared memory: a predecessor /
// successor lookup is then a binary search at shared-memory latency.  Rows with more than ROWCACHE keys (rows whose
// posterior mass is spread over many columns) fall back to the global key list.
constexpr uint32_t ROWCACHE = 1536;
struct RowCache
{
	uint32_t* code[3];  // shared memory
	uint32_t row[3];    // lattice row held by slot i (0xffffffff: none)
	uint32_t cnt[3];
};

__device__ __forceinline__ void rc_load(const SparseArgs& a, RowCache& rc, int slot, uint32_t t, int lane)
{
	rc.row[slot] = 0xffffffffu;
	if (t >= a.T) return;
	const uint64_t r0 = a.rowptr[t], r1 = a.rowptr[t + 1];
	if (r1 - r0 > ROWCACHE || (uint64_t)a.N * a.K > 0xffffffffull) return;
	const uint64_t tNK = (uint64_t)t * a.N * a.K;
	for (uint64_t i = r0 + lane; i < r1; i += 32) rc.code[slot][i - r0] = (uint32_t)(a.keys[i] - tNK);
	rc.row[slot] = t;
	rc.cnt[slot] = (uint32_t)(r1 - r0);
}

__device__ __forceinline__ int64_t sp_find(const SparseArgs& a, const RowCache& rc, uint32_t t, uint32_t n, uint32_t q)
{
	for (int sl = 0; sl < 3; ++sl)
		if (rc.row[sl] == t)
		{
			const uint32_t code = n * a.K + q;
			const uint32_t* c = rc.code[sl];
			uint32_t lo = 0, hi = rc.cnt[sl];
			while (lo < hi)
			{
				const uint32_t mid = (lo + hi) >> 1;
				if (c[mid] < code) lo = mid + 1;
				else hi = mid;
			}
			return (lo < rc.cnt[sl] && c[lo] == code) ? (int64_t)(a.rowptr[t] + lo) : -1;
		}
	const uint64_t key = ((uint64_t)t * a.N + n) * a.K + q;
	uint64_t lo = a.rowptr[t], hi = a.rowptr[t + 1];
	while (lo < hi)
	{
		const uint64_t mid = (lo + hi) >> 1;
		const uint64_t v = a.keys[mid];
		if (v < key) lo = mid + 1;
		else hi = mid;
	}
	return (lo < a.rowptr[t + 1] && a.keys[lo] == key) ? (int64_t)lo : -1;
}

__device__ __forceinline__ double sp_get(const double* M, int64_t idx, int st) { return idx < 0 ? neg_inf() : M[idx * 5 + st]; }

// scoreHD / score (NTK:120-140)
__device__ __forceinline__ double sp_score(const SparseArgs& a, double x, uint32_t kn, uint32_t kk)
{
	int dist = 0;
	if (kn != kk)
	{
		uint32_t u = kn, v = kk;
		for (uint32_t i = 0; i < a.k; ++i)
		{
			dist += ((u & 3u) != (v & 3u));
			u >>= 2;
			v >>= 2;
		}
	}
	return score_kmer(a.c, x, kn) + score_kmer(a.c, x, kk) + (double)(-2 * dist);
}

__device__ __forceinline__ void sp_decode(const SparseArgs& a, uint64_t key, uint32_t& t, uint32_t& n, uint32_t& q)
{
	const uint64_t NK = (uint64_t)a.N * a.K;
	t = (uint32_t)(key / NK);
	const uint64_t r = key % NK;
	n = (uint32_t)(r / a.K);
	q = (uint32_t)(r % a.K);
}

// one pass of logF (MAXP = false) or of the MAP fill of decodeMAP (MAXP = true, NTK:814-863) over all rows
template <bool MAXP>
__device__ void sp_forward(const SparseArgs& a, RowCache& rc, int lane)
{
	const double NI = neg_inf();
	double* M = MAXP ? a.V : a.F;
	rc.row[0] = rc.row[1] = rc.row[2] = 0xffffffffu;
	for (uint32_t t = 0; t < a.T; ++t)
	{
		// slot t&1: this row, slot (t+1)&1: the previous one (loaded when it was "this row")
		__syncwarp();
		rc_load(a, rc, t & 1, t, lane);
		__syncwarp();
		const uint64_t r0 = a.rowptr[t], r1 = a.rowptr[t + 1];
		for (uint64_t idx = r0 + lane; idx < r1; idx += 32)
		{
			uint32_t tt, n, q;
			sp_decode(a, a.keys[idx], tt, n, q);
			double va = NI, vp = NI, vs = NI, ve = NI;
			if (t == 0 && n == 0) ve = 0.0;
			else if (t > 0 && n > 0)
			{
				const double sc = MAXP ? 0.0 : sp_score(a, a.signal[t - 1], (uint32_t)a.kmers[n - 1], q);
				const double* lp = a.LP + idx * 5;
				for (uint32_t pre = q / 4; pre < a.K; pre += a.hp)
				{
					const int64_t ia = sp_find(a, rc, t - 1, n - 1, pre), ip = sp_find(a, rc, t - 1, n, pre);
					if (MAXP)
					{
						va = fmax(va, sp_get(M, ia, 3) + lp[0]);
						va = fmax(va, sp_get(M, ia, 4) + lp[0]);
						vp = fmax(vp, sp_get(M, ip, 2) + lp[1]);
						vp = fmax(vp, sp_get(M, ip, 3) + lp[1]);
						vp = fmax(vp, sp_get(M, ip, 4) + lp[1]);
					}
					else
					{
						va = log_plus(va, sp_get(M, ia, 3) + a.tr[TA1] + sc);
						va = log_plus(va, sp_get(M, ia, 4) + a.tr[TA2] + sc);
						vp = log_plus(vp, sp_get(M, ip, 2) + a.tr[TP1] + sc);
						vp = log_plus(vp, sp_get(M, ip, 3) + a.tr[TP2] + sc);
						vp = log_plus(vp, sp_get(M, ip, 4) + a.tr[TP3] + sc);
					}
				}
				const int64_t is = sp_find(a, rc, t - 1, n - 1, q), ie = sp_find(a, rc, t - 1, n, q);
				if (MAXP)
				{
					vs = fmax(vs, sp_get(M, is, 1) + lp[2]);
					vs = fmax(vs, sp_get(M, is, 3) + lp[2]);
					vs = fmax(vs, sp_get(M, is, 4) + lp[2]);
					ve = fmax(ve, sp_get(M, ie, 0) + lp[3]);
					ve = fmax(ve, sp_get(M, ie, 1) + lp[3]);
					ve = fmax(ve, sp_get(M, ie, 2) + lp[3]);
					ve = fmax(ve, sp_get(M, ie, 3) + lp[3]);
				}
				else
				{
					vs = log_plus(vs, sp_get(M, is, 1) + a.tr[TS1] + sc);
					vs = log_plus(vs, sp_get(M, is, 3) + a.tr[TS2] + sc);
					vs = log_plus(vs, sp_get(M, is, 4) + a.tr[TS3] + sc);
					ve = log_plus(ve, sp_get(M, ie, 0) + sc);
					ve = log_plus(ve, sp_get(M, ie, 1) + a.tr[TE2] + sc);
					ve = log_plus(ve, sp_get(M, ie, 2) + a.tr[TE3] + sc);
					ve = log_plus(ve, sp_get(M, ie, 3) + a.tr[TE4] + sc);
				}
			}
			double* o = M + idx * 5;
			o[0] = va; o[1] = vp; o[2] = vs; o[3] = ve; o[4] = NI;
		}
		__syncwarp();
		// The I state consumes no sample: (t, n-1, q) of the same row, an earlier key (NTK:503-505, 857-858).  Every
		// lane walks the chain of its own key from the first column of the run (t, n0..n, q) upwards — the same
		// operations in the same order as the sequential evaluation, so the values are identical.
		if (t > 0)
		{
			for (uint64_t idx = r0 + lane; idx < r1; idx += 32)
			{
				uint32_t tt, n, q;
				sp_decode(a, a.keys[idx], tt, n, q);
				if (n == 0) continue;
				uint32_t n0 = n;
				while (n0 > 1 && sp_find(a, rc, t, n0 - 1, q) >= 0) --n0;
				double vi = NI;  // I of (t, n0-1, q): absent or column 0
				int64_t prev = (n0 >= 1) ? sp_find(a, rc, t, n0 - 1, q) : -1;
				for (uint32_t m = n0; m <= n; ++m)
				{
					const int64_t cur = (m == n) ? (int64_t)idx : sp_find(a, rc, t, m, q);
					const double pe = sp_get(M, prev, 3);
					const double pi = (m == n0) ? sp_get(M, prev, 4) : vi;
					double x = NI;
					if (MAXP)
					{
						const double l4 = a.LP[cur * 5 + 4];
						x = fmax(x, pe + l4);
						x = fmax(x, pi + l4);
					}
					else
					{
						const double sc = sp_score(a, a.signal[t - 1], (uint32_t)a.kmers[m - 1], q);
						x = log_plus(x, pe + a.tr[TI1] + sc);
						x = log_plus(x, pi + a.tr[TI2] + sc);
					}
					vi = x;
					prev = cur;
				}
				M[idx * 5 + 4] = vi;
			}
		}
	}
	__syncwarp();
}

// logB (NTK:517-607)
__device__ void sp_backward(const SparseArgs& a, RowCache& rc, int lane)
{
	const double NI = neg_inf();
	double* M = a.B;
	const uint32_t T = a.T, N = a.N;
	rc.row[0] = rc.row[1] = rc.row[2] = 0xffffffffu;
	for (uint32_t t = T; t-- > 0;)
	{
		__syncwarp();
		rc_load(a, rc, t & 1, t, lane);  // the other slot still holds row t+1
		__syncwarp();
		const uint64_t r0 = a.rowptr[t], r1 = a.rowptr[t + 1];
		for (uint64_t idx = r0 + lane; idx < r1; idx += 32)
		{
			uint32_t tt, n, q;
			sp_decode(a, a.keys[idx], tt, n, q);
			double va = NI, vp = NI, vs = NI, ve = NI, vi = NI;
			if (t == T - 1 && n == N - 1) ve = 0.0;
			if (t < T - 1)
			{
				const uint32_t s0 = (q % a.hp) * 4;
				const double x = a.signal[t];
				if (n > 0)
				{
					const uint32_t kn = (uint32_t)a.kmers[n - 1];
					double sc = sp_score(a, x, kn, q);
					const int64_t ie = sp_find(a, rc, t + 1, n, q);
					const double fe = sp_get(M, ie, 3);
					va = log_plus(va, fe + sc);
					vp = log_plus(vp, fe + a.tr[TE2] + sc);
					vs = log_plus(vs, fe + a.tr[TE3] + sc);
					ve = log_plus(ve, fe + a.tr[TE4] + sc);
					for (uint32_t suc = s0; suc < s0 + 4; ++suc)
					{
						sc = sp_score(a, x, kn, suc);
						const double fp = sp_get(M, sp_find(a, rc, t + 1, n, suc), 1);
						vs = log_plus(vs, fp + a.tr[TP1] + sc);
						ve = log_plus(ve, fp + a.tr[TP2] + sc);
						vi = log_plus(vi, fp + a.tr[TP3] + sc);
					}
				}
				if (n < N - 1)
				{
					const uint32_t kn = (uint32_t)a.kmers[n];
					double sc = sp_score(a, x, kn, q);
					const double fs = sp_get(M, sp_find(a, rc, t + 1, n + 1, q), 2);
					vp = log_plus(vp, fs + a.tr[TS1] + sc);
					ve = log_plus(ve, fs + a.tr[TS2] + sc);
					vi = log_plus(vi, fs + a.tr[TS3] + sc);
					for (uint32_t suc = s0; suc < s0 + 4; ++suc)
					{
						sc = sp_score(a, x, kn, suc);
						const double fa = sp_get(M, sp_find(a, rc, t + 1, n + 1, suc), 0);
						ve = log_plus(ve, fa + a.tr[TA1] + sc);
						vi = log_plus(vi, fa + a.tr[TA2] + sc);
					}
				}
			}
			double* o = M + idx * 5;
			o[0] = va; o[1] = vp; o[2] = vs; o[3] = ve; o[4] = vi;
			a.LP[idx * 5 + 2] = 0.0;  // LP doubles as scratch for the same-row update below
		}
		__syncwarp();
		// The I move into (t, n+1, q) of the same row, a later key (NTK:592-598): only the I component of the successor
		// enters, so every lane walks its own chain from the last column of the run (t, n..n1, q) downwards, adding
		// the same terms in the same order as the sequential (descending key order) evaluation.
		if (t > 0)
		{
			for (uint64_t idx = r0 + lane; idx < r1; idx += 32)
			{
				uint32_t tt, n, q;
				sp_decode(a, a.keys[idx], tt, n, q);
				if (n >= N - 1) continue;
				uint32_t n1 = n;
				while (n1 + 1 <= N - 1 && sp_find(a, rc, t, n1 + 1, q) >= 0) ++n1;
				if (n1 == n) continue;  // successor absent: the terms are -inf, log_plus leaves e and i unchanged
				// I of (t, m, q) for m = n1 down to n+1, each including its own same-row term
				double fi = 0.0;
				for (uint32_t m = n1; m > n; --m)
				{
					const int64_t cur = sp_find(a, rc, t, m, q);
					double im = M[cur * 5 + 4];  // phase-1 value (terms of row t+1 only)
					if (m < n1 && m < N - 1)
					{
						const double sc = sp_score(a, a.signal[t - 1], (uint32_t)a.kmers[m], q);
						im = log_plus(im, fi + a.tr[TI2] + sc);
					}
					fi = im;
				}
				const double sc = sp_score(a, a.signal[t - 1], (uint32_t)a.kmers[n], q);
				double* o = M + idx * 5;
				const double e_new = log_plus(o[3], fi + a.tr[TI1] + sc);
				const double i_new = log_plus(o[4], fi + a.tr[TI2] + sc);
				// written to separate outputs after all lanes have read the phase-1 I values
				a.LP[idx * 5 + 0] = e_new;
				a.LP[idx * 5 + 1] = i_new;
				a.LP[idx * 5 + 2] = 1.0;  // marker
			}
			__syncwarp();
			for (uint64_t idx = r0 + lane; idx < r1; idx += 32)
			{
				if (a.LP[idx * 5 + 2] == 1.0)
				{
					M[idx * 5 + 3] = a.LP[idx * 5 + 0];
					M[idx * 5 + 4] = a.LP[idx * 5 + 1];
					a.LP[idx * 5 + 2] = 0.0;
				}
			}
		}
	}
	__syncwarp();
}

// formattedMedian (aligner.cpp:247-263) of v[0..m); sorts v in place
__device__ double sp_median(double* v, uint32_t m)
{
	if (m == 0) return 0.0;
	for (uint32_t i = 1; i < m; ++i)
	{
		const double x = v[i];
		uint32_t j = i;
		while (j > 0 && v[j - 1] > x)
		{
			v[j] = v[j - 1];
			--j;
		}
		v[j] = x;
	}
	return (m & 1u) ? v[m / 2] : (v[m / 2 - 1] + v[m / 2]) / 2.0;
}

__device__ void sp_push_segment(const SparseArgs& a, uint32_t& ns, char state, uint64_t seqpos, uint64_t sigpos, double* prob,
	uint32_t& np, uint32_t q)
{
	a.seg_state[ns] = state;
	a.seg_seqpos[ns] = seqpos;
	a.seg_sigpos[ns] = sigpos;
	a.seg_prob[ns] = sp_median(prob, np);
	a.seg_kmer[ns] = q;
	++ns;
	np = 0;
}

// NTKAligner::traceback (NTK:628-801), one thread
__device__ int sp_traceback(const SparseArgs& a, uint32_t t, uint32_t n, uint32_t q)
{
	RowCache rc;
	rc.row[0] = rc.row[1] = rc.row[2] = 0xffffffffu;
	const double* V = a.V;
	const double* LP = a.LP;
	double* prob = a.prob_buf;
	uint32_t np = 0, ns = 0;
	int state = 3;
	int64_t cur = sp_find(a, rc, t, n, q);
	const uint32_t half = a.k / 2;
	uint64_t guard = (uint64_t)a.T + a.N + 8;
	while (t)
	{
		if (guard-- == 0) return 2;
		if (state == 3)
		{
			if (t == 1)
			{
				sp_push_segment(a, ns, 'M', half, 0, prob, np, q);
				break;
			}
			const int64_t prev = sp_find(a, rc, t - 1, n, q);
			const double sc = sp_get(V, cur, 3), ls = sp_get(LP, cur, 3);
			prob[np++] = exp(ls);
			if (sc == sp_get(V, prev, 3) + ls) state = 3;
			else if (sc == sp_get(V, prev, 0) + ls) state = 0;
			else if (sc == sp_get(V, prev, 2) + ls) state = 2;
			else if (sc == sp_get(V, prev, 1) + ls) state = 1;
			--t;
			cur = prev;
		}
		else if (state == 0 || state == 1)
		{
			const bool isA = (state == 0);
			if (t == 1 && (!isA || n == 1))
			{
				sp_push_segment(a, ns, isA ? 'M' : 'P', half, 0, prob, np, q);
				break;
			}
			const double sc = sp_get(V, cur, state), ls = sp_get(LP, cur, state);
			prob[np++] = exp(ls);
			bool moved = false;
			for (uint32_t pre = q / 4; pre < a.K && !moved; pre += a.hp)
			{
				const int64_t prev = sp_find(a, rc, t - 1, isA ? n - 1 : n, pre);
				int nstate = -1;
				if (sc == sp_get(V, prev, 3) + ls) nstate = 3;
				else if (!isA && sc == sp_get(V, prev, 2) + ls) nstate = 2;
				else if (sc == sp_get(V, prev, 4) + ls) nstate = 4;
				if (nstate >= 0)
				{
					sp_push_segment(a, ns, isA ? 'M' : 'P', (uint64_t)n - 1 + half, (uint64_t)t - 1, prob, np, q);
					state = nstate;
					--t;
					if (isA) --n;
					q = pre;
					cur = prev;
					moved = true;
				}
			}
			if (!moved) return 2;  // the reference would spin here
		}
		else if (state == 2)
		{
			if (t == 1 && n == 1) break;
			const int64_t prev = sp_find(a, rc, t - 1, n - 1, q);
			const double sc = sp_get(V, cur, 2), ls = sp_get(LP, cur, 2);
			prob[np++] = exp(ls);
			if (sc == sp_get(V, prev, 3) + ls) state = 3;
			else if (sc == sp_get(V, prev, 1) + ls) state = 1;
			else if (sc == sp_get(V, prev, 4) + ls) state = 4;
			--t;
			--n;
			cur = prev;
		}
		else
		{
			if (n == 1) break;
			const int64_t prev = sp_find(a, rc, t, n - 1, q);
			const double sc = sp_get(V, cur, 4), ls = sp_get(LP, cur, 4);
			prob[np++] = exp(ls);
			if (sc == sp_get(V, prev, 4) + ls) state = 4;
			else if (sc == sp_get(V, prev, 3) + ls) state = 3;
			--n;
			cur = prev;
		}
	}
	*a.seg_n = ns;
	return 0;
}

// logF and logB are independent: grid = 2 single-warp CTAs
__global__ void __launch_bounds__(32) k_ntk_sparse_fb(SparseArgs a)
{
	__shared__ uint32_t codes[2][ROWCACHE];
	RowCache rc;
	rc.code[0] = codes[0];
	rc.code[1] = codes[1];
	rc.code[2] = codes[0];
	if (blockIdx.x == 0) sp_forward<false>(a, rc, threadIdx.x);
	else sp_backward(a, rc, threadIdx.x);
}

__global__ void __launch_bounds__(32) k_ntk_sparse(SparseArgs a)
{
	const int lane = threadIdx.x;
	const double NI = neg_inf();
	__shared__ uint32_t codes[2][ROWCACHE];
	RowCache rc;
	rc.code[0] = codes[0];
	rc.code[1] = codes[1];
	rc.code[2] = codes[0];
	// Zf, Zb (NTK:897-918): sequential log-sum-exp over q of the E state of (T-1, N-1, q) / (0, 0, q)
	double Zf = NI, Zb = NI;
	if (lane == 0)
	{
		for (uint64_t idx = a.rowptr[a.T - 1]; idx < a.rowptr[a.T]; ++idx)
		{
			uint32_t t, n, q;
			sp_decode(a, a.keys[idx], t, n, q);
			if (n == a.N - 1) Zf = log_plus(Zf, a.F[idx * 5 + 3]);
		}
		for (uint64_t idx = a.rowptr[0]; idx < a.rowptr[1]; ++idx)
		{
			uint32_t t, n, q;
			sp_decode(a, a.keys[idx], t, n, q);
			if (n == 0) Zb = log_plus(Zb, a.B[idx * 5 + 3]);
		}
		a.out_z[0] = Zf;
		a.out_z[1] = Zb;
	}
	Zf = __shfl_sync(0xffffffffu, Zf, 0);
	Zb = __shfl_sync(0xffffffffu, Zb, 0);
	const double cells = (double)a.T * (double)a.N * (double)a.K;
	if (fabs(Zf - Zb) / cells >= 1e-8 || isinf(Zf) || isinf(Zb))
	{
		if (lane == 0) *a.out_status = 1;
		return;
	}
	if (lane == 0) *a.out_status = 0;
	if (!a.calc_prob) return;
	// sparse logP (NTK:159-177)
	for (uint64_t i = lane; i < a.nk * 5; i += 32) a.LP[i] = a.F[i] + a.B[i] - Zb;
	__syncwarp();
	sp_forward<true>(a, rc, lane);
	__syncwarp();
	if (lane == 0)
	{
		// end kmer: the LAST q whose score is >= the best so far (NTK:865-876); absent keys count as -inf
		double best = NI;
		uint32_t bestq = a.K - 1;
		bool any = false;
		for (uint64_t idx = a.rowptr[a.T - 1]; idx < a.rowptr[a.T]; ++idx)
		{
			uint32_t t, n, q;
			sp_decode(a, a.keys[idx], t, n, q);
			if (n != a.N - 1) continue;
			const double cand = a.V[idx * 5 + 3];
			if (cand >= best && !(isinf(cand) && cand < 0))
			{
				best = cand;
				bestq = q;
				any = true;
			}
		}
		if (!any) bestq = a.K - 1;
		const int rc = sp_traceback(a, a.T - 1, a.N - 1, bestq);
		if (rc) *a.out_status = rc;
	}
}

} // namespace ntk
} // namespace dyn

#endif // DYN_HOST_EMU
