// Resquiggle ("NTK") mode, the two dense pre-passes (reference NTK_aligner_api.cpp:197-400):
//   TN  the two-state HMM over (signal sample x sequence position), unbanded, T x N
//   TK  the two-state HMM over (signal sample x kmer) on the de-Bruijn graph of all K = 4^k kmers, T x K
// and the per-row log posteriors their top-mass row masks are selected from.
//
// Design (nothing here mirrors the reference's loops):
//   * LINEAR domain in FP64 with one power-of-two scale per lattice row.  The reference adds log-probabilities with
//     logPlus (an exp and a log1p per term, five of them per TK cell); here a cell-update is a handful of DMUL/DFMA and ONE
//     exp (the emission).  A row is scaled by the exponent of the PREVIOUS row's largest value (known when the row
//     starts: no second sweep), and the scales accumulate as integers, so Z and the posteriors keep double accuracy
//     (Z agrees with the reference to ~1e-12 relative; the masks are selected from the same numbers).
//   * The de-Bruijn structure is used instead of gathered: the four predecessors of kmer q are the same for q, q^1, q^2,
//     q^3 (q >> 2 fixed), so ONE sum S[r] = sum_j E[r + j*4^(k-1)] serves four cells of the forward row; the four
//     successors of q are the contiguous kmers 4*(q mod 4^(k-1)) .. +3 and are shared by the four kmers with the same
//     q mod 4^(k-1), so ONE sum R[r] serves four cells of the backward row.  A thread owns one r: 4 coalesced strided
//     reads + 4 contiguous cells forward, 4 contiguous reads + 4 strided cells backward.
//   * Small lattices (TN; TK for 5-mers: 1024 kmers) run as ONE CTA per direction that walks the rows with a block
//     barrier per row — reads are batched across CTAs / streams.  Large lattices (TK for 9-mers: 262 144 kmers = 2 MB per
//     row and state) run ONE LAUNCH PER ROW over the whole GPU (1024 CTAs), the rows staying L2-resident between
//     launches; the kernel boundary is the grid barrier.  That is what makes k = 9 feasible at T in the thousands.
//   * Only the forward lattice is stored (16*T*C bytes).  The backward pass keeps two rows and emits the log posterior
//     of every cell as it goes; Z is folded into the mask threshold afterwards (logsumexp(lp) >= threshold + Z).
#pragma once

#include "dp_common.cuh"

#ifndef DYN_HOST_EMU
#include "ntk_kernels.cuh"

namespace dyn
{
namespace ntk
{

struct PreArgs
{
	const double* signal;     // [S]
	const int32_t* kmers;     // [N-1] kmer id of column n is kmers[n-1] (TN)
	const KmerModel* model;   // [K]
	double half_log_2pi;
	double m, e;              // ntMatch, ntExtend as PROBABILITIES (NTK:95-98 takes their logs)
	uint32_t T, C, hp;        // C = N (TN) or K (TK); hp = 4^(k-1)
	int tk;                   // 0: TN, 1: TK
	double *fM, *fE;          // [T][C] forward lattice, row t scaled by 2^fexp[t]
	double* brow;             // [2][2][C] backward rows (ping-pong; M then E)
	double* LP;               // [T][C] log posterior + Z (natural log)
	unsigned long long* fmax; // [T] bit pattern of the largest value of forward row t (values are >= 0)
	unsigned long long* bmax; // [T]
	int* fexp;                // [T] true value = stored * 2^fexp[t]
	int* bexp;                // [T]
	double* z;                // [2] Zf, Zb (natural log)
};

__device__ __forceinline__ double emis_lin(const PreArgs& a, double x, uint32_t q)
{
	const KmerModel km = a.model[q];
	const double diff = x - km.mean;
	const double zz = diff / km.stdev;
	return exp(-0.5 * zz * zz - km.log_stdev - a.half_log_2pi);  // N(x; mu, sigma), aligner.cpp:287-292
}

// scale of row t from the largest value of its neighbour row: 2^(-exponent), so that the neighbour's maximum would be in [1, 2)
__device__ __forceinline__ int row_shift(unsigned long long maxbits)
{
	if (maxbits == 0ull) return 0;
	return 1023 - (int)((maxbits >> 52) & 0x7ffull);
}
__device__ __forceinline__ double pow2d(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }

// ---- one forward row t (t >= 1) from row t-1; cells idx, idx + stride, ... ------------------------------------------
__device__ __forceinline__ void fwd_row(const PreArgs& a, uint32_t t, uint32_t idx, uint32_t stride, unsigned long long& lmax)
{
	const uint32_t C = a.C;
	const double x = a.signal[t - 1];
	const double* pM = a.fM + (size_t)(t - 1) * C;
	const double* pE = a.fE + (size_t)(t - 1) * C;
	double* cM = a.fM + (size_t)t * C;
	double* cE = a.fE + (size_t)t * C;
	const int sh = row_shift(a.fmax[t - 1]);
	const double sc = pow2d(sh);
	if (!a.tk)
	{
		// TN (NTK:197-218): fM[t][n] = fE[t-1][n-1] * p * m; fE[t][n] = (fM[t-1][n] + fE[t-1][n] * e) * p, p = N(x[t-1]; kmer[n-1])
		for (uint32_t n = idx; n < C; n += stride)
		{
			double vm = 0.0, ve = 0.0;
			if (n >= 1)
			{
				const double p = emis_lin(a, x, (uint32_t)a.kmers[n - 1]) * sc;
				vm = pE[n - 1] * p * a.m;
				ve = (pM[n] + pE[n] * a.e) * p;
			}
			cM[n] = vm;
			cE[n] = ve;
			lmax = max(lmax, (unsigned long long)__double_as_longlong(fmax(vm, ve)));
		}
	}
	else
	{
		// TK (NTK:253-281): the four predecessors of kmers 4r .. 4r+3 are r, r+hp, r+2hp, r+3hp
		const uint32_t hp = a.hp;
		for (uint32_t r = idx; r < hp; r += stride)
		{
			const double S = ((pE[r] + pE[r + hp]) + (pE[r + 2 * hp] + pE[r + 3 * hp])) * a.m;
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j)
			{
				const uint32_t q = 4 * r + j;
				const double p = emis_lin(a, x, q) * sc;
				const double vm = S * p;
				const double ve = (pM[q] + pE[q] * a.e) * p;
				cM[q] = vm;
				cE[q] = ve;
				lmax = max(lmax, (unsigned long long)__double_as_longlong(fmax(vm, ve)));
			}
		}
	}
}

// ---- one backward row t (t <= T-2) from row t+1, and the log posteriors of row t -------------------------------------
// nM / nE: backward row t+1; cM / cE: row t.  LP[t][c] = log(fM*bM + fE*bE) + ln2 * (fexp[t] + bexp[t])   (unnormalised)
__device__ __forceinline__ void bwd_row(const PreArgs& a, uint32_t t, uint32_t idx, uint32_t stride, const double* nM, const double* nE,
	double* cM, double* cE, int bexp_t, double sc, unsigned long long& lmax)
{
	const uint32_t C = a.C;
	const double x = a.signal[t];
	const double* fM = a.fM + (size_t)t * C;
	const double* fE = a.fE + (size_t)t * C;
	double* lp = a.LP + (size_t)t * C;
	const double off = LN2 * (double)(a.fexp[t] + bexp_t);
	if (!a.tk)
	{
		// TN (NTK:220-251)
		for (uint32_t n = idx; n < C; n += stride)
		{
			double ext = 0.0, vm = 0.0;
			if (n + 1 < C) ext = nM[n + 1] * (emis_lin(a, x, (uint32_t)a.kmers[n]) * sc) * a.m;
			if (n > 0)
			{
				const double s = emis_lin(a, x, (uint32_t)a.kmers[n - 1]) * sc;
				vm = nE[n] * s;
				ext += vm * a.e;
			}
			cM[n] = vm;
			cE[n] = ext;
			lmax = max(lmax, (unsigned long long)__double_as_longlong(fmax(vm, ext)));
			lp[n] = log(fM[n] * vm + fE[n] * ext) + off;
		}
	}
	else
	{
		// TK (NTK:283-313): the successors of r, r+hp, r+2hp, r+3hp are the contiguous kmers 4r .. 4r+3
		const uint32_t hp = a.hp;
		for (uint32_t r = idx; r < hp; r += stride)
		{
			double R = 0.0;
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j) R += nM[4 * r + j] * (emis_lin(a, x, 4 * r + j) * sc);
			R *= a.m;
#pragma unroll
			for (uint32_t j = 0; j < 4; ++j)
			{
				const uint32_t q = r + j * hp;
				const double s = emis_lin(a, x, q) * sc;
				const double vm = nE[q] * s;
				const double ve = R + vm * a.e;
				cM[q] = vm;
				cE[q] = ve;
				lmax = max(lmax, (unsigned long long)__double_as_longlong(fmax(vm, ve)));
				lp[q] = log(fM[q] * vm + fE[q] * ve) + off;
			}
		}
	}
}

__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* s_red)
{
	for (int o = 16; o; o >>= 1)
	{
		const unsigned long long u = __shfl_xor_sync(FULL, v, o);
		v = max(v, u);
	}
	if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x < 32)
	{
		v = (threadIdx.x < (blockDim.x + 31) / 32) ? s_red[threadIdx.x] : 0ull;
		for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
		if (threadIdx.x == 0) s_red[0] = v;
	}
	__syncthreads();
	const unsigned long long r = s_red[0];
	__syncthreads();
	return r;
}

// row 0 of the forward lattice / the terminal row of the backward lattice
__device__ __forceinline__ void init_rows(const PreArgs& a, bool fwd, uint32_t idx, uint32_t stride)
{
	const uint32_t C = a.C, T = a.T;
	if (fwd)
	{
		for (uint32_t c = idx; c < C; c += stride)
		{
			a.fM[c] = 0.0;
			a.fE[c] = (a.tk || c == 0) ? 1.0 : 0.0;  // fE[0][0] = 0 (TN, NTK:205) / fE[0][q] = 0 for every q (TK, NTK:262) in log space
		}
		if (idx == 0)
		{
			a.fexp[0] = 0;
			a.fmax[0] = (unsigned long long)__double_as_longlong(1.0);
		}
	}
	else
	{
		double* bM = a.brow + (size_t)((T - 1) & 1u) * 2 * C;
		double* bE = bM + C;
		double* lp = a.LP + (size_t)(T - 1) * C;
		const double* fM = a.fM + (size_t)(T - 1) * C;
		const double* fE = a.fE + (size_t)(T - 1) * C;
		const double off = LN2 * (double)a.fexp[T - 1];
		for (uint32_t c = idx; c < C; c += stride)
		{
			const double ve = (a.tk || c == C - 1) ? 1.0 : 0.0;  // bE[T-1][N-1] = 0 (TN, NTK:228) / every q (TK, NTK:291)
			bM[c] = 0.0;
			bE[c] = ve;
			lp[c] = log(fM[c] * 0.0 + fE[c] * ve) + off;
		}
		if (idx == 0)
		{
			a.bexp[T - 1] = 0;
			a.bmax[T - 1] = (unsigned long long)__double_as_longlong(1.0);
		}
	}
}

// ---- small lattices: one CTA walks all rows of one direction (blockIdx.x: 0 forward, 1 backward needs the forward
// lattice, so the two are separate launches) --------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_pre_forward_cta(PreArgs a)
{
	__shared__ unsigned long long s_red[32];
	init_rows(a, true, threadIdx.x, blockDim.x);
	__syncthreads();
	for (uint32_t t = 1; t < a.T; ++t)
	{
		unsigned long long lmax = 0ull;
		fwd_row(a, t, threadIdx.x, blockDim.x, lmax);
		const unsigned long long rm = block_max_u64(lmax, s_red);  // (also the barrier between rows)
		if (threadIdx.x == 0)
		{
			a.fmax[t] = rm;
			a.fexp[t] = a.fexp[t - 1] - row_shift(a.fmax[t - 1]);
		}
		__syncthreads();
	}
}

__global__ void __launch_bounds__(1024) k_pre_backward_cta(PreArgs a)
{
	__shared__ unsigned long long s_red[32];
	const uint32_t C = a.C, T = a.T;
	init_rows(a, false, threadIdx.x, blockDim.x);
	__syncthreads();
	for (uint32_t t = T - 1; t-- > 0;)
	{
		const double* nM = a.brow + (size_t)((t + 1) & 1u) * 2 * C;
		double* cM = a.brow + (size_t)(t & 1u) * 2 * C;
		const int sh = row_shift(a.bmax[t + 1]);
		const int be = a.bexp[t + 1] - sh;
		unsigned long long lmax = 0ull;
		bwd_row(a, t, threadIdx.x, blockDim.x, nM, nM + C, cM, cM + C, be, pow2d(sh), lmax);
		const unsigned long long rm = block_max_u64(lmax, s_red);
		if (threadIdx.x == 0)
		{
			a.bmax[t] = rm;
			a.bexp[t] = be;
		}
		__syncthreads();
	}
}

// ---- large lattices: one launch per row over the whole GPU ------------------------------------------------------
__global__ void __launch_bounds__(256) k_pre_init(PreArgs a, int fwd)
{
	init_rows(a, fwd != 0, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(256) k_pre_forward_row(PreArgs a, uint32_t t)
{
	__shared__ unsigned long long s_red[32];
	unsigned long long lmax = 0ull;
	fwd_row(a, t, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, lmax);
	const unsigned long long rm = block_max_u64(lmax, s_red);
	if (threadIdx.x == 0)
	{
		atomicMax(&a.fmax[t], rm);  // fmax[t] zeroed by the host
		if (blockIdx.x == 0) a.fexp[t] = a.fexp[t - 1] - row_shift(a.fmax[t - 1]);
	}
}

__global__ void __launch_bounds__(256) k_pre_backward_row(PreArgs a, uint32_t t)
{
	__shared__ unsigned long long s_red[32];
	const uint32_t C = a.C;
	const double* nM = a.brow + (size_t)((t + 1) & 1u) * 2 * C;
	double* cM = a.brow + (size_t)(t & 1u) * 2 * C;
	const int sh = row_shift(a.bmax[t + 1]);
	const int be = a.bexp[t + 1] - sh;
	unsigned long long lmax = 0ull;
	bwd_row(a, t, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, nM, nM + C, cM, cM + C, be, pow2d(sh), lmax);
	const unsigned long long rm = block_max_u64(lmax, s_red);
	if (threadIdx.x == 0)
	{
		atomicMax(&a.bmax[t], rm);
		if (blockIdx.x == 0) a.bexp[t] = be;
	}
}

// Zf / Zb (natural log).  TN: Zf = fE[T-1][N-1], Zb = bE[0][0] (NTK:328-329); TK: the sums over all kmers (NTK:372-376).
__global__ void __launch_bounds__(1024) k_pre_z(PreArgs a)
{
	__shared__ double s_sum[32];
	const uint32_t C = a.C, T = a.T;
	const bool fwd = (blockIdx.x == 0);
	const double* row = fwd ? a.fE + (size_t)(T - 1) * C : a.brow + C;  // backward row 0 sits in ping-pong slot 0
	double v = 0.0;
	if (a.tk)
		for (uint32_t c = threadIdx.x; c < C; c += blockDim.x) v += row[c];
	else if (threadIdx.x == 0) v = fwd ? row[C - 1] : row[0];
	for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
	if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		double s = 0.0;
		for (uint32_t i = 0; i < (blockDim.x + 31) / 32; ++i) s += s_sum[i];
		a.z[fwd ? 0 : 1] = log(s) + LN2 * (double)(fwd ? a.fexp[T - 1] : a.bexp[0]);
	}
}

} // namespace ntk
} // namespace dyn
#endif
