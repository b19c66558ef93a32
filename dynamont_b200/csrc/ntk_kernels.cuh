// Resquiggle ("NTK") mode, pre-pass stages: the dense sequence x signal (TN) and
// kmer x signal (TK) two-state HMMs, their posterior-mass row masks and the sorted key list of the sparse
// 5-state lattice (reference NTK_aligner_api.cpp:120-441).
//
// First correct CUDA path of this mode: FP64 log-space arithmetic that mirrors the reference operation by operation
// (same association order, --fmad=false), one CTA per read with the lattice rows processed in sequence and the
// columns spread over the threads; the four lattice arrays live in HBM (T x C doubles each).  The row masks are
// what the sparse stages consume, so they are the parity surface of this file: membership is decided by a
// descending stable selection with a sequential log-sum-exp, exactly like columnArgsort + the insertion loop of
// preProcTN / preProcTK (:315-400).
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

struct PrepassArgs
{
	const double* signal;  // [S]
	const int32_t* kmers;  // [N-1] kmer id of column n is kmers[n-1]
	uint32_t T, N, K, hp;  // hp = 4^(k-1)
	Consts c;
	double *fM, *fE, *bM, *bE;  // [T][C], C = N (TN) or K (TK)
	double* LP;                 // [T][C]
	double* z;                  // [2] Zf, Zb
};

// ---- TN pre-pass (ppForTN / ppBackTN, NTK:197-251): dense T x N, no band ------------------------------------
__global__ void __launch_bounds__(1024) k_tn_fill(PrepassArgs a)
{
	const uint32_t T = a.T, N = a.N;
	const double NI = neg_inf();
	for (size_t i = threadIdx.x; i < (size_t)T * N; i += blockDim.x)
	{
		a.fM[i] = NI;
		a.fE[i] = NI;
		a.bM[i] = NI;
		a.bE[i] = NI;
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		a.fE[0] = 0.0;
		a.bE[(size_t)T * N - 1] = 0.0;
	}
	__syncthreads();
	for (uint32_t t = 1; t < T; ++t)
	{
		const double x = a.signal[t - 1];
		const double* pM = a.fM + (size_t)(t - 1) * N;
		const double* pE = a.fE + (size_t)(t - 1) * N;
		double* cM = a.fM + (size_t)t * N;
		double* cE = a.fE + (size_t)t * N;
		for (uint32_t n = 1 + threadIdx.x; n < N; n += blockDim.x)
		{
			const double sc = score_kmer(a.c, x, (uint32_t)a.kmers[n - 1]);
			cM[n] = pE[n - 1] + sc + a.c.m;
			cE[n] = log_plus(pM[n] + sc, pE[n] + sc + a.c.e);
		}
		__syncthreads();
	}
	for (uint32_t t = T - 1; t-- > 0;)
	{
		const double x = a.signal[t];
		const double* nM = a.bM + (size_t)(t + 1) * N;
		const double* nE = a.bE + (size_t)(t + 1) * N;
		double* cM = a.bM + (size_t)t * N;
		double* cE = a.bE + (size_t)t * N;
		for (uint32_t n = threadIdx.x; n < N; n += blockDim.x)
		{
			double ext = NI;
			if (n + 1 < N) ext = nM[n + 1] + score_kmer(a.c, x, (uint32_t)a.kmers[n]) + a.c.m;
			if (n > 0)
			{
				const double sc = score_kmer(a.c, x, (uint32_t)a.kmers[n - 1]);
				cM[n] = nE[n] + sc;
				ext = log_plus(ext, nE[n] + sc + a.c.e);
			}
			cE[n] = ext;
		}
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		a.z[0] = a.fE[(size_t)T * N - 1];  // Zf (NTK:328)
		a.z[1] = a.bE[0];                  // Zb (NTK:329)
	}
}

// ---- TK pre-pass (ppForTK / ppBackTK, NTK:253-313): two-state HMM over the de-Bruijn graph of all K kmers ----
__global__ void __launch_bounds__(1024) k_tk_fill(PrepassArgs a)
{
	const uint32_t T = a.T, K = a.K, hp = a.hp;
	const double NI = neg_inf();
	for (size_t i = threadIdx.x; i < (size_t)T * K; i += blockDim.x)
	{
		a.fM[i] = NI;
		a.fE[i] = NI;
		a.bM[i] = NI;
		a.bE[i] = NI;
	}
	__syncthreads();
	for (uint32_t k = threadIdx.x; k < K; k += blockDim.x)
	{
		a.fE[k] = 0.0;
		a.bE[(size_t)(T - 1) * K + k] = 0.0;
	}
	__syncthreads();
	for (uint32_t t = 1; t < T; ++t)
	{
		const double x = a.signal[t - 1];
		const double* pM = a.fM + (size_t)(t - 1) * K;
		const double* pE = a.fE + (size_t)(t - 1) * K;
		double* cM = a.fM + (size_t)t * K;
		double* cE = a.fE + (size_t)t * K;
		for (uint32_t k = threadIdx.x; k < K; k += blockDim.x)
		{
			double mat = NI;
			const double sc = score_kmer(a.c, x, k);
			for (uint32_t pre = k / 4; pre < K; pre += hp) mat = log_plus(mat, pE[pre] + sc + a.c.m);  // predecessorKmer(k, j)
			cM[k] = mat;
			cE[k] = log_plus(pM[k] + sc, pE[k] + sc + a.c.e);
		}
		__syncthreads();
	}
	for (uint32_t t = T - 1; t-- > 0;)
	{
		const double x = a.signal[t];
		const double* nM = a.bM + (size_t)(t + 1) * K;
		const double* nE = a.bE + (size_t)(t + 1) * K;
		double* cM = a.bM + (size_t)t * K;
		double* cE = a.bE + (size_t)t * K;
		for (uint32_t k = threadIdx.x; k < K; k += blockDim.x)
		{
			double ext = NI;
			const uint32_t s0 = (k % hp) * 4;  // successorKmer(k, 0)
			for (uint32_t suc = s0; suc < s0 + 4; ++suc) ext = log_plus(ext, nM[suc] + score_kmer(a.c, x, suc) + a.c.m);
			const double sc = score_kmer(a.c, x, k);
			cM[k] = nE[k] + sc;
			cE[k] = log_plus(ext, nE[k] + sc + a.c.e);
		}
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		// sequential log-sum-exp in the reference's order (NTK:372-376)
		double Zf = NI, Zb = NI;
		const size_t TK = (size_t)T * K;
		for (uint32_t k = 0; k < K; ++k)
		{
			Zf = log_plus(Zf, a.fE[TK - 1 - k]);
			Zb = log_plus(Zb, a.bE[k]);
		}
		a.z[0] = Zf;
		a.z[1] = Zb;
	}
}

// dense logP (NTK:179-195): LP = logPlus(fM + bM - Z, fE + bE - Z)
__global__ void k_dense_logp(PrepassArgs a, size_t size, double Z)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (size_t)gridDim.x * blockDim.x)
		a.LP[i] = log_plus(a.fM[i] + a.bM[i] - Z, a.fE[i] + a.bE[i] - Z);
}

// Row masks (NTK:343-353, 389-399): take the columns of row t in order of LP descending (stable: ties -> smaller
// index first, -inf last) until the sequential log-sum-exp of the taken values reaches SPARSETHRESHOLD (NTK:17: -0.02227639471 = log10(0.95), a mass of 0.978).
// One CTA per row; every round finds the largest remaining value with a block-wide arg-max.
__global__ void __launch_bounds__(256) k_row_mask(const double* LP, uint32_t C, uint32_t words, uint32_t* mask, double threshold)
{
	__shared__ double s_val[256];
	__shared__ uint32_t s_idx[256];
	__shared__ int s_stop;
	const uint32_t t = blockIdx.x;
	const double* row = LP + (size_t)t * C;
	uint32_t* m = mask + (size_t)t * words;
	for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) m[i] = 0u;
	__syncthreads();
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
		s_val[threadIdx.x] = bv;
		s_idx[threadIdx.x] = bi;
		__syncthreads();
		for (int o = 128; o; o >>= 1)
		{
			if ((int)threadIdx.x < o)
			{
				const uint32_t oi = s_idx[threadIdx.x + o];
				const double ov = s_val[threadIdx.x + o];
				const uint32_t mi = s_idx[threadIdx.x];
				const double mv = s_val[threadIdx.x];
				if (oi != 0xffffffffu && (mi == 0xffffffffu || ov > mv || (ov == mv && oi < mi)))
				{
					s_val[threadIdx.x] = ov;
					s_idx[threadIdx.x] = oi;
				}
			}
			__syncthreads();
		}
		if (threadIdx.x == 0)
		{
			const uint32_t i = s_idx[0];
			m[i >> 5] |= 1u << (i & 31);
			sum = log_plus(sum, s_val[0]);
			++taken;
			s_stop = (sum >= threshold || taken >= C) ? 1 : 0;
		}
		__syncthreads();
		if (s_stop) break;
		__syncthreads();
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

} // namespace ntk
} // namespace dyn

#endif // DYN_HOST_EMU
