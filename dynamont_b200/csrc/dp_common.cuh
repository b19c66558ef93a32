// Common device helpers for the warp-per-read banded HMM kernels (sm_100a).
//
// Arithmetic domain: log2.  FP32 state in registers, block-floating-point normalisation with one double
// offset per lane, MUFU ex2/lg2 for the log-sum-exp.  "Impossible" is a large finite negative number
// instead of -inf so that (-inf) - (-inf) never appears in the log-sum-exp.
#pragma once

#include <stdint.h>

#ifndef DYN_HOST_EMU
#include <cuda_runtime.h>
#define DYN_DEV __device__ __forceinline__
#define DYN_HD __host__ __device__ __forceinline__
#else
#define DYN_DEV inline
#define DYN_HD inline
#endif

namespace dyn
{

constexpr unsigned FULL = 0xffffffffu;
constexpr float NEG = -1.0e30f;      // forced value of an impossible / out-of-band cell
constexpr float CNEG = -1.0e25f;     // emission constant of an inactive ring slot
constexpr float BIGB = 1.0e18f;      // Cfg::UNI: b constant of an inactive ring slot (z^2 = 1e36 stays finite)
constexpr float DEADT = -1.0e20f;    // anything below this is "-inf" (real log2-probabilities are > -1e10)
constexpr double LOG2E = 1.4426950408889634074;
constexpr double LN2 = 0.69314718055994530942;

// ---- MUFU wrappers -------------------------------------------------------------------------------------
DYN_DEV float ex2(float x)
{
#ifndef DYN_HOST_EMU
	float y;
	asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
#else
	const float y = exp2f(x);
	return y < 1.17549435e-38f ? 0.0f : y;  // .ftz
#endif
}

DYN_DEV float lg2(float x)
{
#ifndef DYN_HOST_EMU
	float y;
	asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
#else
	return log2f(x);
#endif
}

// log2(2^u + 2^v); 2 MUFU + 4 FP32 ops.  Both operands finite (>= about -1e38).
DYN_DEV float logplus2(float u, float v)
{
	const float mx = fmaxf(u, v);
	const float e = ex2(-fabsf(u - v));
	return mx + lg2(1.0f + e);
}

// emission in the log2 domain: log2 N(x; mu, sigma) = c - (x*a - b)^2
DYN_DEV float emis2(float x, float a, float b, float c)
{
	const float z = fmaf(x, a, -b);
	return fmaf(-z, z, c);
}

DYN_DEV double shfl_f64(double v, int src)
{
	int lo = __double2loint(v), hi = __double2hiint(v);
	lo = __shfl_sync(FULL, lo, src);
	hi = __shfl_sync(FULL, hi, src);
	return __hiloint2double(hi, lo);
}

// band centre of row t: (size_t)(t * ratio) in IEEE double, exactly as the reference (NT_aligner_api.cpp:96-100)
DYN_DEV uint32_t band_mid(uint32_t t, double ratio)
{
	return (uint32_t)__double2ull_rz(__dmul_rn((double)t, ratio));
}

// Per-position emission constants (one per lattice column n; column n scores kmer[n-1]).
struct __align__(16) PosConst
{
	float a, b, c, pad;
};

// One sparse posterior record = the log2-posteriors (match, extend) of all CPL cells of ONE lane for one row.
// A lane is recorded when any of its cells has a non-negligible posterior; per row that is typically one
// lane (the alignment path crosses a lane boundary every CPL columns).  16-byte aligned for vector stores.
template <int CPL>
struct __align__(16) LaneRec
{
	static constexpr int NF = (2 * CPL + 1 + 3) / 4 * 4;  // floats, rounded up to a multiple of 4
	float v[NF];  // [0, CPL): match, [CPL, 2*CPL): extend, [2*CPL]: lane id (as int bits)
};

enum Status : int32_t
{
	ST_OK = 0,
	ST_SIGNAL_EMPTY = 1,        // "Signal is empty"                                  aligner.cpp:151
	ST_SEQ_SHORT = 2,           // "Sequence shorter than model kmer size"            aligner.cpp:156
	ST_SIGNAL_SHORT = 3,        // "Signal too short compared to sequence"            aligner.cpp:162
	ST_INVALID_NT = 4,          // "Invalid nucleotide: X"                            aligner.cpp:182,194
	ST_ALIGN_FAILED = 5,        // "Alignment failed: alignment scores do not match"  NT_aligner_api.cpp:291
	ST_TRAIN_FAILED = 6,        // "Training failed: alignment scores do not match"   NT_aligner_api.cpp:625
	ST_REC_OVERFLOW = 7,        // internal: sparse posterior buffer too small, host retries with a full-size buffer
	ST_INTERNAL = 8,
	ST_BAND_UNSUPPORTED = 9,
	ST_LIN_FAULT = 10,          // internal: the linear-domain kernel hit an FP32 range fault, host re-runs the read in the log2 domain
};

// Everything a warp needs to know about one read.  Built on the host, resident in HBM.
struct ReadDesc
{
	uint64_t sig_off;   // first sample in the batch signal array
	uint64_t pc_off;    // first PosConst (column 0) in the batch constants array
	uint64_t out_off;   // first segment in the batch output arrays (Kc entries)
	uint32_t S;         // samples;  T = S + 1 lattice rows
	uint32_t N;         // lattice columns = Kc + 1
	uint32_t bw;        // half band width = min(band/2, N/2)          NT_aligner_api.cpp:243
	int32_t status;     // Status; reads with status != ST_OK on entry are skipped
	double ratio;       // (double)N / (double)T                          NT_aligner_api.cpp:96
};

// Per-read results written by the kernels.
struct ReadOut
{
	double Z;       // natural-log partition function (backward), NT_aligner_api.cpp:286,293
	double dZ;      // Zf - Zb (natural log), for the reference's consistency check NT_aligner_api.cpp:288-291
	uint32_t nrec;  // sparse posterior records written
	int32_t status;
	double xi_m;    // training: expected number of E->M transitions
	double xi_e;    // training: expected number of E->E transitions
};

} // namespace dyn
