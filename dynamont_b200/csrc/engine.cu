// Host engine + C ABI of dynamont_b200 (see include/dynamont_b200.h).
//
// Host side of the hot path: model loading (reference aligner.cpp:88-143), input validation (:145-164),
// batch marshalling, scratch management, kernel launches, result assembly.  Built by nvcc for sm_100a.
// (tests/emu/ compiles this same file with g++ against a SIMT emulator for GPU-less unit tests of the
// kernels; that build is test infrastructure and is never loaded by the product.)
#include "../../include/dynamont_b200.h"
#include "dp_kernels.cuh"
#include "dp_linear.cuh"
#include "ntk_kernels.cuh"
#include "ntk_prepass.cuh"
#include "ribbon.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <numeric>
#include <set>
#include <sstream>
#include <thread>
#include <type_traits>
#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>

using namespace dyn;

// ---------------------------------------------------------------------------------------------------------
// runtime shim: CUDA runtime in the product, plain memory under the test emulator
// ---------------------------------------------------------------------------------------------------------
#ifndef DYN_HOST_EMU
#define CK_CUDA(x)                                                                                  \
	do                                                                                              \
	{                                                                                               \
		cudaError_t e_ = (x);                                                                       \
		if (e_ != cudaSuccess)                                                                      \
			throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + \
				__FILE__ + ":" + std::to_string(__LINE__));                                         \
	} while (0)

// device memory exhausted: run_batch releases the shared ribbon scratch pool and retries once
struct DevOom : std::runtime_error
{
	using std::runtime_error::runtime_error;
};

struct Rt
{
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[12] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
	int device = 0;
	int sms = 148;
	size_t smem_optin = 0;
	void init(int dev)
	{
		int count = 0;
		cudaError_t e = cudaGetDeviceCount(&count);
		if (e != cudaSuccess || count == 0)
			throw std::runtime_error("dynamont_b200: no usable CUDA device (there is no CPU fallback)");
		if (dev < 0) CK_CUDA(cudaGetDevice(&dev));
		device = dev;
		CK_CUDA(cudaSetDevice(device));
		cudaDeviceProp p;
		CK_CUDA(cudaGetDeviceProperties(&p, device));
		sms = p.multiProcessorCount;
		smem_optin = p.sharedMemPerBlockOptin;
		CK_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
		for (auto& v : ev) CK_CUDA(cudaEventCreate(&v));
	}
	bool own_stream = true;
	void use_stream(void* sh)
	{
		if (own_stream && stream) cudaStreamDestroy(stream);
		stream = (cudaStream_t)sh;
		own_stream = false;
	}
	void fini()
	{
		if (stream && own_stream) cudaStreamDestroy(stream);
		for (auto& v : ev)
			if (v) cudaEventDestroy(v);
	}
	void bind() { CK_CUDA(cudaSetDevice(device)); }
	bool async_alloc = false;  // stream-ordered allocation: no device-wide synchronisation (worker pools)
	void* dmalloc(size_t n)
	{
		void* p = nullptr;
		const cudaError_t e = async_alloc ? cudaMallocAsync(&p, n ? n : 1, stream) : cudaMalloc(&p, n ? n : 1);
		if (e == cudaErrorMemoryAllocation)
		{
			cudaGetLastError();  // not sticky: clear it
			throw DevOom("CUDA error: out of memory allocating " + std::to_string(n >> 20) + " MB of device memory");
		}
		CK_CUDA(e);
		return p;
	}
	void dfree(void* p)
	{
		if (async_alloc) cudaFreeAsync(p, stream);
		else cudaFree(p);
	}
	void* hmalloc(size_t n)
	{
		void* p = nullptr;
		CK_CUDA(cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault));  // pinned: device-to-host copies at link speed
		return p;
	}
	void hfree(void* p) { cudaFreeHost(p); }
	void h2d(void* d, const void* h, size_t n) { CK_CUDA(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, stream)); }
	void d2h(void* h, const void* d, size_t n) { CK_CUDA(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, stream)); }
	void zero(void* d, size_t n) { CK_CUDA(cudaMemsetAsync(d, 0, n, stream)); }
	void fill_ff(void* d, size_t n) { CK_CUDA(cudaMemsetAsync(d, 0xff, n, stream)); }
	void sync() { CK_CUDA(cudaStreamSynchronize(stream)); }
	size_t free_bytes()
	{
		size_t f = 0, t = 0;
		CK_CUDA(cudaMemGetInfo(&f, &t));
		return f;
	}
	void mark(int i) { CK_CUDA(cudaEventRecord(ev[i], stream)); }
	void mark_on(int i, const Rt& other) { CK_CUDA(cudaEventRecord(ev[i], other.stream)); }  // my event, recorded on other's stream
	void wait_on(const Rt& other, int i) { CK_CUDA(cudaStreamWaitEvent(other.stream, ev[i], 0)); }  // other's stream waits for my event
	void wait_self(int i) { CK_CUDA(cudaStreamWaitEvent(stream, ev[i], 0)); }
	double elapsed(int a, int b)
	{
		float ms = 0;
		CK_CUDA(cudaEventElapsedTime(&ms, ev[a], ev[b]));
		return ms;
	}
};
#else
#include <chrono>
struct Rt
{
	int device = 0;
	int sms = 2;
	bool async_alloc = false;
	size_t smem_optin = 227 * 1024;
	double tm[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	void init(int) {}
	void use_stream(void*) {}
	void fini() {}
	void bind() {}
	void* dmalloc(size_t n) { return malloc(n ? n : 1); }
	void dfree(void* p) { free(p); }
	void* hmalloc(size_t n) { return malloc(n ? n : 1); }
	void hfree(void* p) { free(p); }
	void h2d(void* d, const void* h, size_t n) { memcpy(d, h, n); }
	void d2h(void* h, const void* d, size_t n) { memcpy(h, d, n); }
	void zero(void* d, size_t n) { memset(d, 0, n); }
	void fill_ff(void* d, size_t n) { memset(d, 0xff, n); }
	void sync() {}
	size_t free_bytes() { return (size_t)4 << 30; }
	void mark(int i) { tm[i] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
	void mark_on(int i, const Rt&) { mark(i); }
	void wait_on(const Rt&, int) {}
	void wait_self(int) {}
	double elapsed(int a, int b) { return tm[b] - tm[a]; }
};
#endif

// ---------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------
namespace
{

// Build-time geometry variants: 13 columns per lane (ring of 416 >= 2*200+2), backward renormalisation every
// 4 rows, Viterbi renormalisation every 8 rows; checkpoint spacing CK and the register budget (resident
// single-warp CTAs per SM) trade shared memory / HBM scratch against latency hiding:
//   variant 0: CK = 16, 7 CTAs/SM (<= 255 registers, 30 KB smem)      scratch 224 B/row for checkpoints
//   variant 1: CK =  8, 10 CTAs/SM (<= 200 registers, 16 KB smem)     scratch 448 B/row
//   variant 2: CK =  8, 12 CTAs/SM (<= 168 registers, 16 KB smem)     scratch 448 B/row
//   variant 3: CK =  8, 8 CTAs/SM (<= 255 registers, 16 KB smem)      scratch 448 B/row
//   variants 4 .. 8: the uniform-sigma linear-domain kernels (Cfg::UNI) at 8 / 9 / 10 / 11 / 12 CTAs/SM (255 / 224 / 200 /
//                    184 / 168 registers); chosen only when every kmer of the model has the same sigma, otherwise they
//                    map to the general variants 3 / 9 / 1 / 1 / 2.  The log2-domain re-run of faulted reads uses the
//                    general kernels of variant 3.
//   variant 9: CK = 8, 9 CTAs/SM (<= 224 registers), general kernels
using Cfg16 = Cfg<13, 16, 4, 8>;
using Cfg8 = Cfg<13, 8, 4, 8>;
using Cfg8U = Cfg<13, 8, 4, 8, true>;
//   variants 10 / 11: renormalisation of the forward / backward values every 8 rows instead of 4 (uniform-sigma / general
//                     kernels at 8 CTAs/SM).  Reads they cannot represent (noisy reads: twice as many as with period 4)
//                     are re-run by the period-4 linear kernels of variant 4 / 3, and what those cannot represent by the
//                     log2-domain kernels of variant 3.
using Cfg8R8 = Cfg<13, 8, 8, 8>;
using Cfg8UR8 = Cfg<13, 8, 8, 8, true>;
//   variants 12 / 13: variants 10 / 11 launched as CTAs of 8 warps that run their reads pass by pass in step
constexpr int N_VARIANTS = 14;
// experiment builds: -DDYN_ONLY_VARIANT=n compiles the kernels of one variant only (seconds instead of minutes)
#ifndef DYN_V12_WPC
#define DYN_V12_WPC 8  // warps per CTA of the phase-synchronised launch shape (one CTA per SM)
#endif
#ifndef DYN_V10_MINB
#define DYN_V10_MINB 8  // resident CTAs per SM of the default uniform-sigma kernels (experiment builds override it)
#endif
#ifdef DYN_ONLY_VARIANT
#define DYN_HAS(n) ((n) == DYN_ONLY_VARIANT)
#else
#define DYN_HAS(n) 1
#endif
constexpr int DEFAULT_VARIANT = 13;      // measured fastest on B200
constexpr int DEFAULT_VARIANT_UNI = 12;  // uniform-sigma models

struct EncodeArgs
{
	ReadDesc* reads;
	uint32_t n_reads;
	const char* seq;
	const uint64_t* seq_off;
	int k;
	const PosConst* table;
	PosConst* pc;
	int32_t* kmers;
	uint32_t* bad_pos;
	ReadOut* out;      // zeroed here
	uint32_t* queue;   // work counters of the DP kernels, zeroed here
};

DYN_DEV int base_digit(unsigned char ch)
{
	// aligner.cpp:46-60 (N/n map to 4, which is >= the alphabet size and therefore invalid, :180)
	switch (ch)
	{
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': case 'U': case 'u': return 3;
	case 'N': case 'n': return 4;
	default: return -1;
	}
}

// K1 (emission constants): kmer encoding (aligner.cpp:166-205) fused with the gather of the per-column
// Gaussian constants from the pore-model table, so the 4^k table is touched N times per read and never in
// the DP inner loop.  One warp-CTA per read.
DYN_DEV void encode_read(const EncodeArgs& a, uint32_t r, int lane)
{
	ReadDesc rd = a.reads[r];
	if (lane == 0)
	{
		ReadOut o;
		o.Z = 0.0; o.dZ = 0.0; o.nrec = 0; o.status = 0; o.xi_m = 0.0; o.xi_e = 0.0;
		a.out[r] = o;
		a.bad_pos[r] = 0xffffffffu;
		if (r == 0) a.queue[0] = 0u;
	}
	if (rd.status != ST_OK) return;
	const char* s = a.seq + a.seq_off[r];
	const uint32_t Kc = rd.N - 1;
	PosConst* pc = a.pc + rd.pc_off;
	int32_t* km = a.kmers + rd.pc_off;
	if (lane == 0)
	{
		PosConst z;
		z.a = 0.0f; z.b = 0.0f; z.c = CNEG; z.pad = 0.0f;
		pc[0] = z;  // column 0 scores no kmer
		km[0] = -1;
	}
	uint32_t bad = 0xffffffffu;
	for (uint32_t c = lane; c < Kc; c += 32)
	{
		int id = 0;
		for (int i = 0; i < a.k; ++i)
		{
			const int d = base_digit((unsigned char)s[c + i]);
			if (d < 0 || d > 3)
			{
				bad = min(bad, c + (uint32_t)i);
				id = -1;
				break;
			}
			id = id * 4 + d;
		}
		if (id >= 0)
		{
			pc[c + 1] = a.table[id];
			km[c + 1] = id;
		}
	}
	for (int o = 16; o; o >>= 1) bad = min(bad, __shfl_sync(FULL, bad, (lane + o) & 31));
	if (lane == 0)
	{
		a.bad_pos[r] = bad;
		if (bad != 0xffffffffu) a.reads[r].status = ST_INVALID_NT;
	}
}

// WPC = 1: a persistent grid of single-warp CTAs, each pulling one read at a time.  WPC > 1 (linear-domain kernels
// only): CTAs of WPC warps pull WPC consecutive reads of the processing order (sorted by size, so nearly equal) and run
// them pass by pass in step (lin::cta_sync).
template <class CFG, int MODE, bool LIN, int WPC = 1>
DYN_DEV void align_worker(const BatchArgs& args, unsigned char* smem_all, int tid, unsigned cta)
{
	const int lane = tid & 31, wid = tid >> 5;
	unsigned char* smem_raw = smem_all + (size_t)wid * CFG::SMEM_BYTES;
	const SlotScratch sc = args.slots[cta * WPC + wid];
#ifndef DYN_HOST_EMU
	__shared__ uint32_t s_base, s_kb;
#endif
	while (true)
	{
		uint32_t i = 0;
		if (WPC == 1)
		{
			if (lane == 0) i = atomicAdd(args.queue, 1u);
			i = __shfl_sync(FULL, i, 0);
		}
#ifndef DYN_HOST_EMU
		else
		{
			if (tid == 0)
			{
				s_base = atomicAdd(args.queue, (uint32_t)WPC);
				s_kb = 0u;
			}
			__syncthreads();
			i = s_base + (uint32_t)wid;
			__syncthreads();
			if (i - (uint32_t)wid >= args.n_reads) break;  // the whole CTA leaves together
		}
#endif
		const bool live = i < args.n_reads;
		if (WPC == 1 && !live) break;
		const uint32_t ridx = live ? args.order[i] : 0u;
		ReadDesc rd;
		if (live) rd = args.reads[ridx];
		else rd.status = ST_INTERNAL;
		if (live && rd.status != ST_OK)
		{
			if (lane == 0)
			{
				ReadOut o;
				o.Z = 0.0; o.dZ = 0.0; o.nrec = 0; o.status = rd.status; o.xi_m = 0.0; o.xi_e = 0.0;
				args.out[ridx] = o;
			}
			if (WPC == 1) continue;
		}
		uint32_t kb_sync = 0;
#ifndef DYN_HOST_EMU
		if (WPC > 1)
		{
			// the largest number of pass-2 blocks among the CTA's reads (lin::forward_posterior_pass)
			if (lane == 0 && live && rd.status == ST_OK) atomicMax(&s_kb, rd.S / (uint32_t)CFG::CK);
			__syncthreads();
			kb_sync = s_kb;
		}
#endif
		if (LIN) lin::align_read<CFG, MODE, WPC>(args, rd, ridx, sc, smem_raw, lane, live && rd.status == ST_OK, kb_sync);
		else align_read<CFG, MODE>(args, rd, ridx, sc, smem_raw, lane);
		__syncwarp();
	}
}

struct FoldArgs
{
	const ReadDesc* reads;
	const ReadOut* out;
	uint32_t n_reads;
	const int32_t* kmers;
	const double* read_w;
	const double* read_x;
	const double* read_xx;
	double* stat_w;
	double* stat_x;
	double* stat_xx;
};

// folds per-column training statistics into the pooled per-kmer table (one warp-CTA per read)
DYN_DEV void fold_read(const FoldArgs& a, uint32_t r, int lane)
{
	const ReadDesc rd = a.reads[r];
	if (a.out[r].status != ST_OK) return;
	for (uint32_t n = 1 + lane; n < rd.N; n += 32)
	{
		const double wv = a.read_w[rd.pc_off + n];
		if (wv == 0.0) continue;
		const int32_t q = a.kmers[rd.pc_off + n];
		atomicAdd(&a.stat_w[q], wv);
		atomicAdd(&a.stat_x[q], a.read_x[rd.pc_off + n]);
		atomicAdd(&a.stat_xx[q], a.read_xx[rd.pc_off + n]);
	}
}

// Signal preprocessing of the front end, the step right before the DP: z-normalisation with the
// basecaller's shift / scale (segment.py:151-152, train.py:168-169) and the Hampel outlier filter
// (utils.py:16-43; window 3 / 3 sigma for segmentation, 7 / 5 sigma for training) — float64 like the numpy code,
// rounded to FP32 (what the DP consumes) at the very end.
struct PreprocArgs
{
	const float* raw;        // concatenated raw samples
	const uint64_t* sig_off; // [n_reads + 1]
	uint32_t n_reads;
	const double* shift;     // [n_reads]
	const double* scale;     // [n_reads]
	int window;              // <= 15
	double n_sigmas;
	float* out;              // concatenated, same offsets
};

DYN_DEV double small_median(double* v, int m)
{
	for (int i = 1; i < m; ++i)
	{
		const double x = v[i];
		int j = i;
		while (j > 0 && v[j - 1] > x)
		{
			v[j] = v[j - 1];
			--j;
		}
		v[j] = x;
	}
	return (m & 1) ? v[m / 2] : (v[m / 2 - 1] + v[m / 2]) / 2.0;  // numpy.median
}

DYN_DEV void preprocess_sample(const PreprocArgs& a, uint64_t g)
{
	// read of sample g: largest r with sig_off[r] <= g
	uint32_t lo = 0, hi = a.n_reads;
	while (hi - lo > 1)
	{
		const uint32_t mid = (lo + hi) / 2;
		if (a.sig_off[mid] <= g) lo = mid;
		else hi = mid;
	}
	const uint64_t base = a.sig_off[lo], n = a.sig_off[lo + 1] - base, i = g - base;
	const double sh = a.shift[lo], sc = a.scale[lo];
	const float* raw = a.raw + base;
	const int W = a.window, h = W / 2;
	double v = ((double)raw[i] - sh) / sc;
	// utils.py:26-40: centres original[W//2 : n - W//2 - 1], window j = original[j : j + W]
	if (n > (uint64_t)W && i >= (uint64_t)h && i + h + 1 < n)
	{
		double w[16], d[16];
		for (int q = 0; q < W; ++q) w[q] = ((double)raw[i - h + q] - sh) / sc;
		for (int q = 0; q < W; ++q) d[q] = w[q];
		const double med = small_median(d, W);
		for (int q = 0; q < W; ++q) d[q] = fabs(w[q] - med);
		const double sigma = 1.4826 * small_median(d, W);
		if (fabs(v - med) > a.n_sigmas * sigma) v = med;
	}
	a.out[g] = (float)v;
}

#ifndef DYN_HOST_EMU
__global__ void k_preprocess(PreprocArgs a, uint64_t total)
{
	for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (uint64_t)gridDim.x * blockDim.x)
		preprocess_sample(a, g);
}
__global__ void __launch_bounds__(32) k_encode(EncodeArgs a)
{
	for (uint32_t r = blockIdx.x; r < a.n_reads; r += gridDim.x) encode_read(a, r, threadIdx.x);
}
// register budget for MINB resident single-warp CTAs per SM (64 K registers, allocation granularity 8 per thread).
// __launch_bounds__(32, MINB) makes ptxas fall to 168 registers for every MINB > 8; __maxnreg__ gives the exact budget.
constexpr int max_regs(int minb) { return (65536 / (32 * minb) / 8 * 8) > 255 ? 255 : (65536 / (32 * minb) / 8 * 8); }
template <class CFG, int MODE, int MINB, bool LIN, int WPC = 1>
__global__ void __launch_bounds__(32 * WPC) __maxnreg__(max_regs(MINB * WPC)) k_align(BatchArgs args)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	align_worker<CFG, MODE, LIN, WPC>(args, smem_raw, threadIdx.x, blockIdx.x);
}
__global__ void __launch_bounds__(32) k_fold(FoldArgs a)
{
	for (uint32_t r = blockIdx.x; r < a.n_reads; r += gridDim.x) fold_read(a, r, threadIdx.x);
}
#endif

void launch_encode(Rt& rt, const EncodeArgs& a)
{
	if (!a.n_reads) return;
#ifndef DYN_HOST_EMU
	const unsigned grid = std::min<uint32_t>(a.n_reads, 65535u * 16u);
	k_encode<<<grid, 32, 0, rt.stream>>>(a);
	CK_CUDA(cudaGetLastError());
#else
	(void)rt;
	simt::launch(a.n_reads, 0, [&]() { encode_read(a, blockIdx.x, threadIdx.x); });
#endif
}

// grid = resident warps ("slots"); WPC warps per CTA (see align_worker)
template <class CFG, int MINB, bool LIN, int WPC = 1>
void launch_align_t(Rt& rt, const BatchArgs& args, unsigned grid, int mode)
{
	const size_t smem = CFG::SMEM_BYTES * WPC;
#ifndef DYN_HOST_EMU
	// the shared-memory opt-in is a per-DEVICE attribute of the function: set it for the kernel about to be launched on
	// the current device every time (cheap), so that a second handle on another device of the same process works too
	if (mode == 1) CK_CUDA(cudaFuncSetAttribute(k_align<CFG, 1, MINB, LIN, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	else if (mode == 2) CK_CUDA(cudaFuncSetAttribute(k_align<CFG, 2, MINB, LIN, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	const unsigned ctas = std::max(1u, grid / WPC);  // the host sizes the slot table for grid warps
	if (mode == 0) k_align<CFG, 0, MINB, LIN, WPC><<<ctas, 32 * WPC, 0, rt.stream>>>(args);  // the backward pass alone needs no shared memory
	else if (mode == 1) k_align<CFG, 1, MINB, LIN, WPC><<<ctas, 32 * WPC, smem, rt.stream>>>(args);
	else k_align<CFG, 2, MINB, LIN, WPC><<<ctas, 32 * WPC, smem, rt.stream>>>(args);
	CK_CUDA(cudaGetLastError());
#else
	(void)rt;
	if (mode == 0) simt::launch(grid, smem, [&]() { align_worker<CFG, 0, LIN>(args, simt::dyn_smem(), threadIdx.x, blockIdx.x); });
	else if (mode == 1) simt::launch(grid, smem, [&]() { align_worker<CFG, 1, LIN>(args, simt::dyn_smem(), threadIdx.x, blockIdx.x); });
	else simt::launch(grid, smem, [&]() { align_worker<CFG, 2, LIN>(args, simt::dyn_smem(), threadIdx.x, blockIdx.x); });
#endif
}

// lin: the linear-domain kernels (dp_linear.cuh); otherwise the log2-domain kernels (dp_kernels.cuh)
template <class CFG, int MINB, class CFGLIN, int MINB_FB, int WPC>
void launch_align(Rt& rt, const BatchArgs& args, unsigned grid, int mode, bool lin)
{
	if (lin) launch_align_t<CFGLIN, MINB / WPC, true, WPC>(rt, args, grid, mode);
	else launch_align_t<CFG, MINB_FB, false>(rt, args, grid, mode);
}

// pooled training on device-resident statistics: stats[3K + 4] = w[K], x[K], xx[K], xi_m, xi_e, sum Z, reads ok
struct SumOutArgs
{
	const ReadOut* out;
	uint32_t n_reads;
	double* tail;  // stats + 3K
};

DYN_DEV void sum_out_lane(const SumOutArgs& a, uint32_t first, uint32_t step, int lane)
{
	double xm = 0.0, xe = 0.0, z = 0.0, ok = 0.0;
	for (uint32_t r = first + lane; r < a.n_reads; r += step)
	{
		const ReadOut o = a.out[r];
		if (o.status != ST_OK) continue;
		xm += o.xi_m;
		xe += o.xi_e;
		z += o.Z;
		ok += 1.0;
	}
	for (int off = 16; off; off >>= 1)
	{
		xm += shfl_f64(xm, (lane + off) & 31);
		xe += shfl_f64(xe, (lane + off) & 31);
		z += shfl_f64(z, (lane + off) & 31);
		ok += shfl_f64(ok, (lane + off) & 31);
	}
	if (lane == 0 && ok > 0.0)
	{
		atomicAdd(&a.tail[0], xm);
		atomicAdd(&a.tail[1], xe);
		atomicAdd(&a.tail[2], z);
		atomicAdd(&a.tail[3], ok);
	}
}

// M-step of NT_aligner_api.cpp:519-535 on pooled statistics, one thread per kmer: kmers without weight keep the model
struct MStepArgs
{
	const double* stats;
	uint64_t K;
	double* mean;   // in: old model, out: new model
	double* stdev;
};

DYN_DEV void mstep_kmer(const MStepArgs& a, uint64_t q)
{
	const double w = a.stats[q];
	if (!(w > 0.0)) return;
	const double mu = a.stats[a.K + q] / w;
	double var = a.stats[2 * a.K + q] / w - mu * mu;
	if (var < 1e-12) var = 1e-12;
	a.mean[q] = mu;
	a.stdev[q] = sqrt(var);
}

#ifndef DYN_HOST_EMU
__global__ void __launch_bounds__(32) k_sum_out(SumOutArgs a) { sum_out_lane(a, blockIdx.x * 32u, gridDim.x * 32u, threadIdx.x); }
__global__ void k_mstep(MStepArgs a)
{
	for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < a.K; q += (uint64_t)gridDim.x * blockDim.x) mstep_kmer(a, q);
}
#endif

void launch_fold(Rt& rt, const FoldArgs& a)
{
	if (!a.n_reads) return;
#ifndef DYN_HOST_EMU
	const unsigned grid = std::min<uint32_t>(a.n_reads, 65535u * 16u);
	k_fold<<<grid, 32, 0, rt.stream>>>(a);
	CK_CUDA(cudaGetLastError());
#else
	(void)rt;
	simt::launch(a.n_reads, 0, [&]() { fold_read(a, blockIdx.x, threadIdx.x); });
#endif
}

// ---------------------------------------------------------------------------------------------------------
// device buffer that only ever grows
// ---------------------------------------------------------------------------------------------------------
struct DevBuf
{
	void* p = nullptr;
	size_t cap = 0;
	void* get(Rt& rt, size_t n)
	{
		if (n > cap)
		{
			const bool first = (cap == 0);
			if (p) rt.dfree(p);
			p = nullptr;
			cap = 0;
			// growing costs a cudaFree + cudaMalloc, which synchronise the whole device (and stall the other lane's running
			// kernel): after the first allocation grow with slack so that batches of varying size settle quickly
			const size_t want = rt.async_alloc ? n + n / 2 : (first ? n : n + n / 8);
			p = rt.dmalloc(want);
			cap = want;
		}
		return p;
	}
	void release(Rt& rt)
	{
		if (p) rt.dfree(p);
		p = nullptr;
		cap = 0;
	}
};

// pinned host staging buffer that only ever grows (results of a batch land here before they are fanned out into the
// caller's arrays by a few host threads)
struct HostBuf
{
	void* p = nullptr;
	size_t cap = 0;
	void* get(Rt& rt, size_t n)
	{
		if (n > cap)
		{
			if (p) rt.hfree(p);
			p = nullptr;
			cap = 0;
			const size_t want = n + n / 4;
			p = rt.hmalloc(want);
			cap = want;
		}
		return p;
	}
	void release(Rt& rt)
	{
		if (p) rt.hfree(p);
		p = nullptr;
		cap = 0;
	}
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// host phases of a batch call, printed to stderr when DYN_TIMING is set (bench.py reports the kernel time separately)
struct HostTimer
{
	bool on;
	std::chrono::steady_clock::time_point t0;
	std::string log;
	HostTimer() : on(getenv("DYN_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
	void lap(const char* what)
	{
		if (!on) return;
		const auto t1 = std::chrono::steady_clock::now();
		char b[96];
		snprintf(b, sizeof(b), " %s=%.1fms", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
		log += b;
		t0 = t1;
	}
	~HostTimer()
	{
		if (on && !log.empty()) fprintf(stderr, "[dyn timing]%s\n", log.c_str());
	}
};

// fn(begin, end) over [0, n) on up to 8 host threads (memory-bound fan-out of results; first-touch page faults of the
// caller's arrays are spread over the threads too)
template <class F>
void parallel_ranges(uint32_t n, uint64_t work, F fn)
{
	unsigned nt = std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
	if (work < (1u << 20) || n < 2 * nt) nt = 1;
	if (nt == 1)
	{
		fn(0u, n);
		return;
	}
	std::vector<std::thread> th;
	for (unsigned i = 0; i < nt; ++i)
	{
		const uint32_t a = (uint32_t)((uint64_t)n * i / nt), b = (uint32_t)((uint64_t)n * (i + 1) / nt);
		th.emplace_back([=]() { fn(a, b); });
	}
	for (auto& t : th) t.join();
}

} // namespace

// ---------------------------------------------------------------------------------------------------------
// the aligner handle
// ---------------------------------------------------------------------------------------------------------
struct dyn_aligner
{
	Rt rt;
	std::mutex mu;
	std::string last_error;
	// pore / model (aligner.cpp:13-143)
	bool rna = false;
	int k = 0;
	int alphabet = 0;
	uint64_t K = 0;
	int band = 400;
	double trans[3] = {0, 0, 0};  // log m1, e1, e2 (NT:84-86)
	std::vector<double> mean, stdev;
	// tuning
	int warps_per_sm = 0;  // 0 = the variant's own occupancy
	int variant = -1;  // -1: DEFAULT_VARIANT / DEFAULT_VARIANT_UNI (measured fastest on B200)
	int fwd_fast = 1;  // linear-domain pass 2: branch-free row body for groups of rows without a band slide
	bool uniform = false;  // every kmer has the same sigma (set by upload_table): the Cfg::UNI kernels apply
	float uni_a = 0.0f, uni_c = 0.0f;
	bool ntk_pool_used = false;  // dyn_ntk_align_batch raised the release threshold of the device's memory pool
	bool ntk = false; // resquiggle (NTK) mode: only the pre-pass stages are built (dyn_ntk_prepass)
	double ntk_trans[18] = {0};  // log a1,a2,p1-3,s1-3,e1-4,i1,i2, then log ntMatch/ntExtend for TN and TK (NTK:35-104)
	int arith = 0;    // 0: linear-domain kernels, reads with an FP32 range fault re-run in the log2 domain; 1: log2 domain only
	uint64_t n_fallback = 0;  // reads of the last batch that were re-run in the log2 domain
	int last_variant = -1;    // resolved build variant of the last batch
	uint64_t n_retry_lin = 0; // reads of the last batch that were re-run by the second-tier linear-domain kernels
	double thr2 = -22.0;
	// ribbon kernels (dp_ribbon.cuh): first tier for every read whose band is wider than the window
	int ribbon = 2;            // lattice columns per lane of the window (2: 63 columns, 4: 127 columns); 0: off
	int rib_guard = 40;        // the window's edge lanes must stay this many bits below the row maximum
	double thr_rib = -16.0;    // log2 of the posterior above which a lane is recorded (unrecorded path cells count as 0)
	double rib_recs_per_row = 2.5;   // lane records per lattice row budgeted (measured use: 1.2; a read that needs more faults to the full-band kernels)
	int rib_min_bw = 0;        // reads whose half band is narrower than this skip the ribbon tier (0: none; the band is clipped exactly inside the window)
	int rib_two_level = -1;    // checkpoints of every 8th group only: -1 when the scratch would not fit otherwise, 0 never, 1 always
	int rib_log = 1;           // 1: reads the linear-domain ribbon loses to FP32 range are re-run by the log2-domain ribbon (align)
	uint64_t n_rib_log = 0, n_rib_log_fault = 0;  // last batch: reads given to the log2-domain ribbon / reads it handed on
	int rib_gather = -1;       // records-free scratch + second forward sweep for the path posteriors (implies two-level):
	                           // -1 when the resident warps' scratch would not fit otherwise, 0 never, 1 always
	double rib_last_two_level = 0;
	double rib_recs_used = 0;  // lane records per lattice row the last batch actually wrote (align mode)
	int rib_bps = 0;           // resident CTAs (of 4 warps) per SM of the ribbon kernels; 0: the build's default
	uint64_t n_ribbon = 0;     // reads of the last batch the ribbon kernels were given ...
	uint64_t n_rib_fault = 0;  // ... and how many of them they handed on to the full-band kernels
	uint64_t rib_reason[16] = {0};  // cumulative over the handle's life: faults by reason (dp_ribbon.cuh ribbon_read)
	double ribbon_ms = 0.0;
	double recs_per_row = 1.6;  // lane records per row (typical use: ~1.1; a read that overflows is retried alone with a full buffer)
	double mem_fraction = 0.92;  // share of the free HBM the scratch of the resident warps may take
	// device state
	DevBuf d_table, d_sig, d_seq, d_seqoff, d_desc, d_order, d_pc, d_kmers, d_bad, d_out, d_sigpos, d_prob, d_scratch,
		d_slots, d_queue, d_rw, d_rx, d_rxx, d_sw, d_sx, d_sxx;
	HostBuf h_sigpos, h_prob;  // pinned staging of a batch's segment borders / probabilities
	// pinned bump arena for the small per-call tables (descriptors, orders, slot tables, per-read outputs).  Copies from / to
	// pageable memory are staged by the driver in a way that waits for kernels running on OTHER streams (measured: a
	// 1 MB pageable D2H issued next to another lane's ribbon kernel returned when that kernel ended, 300 ms later), which
	// serialised the lanes; everything a batch call copies in steady state goes through pinned memory instead.
	HostBuf h_arena;
	size_t arena_off = 0;
	void* stage(size_t bytes)
	{
		const size_t at = (arena_off + 255) / 256 * 256;
		if (at + bytes > h_arena.cap) throw std::runtime_error("dynamont_b200: pinned staging arena exhausted");
		arena_off = at + bytes;
		return (unsigned char*)h_arena.p + at;
	}
	void h2d_staged(void* d, const void* h, size_t bytes)
	{
		if (!bytes) return;
		void* st = stage(bytes);
		std::memcpy(st, h, bytes);
		rt.h2d(d, st, bytes);
	}
	bool table_dirty = true;
	double timing[3] = {0, 0, 0};
	double* ext_stats = nullptr;  // dyn_train_accumulate: caller-owned device buffer [3K + 4] the statistics are added to
	// asynchronous entry points (dyn_align_submit / dyn_align_wait): calls alternate between LANES child handles on the
	// same device, each with its own stream and buffers, so that the host-to-device copy of call i+1 and the result
	// copy + fan-out of call i-1 overlap the kernels of call i
	static constexpr int LANES = 3;  // (a third lane lets a long full-band hand-over of one batch overlap two ribbon kernels: config 4)
	dyn_aligner* lane[LANES] = {nullptr, nullptr, nullptr};
	// The lanes run their ribbon kernels on the ROOT handle's stream and in the root's scratch: kernels of successive
	// batches execute one after the other anyway (each fills the GPU), so one scratch pool — the largest allocation by
	// far, tens of GB — serves all of them.  compute_mu orders "size the pool, build the slot table, enqueue the kernel".
	dyn_aligner* root = nullptr;  // nullptr: this handle is a root
	std::mutex compute_mu;
	DevBuf d_rib_scratch;
	std::string model_path_, pore_, mode_;
	int device_ = -1;
	struct Job
	{
		std::thread th;
		int rc = 0;
		bool joined = false;
	};
	std::mutex jobs_mu;
	std::vector<Job*> jobs;
	std::vector<std::pair<std::string, double>> options;  // replayed on the lanes

	void upload_table();
};

namespace
{

struct PoreInfo
{
	const char* name;
	bool rna;
	int k;
	double m1, e1, e2;
};
// aligner.cpp:62-86 and NT_aligner_api.cpp:36-82
const PoreInfo PORES[] = {
	{"rna002", true, 5, 0.019889650396799997, 1.0, 0.9801103496029998},
	{"rna004", true, 9, 0.031111753637096777, 1.0, 0.9688882463622581},
	{"dna_r9", false, 5, 1.0, 1.0, 1.0},
	{"dna_r10_260bps", false, 9, 0.031111753637096777, 1.0, 0.9688882463622581},
	{"dna_r10_400bps", false, 9, 0.031111753637096777, 1.0, 0.9688882463622581},
};

int host_digit(unsigned char ch)
{
	switch (ch)
	{
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': case 'U': case 'u': return 3;
	case 'N': case 'n': return 4;
	default: return -1;
	}
}

// Aligner::loadModel (aligner.cpp:88-143) incl. its error strings
void load_model(dyn_aligner& A, const std::string& path)
{
	std::ifstream file(path);
	if (!file) throw std::runtime_error("Could not open model file, please prove a valid model path " + path);
	std::string line;
	std::getline(file, line);
	std::set<char> alphabet;
	std::vector<std::string> rows;
	while (std::getline(file, line))
	{
		const std::string kmer = line.substr(0, line.find('\t'));
		if (kmer.size() != (size_t)A.k) throw std::runtime_error("Inconsistent kmer size in model");
		for (char c : kmer) alphabet.insert(c);
		rows.push_back(line);
	}
	A.alphabet = (int)alphabet.size();
	A.K = (uint64_t)std::pow((double)A.alphabet, (double)A.k);
	if (A.alphabet != 4)
		throw std::runtime_error("dynamont_b200: pore models must use a 4-letter alphabet (found " +
			std::to_string(A.alphabet) + ")");
	A.mean.assign(A.K, 0.0);
	A.stdev.assign(A.K, 0.0);
	for (const std::string& row : rows)
	{
		std::stringstream ss(row);
		std::string kmer, m, s;
		std::getline(ss, kmer, '\t');
		std::getline(ss, m, '\t');
		std::getline(ss, s, '\t');
		if (A.rna) std::reverse(kmer.begin(), kmer.end());
		uint64_t v = 0;
		for (char c : kmer)
		{
			const int d = host_digit((unsigned char)c);
			if (d < 0 || d >= A.alphabet) throw std::runtime_error("Invalid nucleotide in k-mer: " + kmer);
			v = v * (uint64_t)A.alphabet + (uint64_t)d;
		}
		A.mean[v] = std::stod(m);
		A.stdev[v] = std::stod(s);
	}
}

} // namespace

void dyn_aligner::upload_table()
{
	// per-kmer constants of log2 N(x; mu, sigma) = c - (x*a - b)^2  (aligner.cpp:287-292 in the log2 domain)
	std::vector<PosConst> t(K);
	for (uint64_t q = 0; q < K; ++q)
	{
		const double sd = stdev[q], mu = mean[q];
		const double a = std::sqrt(0.5 * LOG2E) / sd;
		t[q].a = (float)a;
		t[q].b = (float)(mu * a);
		t[q].c = (float)(-std::log2(sd) - 0.5 * std::log2(2.0 * M_PI));
		t[q].pad = (float)mu;  // centre of the ribbon kernels' training statistics
	}
	uniform = K > 0;
	for (uint64_t q = 1; q < K && uniform; ++q) uniform = (t[q].a == t[0].a && t[q].c == t[0].c);
	uni_a = K ? t[0].a : 0.0f;
	uni_c = K ? t[0].c : 0.0f;
	void* d = d_table.get(rt, K * sizeof(PosConst));
	rt.h2d(d, t.data(), K * sizeof(PosConst));
	rt.sync();
	table_dirty = false;
}

// ---------------------------------------------------------------------------------------------------------
// batch driver
// ---------------------------------------------------------------------------------------------------------
namespace
{

struct BatchIO
{
	const float* sig_host = nullptr;    // exactly one of sig_host / sig_host64 / sig_dev is set
	const double* sig_host64 = nullptr;
	const float* sig_dev = nullptr;
	const char* seq_host = nullptr;     // or seq_dev
	const char* seq_dev = nullptr;
	const uint64_t* sig_off = nullptr;
	const uint64_t* seq_off = nullptr;
	uint32_t n = 0;
};

struct BatchResult
{
	std::vector<ReadOut> out;
	std::vector<uint32_t> bad_pos;
	std::vector<uint64_t> seg_off;  // n+1
	std::vector<ReadDesc> desc;
};

// mode: 0 Z only, 1 align, 2 train.  sigpos/prob (host) receive the segment arrays for mode 1.
// CFGLIN: configuration of the linear-domain kernels of the first launch.  CFGLIN2 (optional): linear-domain kernels
// with a shorter renormalisation period that re-run the reads CFGLIN could not represent; what they cannot represent
// either goes to the log2-domain kernels (CFG, MINB_FB resident CTAs per SM).  All three share CK, i.e. the scratch.
template <class CFG, int MINB, class CFGLIN = CFG, int MINB_FB = MINB, class CFGLIN2 = void, int WPC = 1>
void run_batch_t(dyn_aligner& A, const BatchIO& io, int mode, BatchResult& res, uint32_t* sigpos_h, double* prob_h,
	double* pooled, double* per_read_w)
{
	Rt& rt = A.rt;
	rt.bind();
	HostTimer tm;
	if (A.table_dirty) A.upload_table();
	const uint32_t n = io.n;
	res.out.assign(n, ReadOut{});
	res.bad_pos.assign(n, 0xffffffffu);
	res.seg_off.assign((size_t)n + 1, 0);
	res.desc.assign(n, ReadDesc{});
	A.timing[0] = A.timing[1] = A.timing[2] = 0.0;
	if (n == 0) return;
	ReadOut* h_out_stage = nullptr;  // pinned landing zone of the per-read outputs
	// the previous call has synchronised its stream: the arena is free
	A.h_arena.get(rt, (size_t)n * (sizeof(ReadDesc) + 4 * 6 + 8 + 2 * sizeof(ReadOut) + 8) + ((size_t)8 << 20));
	A.arena_off = 0;
	h_out_stage = (ReadOut*)A.stage((size_t)n * sizeof(ReadOut));
	uint32_t* h_bad_stage = (uint32_t*)A.stage((size_t)n * 4);

	// ---- host: validation (aligner.cpp:145-164) and geometry (NT:240-247) ------------------------------------
	uint64_t pc_total = 0, seg_total = 0;
	uint32_t maxT = 0;
	std::vector<uint32_t> order;
	order.reserve(n);
	for (uint32_t r = 0; r < n; ++r)
	{
		ReadDesc& d = res.desc[r];
		const uint64_t S = io.sig_off[r + 1] - io.sig_off[r];
		const uint64_t L = io.seq_off[r + 1] - io.seq_off[r];
		d.sig_off = io.sig_off[r];
		d.status = ST_OK;
		res.seg_off[r] = seg_total;
		if (S < 1) d.status = ST_SIGNAL_EMPTY;
		else if (L < (uint64_t)A.k) d.status = ST_SEQ_SHORT;
		else if (S < 2 * (L - A.k + 1)) d.status = ST_SIGNAL_SHORT;
		else if (S >= 0x7fffff00ull) d.status = ST_INTERNAL;
		if (L >= (uint64_t)A.k) seg_total += L - A.k + 1;
		if (d.status != ST_OK)
		{
			d.S = 0; d.N = 0; d.bw = 0; d.ratio = 0; d.pc_off = pc_total; d.out_off = res.seg_off[r];
			continue;
		}
		const uint64_t Kc = L - A.k + 1;
		d.S = (uint32_t)S;
		d.N = (uint32_t)(Kc + 1);
		d.bw = (uint32_t)std::min<uint64_t>((uint64_t)A.band / 2, (uint64_t)d.N / 2);
		d.ratio = (double)d.N / (double)(S + 1);
		d.pc_off = pc_total;
		d.out_off = res.seg_off[r];
		pc_total += d.N;
		if (2 * (int)d.bw + 2 > CFG::SLOTS) d.status = ST_BAND_UNSUPPORTED;
		else
		{
			order.push_back(r);
			maxT = std::max(maxT, d.S + 1);
		}
	}
	res.seg_off[n] = seg_total;
	// longest first: the work queue then behaves like LPT scheduling
	std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
		return (uint64_t)res.desc[x].S * (2 * res.desc[x].bw + 1) > (uint64_t)res.desc[y].S * (2 * res.desc[y].bw + 1);
	});
	for (uint32_t r = 0; r < n; ++r)
		if (res.desc[r].status != ST_OK) res.out[r].status = res.desc[r].status;
	if (order.empty()) return;
	tm.lap("host_prep");

	// ---- device inputs --------------------------------------------------------------------------------------------
	const uint64_t sig_total = io.sig_off[n];
	const uint64_t seq_total = io.seq_off[n];
	const float* d_sig = io.sig_dev;
	std::vector<float> conv;
	if (!d_sig)
	{
		float* p = (float*)A.d_sig.get(rt, sig_total * sizeof(float));
		if (io.sig_host64)
		{
			// round to FP32 (what the DP consumes) on up to 8 host threads
			conv.resize(sig_total);
			const double* src = io.sig_host64;
			float* dstf = conv.data();
			const uint32_t parts = (uint32_t)std::min<uint64_t>(4096, sig_total / 65536 + 1);
			parallel_ranges(parts, sig_total, [&](uint32_t a, uint32_t b) {
				const uint64_t i0 = sig_total * a / parts, i1 = sig_total * b / parts;
				for (uint64_t i = i0; i < i1; ++i) dstf[i] = (float)src[i];
			});
			rt.h2d(p, conv.data(), sig_total * sizeof(float));
		}
		else
			rt.h2d(p, io.sig_host, sig_total * sizeof(float));
		d_sig = p;
	}
	const char* d_seq = io.seq_dev;
	if (!d_seq)
	{
		char* p = (char*)A.d_seq.get(rt, seq_total);
		rt.h2d(p, io.seq_host, seq_total);
		d_seq = p;
	}
	uint64_t* d_seqoff = (uint64_t*)A.d_seqoff.get(rt, ((size_t)n + 1) * 8);
	A.h2d_staged(d_seqoff, io.seq_off, ((size_t)n + 1) * 8);
	ReadDesc* d_desc = (ReadDesc*)A.d_desc.get(rt, (size_t)n * sizeof(ReadDesc));
	A.h2d_staged(d_desc, res.desc.data(), (size_t)n * sizeof(ReadDesc));
	uint32_t* d_order = (uint32_t*)A.d_order.get(rt, order.size() * 4);
	A.h2d_staged(d_order, order.data(), order.size() * 4);
	PosConst* d_pc = (PosConst*)A.d_pc.get(rt, pc_total * sizeof(PosConst));
	int32_t* d_kmers = (int32_t*)A.d_kmers.get(rt, pc_total * 4);
	uint32_t* d_bad = (uint32_t*)A.d_bad.get(rt, (size_t)n * 4);
	ReadOut* d_out = (ReadOut*)A.d_out.get(rt, (size_t)n * sizeof(ReadOut));
	uint32_t* d_queue = (uint32_t*)A.d_queue.get(rt, 64);
	uint32_t* d_sigpos = nullptr;
	double* d_prob = nullptr;
	if (mode == 1)
	{
		d_sigpos = (uint32_t*)A.d_sigpos.get(rt, seg_total * 4);
		d_prob = (double*)A.d_prob.get(rt, seg_total * 8);
	}

	// ---- K1: kmer encoding + emission constants ---------------------------------------------------------------
	EncodeArgs ea;
	ea.reads = d_desc; ea.n_reads = n; ea.seq = d_seq; ea.seq_off = d_seqoff; ea.k = A.k;
	ea.table = (const PosConst*)A.d_table.p; ea.pc = d_pc; ea.kmers = d_kmers; ea.bad_pos = d_bad;
	ea.out = d_out; ea.queue = d_queue;
	rt.mark(0);
	launch_encode(rt, ea);
	rt.mark(1);

	// ---- arguments shared by every DP launch of this batch ---------------------------------------------------
	BatchArgs ba;
	memset(&ba, 0, sizeof(ba));
	ba.reads = d_desc; ba.order = d_order; ba.n_reads = (uint32_t)order.size(); ba.queue = d_queue;
	ba.signal = d_sig; ba.pc = d_pc; ba.out = d_out;
	ba.out_sigpos = d_sigpos; ba.out_prob = d_prob;
	ba.m1 = (float)(A.trans[0] * LOG2E);
	ba.e2 = (float)(A.trans[2] * LOG2E);
	// full-band kernels: training statistics are gathered from the sparse posterior records, so the record threshold is
	// what truncates a kmer's weight: 2^-40 keeps every kmer above the documented weight threshold (1e-3) exact to
	// ~1e-9 relative (2^-22, the alignment threshold, cost 2e-3 on the trained stdev of light kmers)
	const double thr2_eff = (mode == 2) ? std::min(A.thr2, -40.0) : A.thr2;
	ba.thr2 = (float)thr2_eff;
	ba.m1_lin = (float)std::exp(A.trans[0]);
	ba.e2_lin = (float)std::exp(A.trans[2]);
	ba.thr_lin = (float)std::exp2(thr2_eff);
	ba.mode = mode;
	ba.uni_a = A.uni_a;
	ba.uni_c = A.uni_c;
	ba.fwd_fast = A.fwd_fast;
	ba.kmers = d_kmers;
	ba.thr_rib = (float)std::exp2(A.thr_rib);
	ba.rib_guard = A.rib_guard;
	if (mode == 2)
	{
		ba.read_w = (double*)A.d_rw.get(rt, pc_total * 8);
		ba.read_x = (double*)A.d_rx.get(rt, pc_total * 8);
		ba.read_xx = (double*)A.d_rxx.get(rt, pc_total * 8);
		rt.zero(ba.read_w, pc_total * 8);
		rt.zero(ba.read_x, pc_total * 8);
		rt.zero(ba.read_xx, pc_total * 8);
		if (A.ext_stats)
		{
			// pooled statistics stay on the device, in the caller's buffer (accumulated, not zeroed)
			ba.stat_w = A.ext_stats;
			ba.stat_x = A.ext_stats + A.K;
			ba.stat_xx = A.ext_stats + 2 * A.K;
		}
		else
		{
			ba.stat_w = (double*)A.d_sw.get(rt, A.K * 8);
			ba.stat_x = (double*)A.d_sx.get(rt, A.K * 8);
			ba.stat_xx = (double*)A.d_sxx.get(rt, A.K * 8);
			rt.zero(ba.stat_w, A.K * 8);
			rt.zero(ba.stat_x, A.K * 8);
			rt.zero(ba.stat_xx, A.K * 8);
		}
	}
	A.n_fallback = 0;
	A.n_retry_lin = 0;
	A.n_ribbon = 0;
	A.n_rib_fault = 0;
	A.n_rib_log = 0;
	A.n_rib_log_fault = 0;
	A.ribbon_ms = 0.0;
	int launches = 1;
	// reads the ribbon kernels lost COMPLETELY (every value of a group underflowed, or Zb is not finite): the same FP32
	// linear arithmetic over the full band cannot hold them either, they go straight to the log2-domain tier
	std::vector<uint32_t> direct_log2;

	// ---- tier 0: the ribbon kernels (dp_ribbon.cuh) -----------------------------------------------------------
	// every read whose reference band leaves room for the window; what they cannot represent (ST_LIN_FAULT) and the
	// short reads go on to the full-band kernels below
	rib::Geometry rg;
	if (A.arith == 0 && A.ribbon > 0 && rib::geometry(A.ribbon, A.rib_bps, rg))
	{
		std::vector<uint32_t> rorder, rest;
		uint32_t maxTr = 0;
		for (uint32_t r : order)
		{
			if ((int)res.desc[r].bw >= A.rib_min_bw)
			{
				rorder.push_back(r);
				maxTr = std::max(maxTr, res.desc[r].S + 1);
			}
			else rest.push_back(r);
		}
		if (!rorder.empty())
		{
			const size_t resident = (size_t)rt.sms * (A.warps_per_sm > 0 ? (size_t)A.warps_per_sm : (size_t)rg.blocks_per_sm * rg.warps_per_block);
			unsigned gridw = (unsigned)std::min<size_t>(resident, rorder.size());
			size_t per_slot = 0, o_sch = 0, o_ck = 0, o_ob = 0, o_hdr = 0, o_rec = 0, o_pp = 0;
			uint64_t rcap = 0;
			size_t o_ring = 0;
			bool two_level = false, gather = false;
			if (mode != 0)
			{
				const size_t budget = (size_t)((double)(rt.free_bytes() + (A.root ? A.root : &A)->d_rib_scratch.cap) * A.mem_fraction);
				// checkpoints of every group (640 B per 8 rows at 2 columns per lane) unless the scratch of the resident warps
				// would not fit: then only every 8th group's, the others replayed in pass 2 (long reads; A.rib_two_level forces)
				// ... and if that does not fit either (config 4: 2 M rows), the records-free layout: row header = the decision
				// words, no posterior records; the path posteriors come from a second forward sweep after the traceback
				// (23 instead of 71 bytes per lattice row)
				for (int pass = 0; pass < 3; ++pass)
				{
					gather = (mode == 1) && ((A.rib_gather > 0) || (A.rib_gather < 0 && pass == 2));
					two_level = gather || (A.rib_two_level > 0) || (A.rib_two_level < 0 && pass >= 1);
					const size_t ngr = (size_t)maxTr / rg.ck + 2;
					const size_t nck = two_level ? ngr / 8 + 2 : ngr;
					size_t o = 0;
					o_sch = o; o = align_up(o + ngr * 8, 256);
					o_ck = o; o = align_up(o + nck * rg.ckf * 4, 256);
					o_ob = o; o = align_up(o + nck * 32 * 4, 256);
					o_ring = o; o = align_up(o + (two_level ? 8 * ((size_t)rg.ckf * 4 + 128) : 0), 256);
					if (mode == 1 && gather)
					{
						rcap = 0;
						o_hdr = o; o = align_up(o + ((size_t)maxTr + 32) * rg.cpl * 4, 256);
						o_rec = o;
						o_pp = o; o = align_up(o + ((size_t)maxTr + 32) * 4, 256);
					}
					else if (mode == 1)
					{
						rcap = (uint64_t)std::min<double>((double)maxTr * A.rib_recs_per_row, (double)maxTr * 32.0) + 64 + (uint64_t)(rg.ck + 1) * 32;
						o_hdr = o; o = align_up(o + ((size_t)maxTr + 32) * rg.hdrw * 4, 256);
						o_rec = o; o = align_up(o + rcap * rg.recf * 4, 256);
						o_pp = o; o = align_up(o + ((size_t)maxTr + 32) * 4, 256);
					}
					per_slot = o;
					if (budget / per_slot >= gridw) break;  // every resident warp gets its scratch: take this layout
					// is a smaller layout allowed?
					if (pass == 0 && A.rib_two_level == 0 && A.rib_gather <= 0) break;
					if (pass == 0 && A.rib_two_level > 0 && A.rib_gather == 0) break;
					if (pass == 1 && (A.rib_gather == 0 || mode != 1)) break;
				}
				const size_t fit = std::max<size_t>(1, budget / per_slot);
				if (tm.on)
					fprintf(stderr, "[dyn timing] ribbon scratch: %.1f MB per resident warp x %u wanted, budget %.1f GB -> %zu fit%s%s\n",
						per_slot / 1048576.0, gridw, budget / 1073741824.0, fit, two_level ? " (two-level checkpoints)" : "",
						gather ? " (records-free, gather sweep)" : "");
				gridw = (unsigned)std::min<size_t>(gridw, fit);
			}
			{
			dyn_aligner& R = A.root ? *A.root : A;
			std::lock_guard<std::mutex> cl(R.compute_mu);
			std::vector<SlotScratch> slots(gridw);
			memset(slots.data(), 0, slots.size() * sizeof(SlotScratch));
			if (mode != 0)
			{
				if (per_slot * gridw > R.d_rib_scratch.cap)
				{
					// growing the shared pool: nothing enqueued may still use the old one (holding compute_mu keeps the
					// other lanes from enqueueing; cudaFree synchronises the device)
					R.rt.sync();
					rt.sync();
				}
				unsigned char* base = (unsigned char*)R.d_rib_scratch.get(R.rt, per_slot * gridw);
				for (unsigned q = 0; q < gridw; ++q)
				{
					unsigned char* b = base + per_slot * q;
					slots[q].sched = (uint2*)(b + o_sch);
					slots[q].ckpt = (float*)(b + o_ck);
					slots[q].ckpt_ob = (double*)(b + o_ob);
					slots[q].ring = (float*)(b + o_ring);
					if (mode == 1)
					{
						slots[q].hdr = (uint32_t*)(b + o_hdr);
						slots[q].recs = (void*)(b + o_rec);
						slots[q].pp = (float*)(b + o_pp);
					}
				}
			}
			SlotScratch* d_slots = (SlotScratch*)A.d_slots.get(rt, (size_t)gridw * sizeof(SlotScratch));
			A.h2d_staged(d_slots, slots.data(), (size_t)gridw * sizeof(SlotScratch));
			A.h2d_staged(d_order, rorder.data(), rorder.size() * 4);
			BatchArgs rb = ba;
			rb.n_reads = (uint32_t)rorder.size();
			rb.slots = d_slots;
			rb.n_slots = gridw;
			rb.rec_cap = rcap;
			tm.lap("enqueue_ribbon");
			// inputs ready on this handle's stream -> kernel on the root's stream -> results back on this handle's stream
			const bool cross = (&R != &A);
			if (cross)
			{
				rt.mark(8);
				rt.wait_on(R.rt, 8);
			}
			rt.mark_on(6, R.rt);
#ifndef DYN_HOST_EMU
			const int le = rib::launch((void*)R.rt.stream, rb, gridw, gather ? 3 : mode, A.ribbon, A.rib_bps, two_level);
			if (le != 0) throw std::runtime_error(std::string("CUDA error launching the ribbon kernel: ") + cudaGetErrorString((cudaError_t)le));
#else
			rib::launch(nullptr, rb, gridw, gather ? 3 : mode, A.ribbon, A.rib_bps, two_level);
#endif
			rt.mark_on(7, R.rt);
			if (cross) rt.wait_self(7);
			++launches;
			}
			rt.d2h(h_out_stage, d_out, (size_t)n * sizeof(ReadOut));
			rt.sync();
			std::memcpy(res.out.data(), h_out_stage, (size_t)n * sizeof(ReadOut));
			tm.lap("ribbon_kernel");
			A.ribbon_ms = rt.elapsed(6, 7);
			A.n_ribbon = rorder.size();
			{
				double nrec = 0.0, nrow = 0.0;
				for (uint32_t r : rorder)
					if (res.out[r].status == ST_OK)
					{
						nrec += (double)res.out[r].nrec;
						nrow += (double)res.desc[r].S;
					}
				A.rib_last_two_level = gather ? 2.0 : (two_level ? 1.0 : 0.0);
				A.rib_recs_used = nrow > 0 ? nrec / nrow : 0.0;
			}
			// ---- tier 0b: the log2-domain ribbon (RCfg<2, true>) for the reads whose loss is one of RANGE — everything
			// underflowed (an alignment forced along the band's edge multiplies by 2^-600 per row), the posterior mass or the
			// Viterbi scores died, Zf != Zb, the range guard tripped.  Same window, same passes, unlimited range; the reads whose
			// mass sits at the window's EDGE (reasons 2, 7: a window problem, not a range problem) go to the full-band kernels.
			std::vector<uint32_t> logq;
			const bool force_log = (A.rib_log == 2);  // test hook: every read of the tier through the log2-domain ribbon as well
			if ((mode == 1 || mode == 2) && A.rib_log != 0)
				for (uint32_t r : rorder)
				{
					if (res.out[r].status == ST_LIN_FAULT)
					{
						const uint32_t why = res.out[r].nrec;
						if (force_log || why == 1u || why == 5u || why == 6u || why == 8u || why == 9u || why == 12u) logq.push_back(r);
					}
					else if (force_log && res.out[r].status == ST_OK) logq.push_back(r);
				}
			std::vector<uint32_t> first_why(logq.size());
			if (!logq.empty())
			{
				dyn_aligner& R = A.root ? *A.root : A;
				uint32_t maxTl = 0;
				for (size_t i = 0; i < logq.size(); ++i)
				{
					maxTl = std::max(maxTl, res.desc[logq[i]].S + 1);
					first_why[i] = (res.out[logq[i]].status == ST_LIN_FAULT) ? res.out[logq[i]].nrec : 0u;
				}
				// records-free layout with two-level checkpoints (the only one the log2-domain kernel is built for)
				const size_t ngr = (size_t)maxTl / rib::LOG_GROUP_ROWS + 2, nck = ngr / 8 + 2;
				size_t o = 0;
				const size_t l_sch = o; o = align_up(o + ngr * 8, 256);
				const size_t l_ck = o; o = align_up(o + nck * rg.ckf * 4, 256);
				const size_t l_ob = o; o = align_up(o + nck * 32 * 4, 256);
				const size_t l_ring = o; o = align_up(o + 8 * ((size_t)rg.ckf * 4 + 128), 256);
				const size_t l_hdr = o; o = align_up(o + ((size_t)maxTl + 32) * rg.cpl * 4, 256);
				const size_t l_pp = o; o = align_up(o + ((size_t)maxTl + 32) * 4, 256);
				const size_t per_l = o;
				std::lock_guard<std::mutex> cl(R.compute_mu);
				const size_t budget = (size_t)((double)(rt.free_bytes() + R.d_rib_scratch.cap) * A.mem_fraction);
				const unsigned gridl = (unsigned)std::max<size_t>(1, std::min<size_t>(logq.size(), budget / per_l));
				if (per_l * gridl > R.d_rib_scratch.cap)
				{
					R.rt.sync();
					rt.sync();
				}
				unsigned char* base = (unsigned char*)R.d_rib_scratch.get(R.rt, per_l * gridl);
				std::vector<SlotScratch> slots(gridl);
				memset(slots.data(), 0, slots.size() * sizeof(SlotScratch));
				for (unsigned q = 0; q < gridl; ++q)
				{
					unsigned char* b = base + per_l * q;
					slots[q].sched = (uint2*)(b + l_sch);
					slots[q].ckpt = (float*)(b + l_ck);
					slots[q].ckpt_ob = (double*)(b + l_ob);
					slots[q].ring = (float*)(b + l_ring);
					slots[q].hdr = (uint32_t*)(b + l_hdr);
					slots[q].pp = (float*)(b + l_pp);
				}
				SlotScratch* d_slots = (SlotScratch*)A.d_slots.get(rt, (size_t)gridl * sizeof(SlotScratch));
				A.h2d_staged(d_slots, slots.data(), (size_t)gridl * sizeof(SlotScratch));
				A.h2d_staged(d_order, logq.data(), logq.size() * 4);
				rt.zero(d_queue, 64);
				if (mode == 2)
					for (uint32_t r : logq)
					{
						// columns the read flushed before its fault was detected: the re-run accumulates from zero
						const ReadDesc& d = res.desc[r];
						rt.zero(ba.read_w + d.pc_off, (size_t)d.N * 8);
						rt.zero(ba.read_x + d.pc_off, (size_t)d.N * 8);
						rt.zero(ba.read_xx + d.pc_off, (size_t)d.N * 8);
					}
				BatchArgs rb = ba;
				rb.n_reads = (uint32_t)logq.size();
				rb.slots = d_slots;
				rb.n_slots = gridl;
				rb.rec_cap = 0;
				const bool cross = (&R != &A);
				if (cross)
				{
					rt.mark(8);
					rt.wait_on(R.rt, 8);
				}
				rt.mark_on(10, R.rt);
#ifndef DYN_HOST_EMU
				const int le = rib::launch((void*)R.rt.stream, rb, gridl, mode == 2 ? 2 : 3, 2, 0, true, true);
				if (le != 0) throw std::runtime_error(std::string("CUDA error launching the log2-domain ribbon kernel: ") + cudaGetErrorString((cudaError_t)le));
#else
				rib::launch(nullptr, rb, gridl, mode == 2 ? 2 : 3, 2, 0, true, true);
#endif
				rt.mark_on(11, R.rt);
				if (cross) rt.wait_self(11);
				++launches;
				rt.d2h(h_out_stage, d_out, (size_t)n * sizeof(ReadOut));
				rt.sync();
				std::memcpy(res.out.data(), h_out_stage, (size_t)n * sizeof(ReadOut));
				A.ribbon_ms += rt.elapsed(10, 11);
				A.n_rib_log = logq.size();
				tm.lap("log_ribbon_kernel");
			}
			{
				// a read the log2-domain ribbon kept is no longer a fault of the tier; one it lost too keeps its FIRST reason
				size_t i = 0;
				for (uint32_t r : logq)
				{
					if (first_why[i] != 0u)
					{
						++A.rib_reason[std::min<uint32_t>(first_why[i], 15u)];
						++A.n_rib_fault;
					}
					if (res.out[r].status == ST_LIN_FAULT)
					{
						++A.n_rib_log_fault;
						direct_log2.push_back(r);  // the linear full-band tiers would lose it for the same reason
						if (mode == 2)
						{
							const ReadDesc& d = res.desc[r];
							rt.zero(ba.read_w + d.pc_off, (size_t)d.N * 8);
							rt.zero(ba.read_x + d.pc_off, (size_t)d.N * 8);
							rt.zero(ba.read_xx + d.pc_off, (size_t)d.N * 8);
						}
					}
					else ++A.rib_reason[15];  // [15]: cumulative reads the log2-domain ribbon kept
					++i;
				}
			}
			for (uint32_t r : rorder)
				if (res.out[r].status == ST_LIN_FAULT && std::find(logq.begin(), logq.end(), r) == logq.end())
				{
					const uint32_t why = res.out[r].nrec;
					// ... and so do long reads whatever the reason: a full-band tier is one warp for seconds per read
					// (config 4: ~3 s per tier for 2 M rows), not worth two attempts that usually fail for the same reason
					if (why == 1u || why == 3u || why == 12u || res.desc[r].S > 400000u) direct_log2.push_back(r);
					else rest.push_back(r);
					++A.n_rib_fault;
					++A.rib_reason[std::min<uint32_t>(why, 15u)];
					if (mode == 2)
					{
						// columns the read flushed before its fault was detected: the full-band kernels accumulate
						const ReadDesc& d = res.desc[r];
						rt.zero(ba.read_w + d.pc_off, (size_t)d.N * 8);
						rt.zero(ba.read_x + d.pc_off, (size_t)d.N * 8);
						rt.zero(ba.read_xx + d.pc_off, (size_t)d.N * 8);
					}
				}
			std::stable_sort(rest.begin(), rest.end(), [&](uint32_t x, uint32_t y) {
				return (uint64_t)res.desc[x].S * (2 * res.desc[x].bw + 1) > (uint64_t)res.desc[y].S * (2 * res.desc[y].bw + 1);
			});
			order.swap(rest);
			maxT = 0;
			for (uint32_t r : order) maxT = std::max(maxT, res.desc[r].S + 1);
			for (uint32_t r : direct_log2) maxT = std::max(maxT, res.desc[r].S + 1);
			if (!order.empty())
			{
				A.h2d_staged(d_order, order.data(), order.size() * 4);
				rt.zero(d_queue, 64);
			}
			ba.n_reads = (uint32_t)order.size();
		}
	}

	double fallback_ms = 0.0, main_ms = 0.0;
	if (!order.empty() || !direct_log2.empty())
	{
	const bool lin_first = (A.arith == 0);
	// ---- scratch: one slot per resident warp, sized for the longest read --------------------------------------
	// Z-only runs the backward pass alone: <= 128 registers and no shared memory, i.e. 16 single-warp CTAs per SM
	const int resident = (mode == 0) ? 16 : MINB;
	unsigned grid = (unsigned)std::min<size_t>((size_t)rt.sms * (A.warps_per_sm > 0 ? A.warps_per_sm : resident), order.size() + direct_log2.size());
	size_t per_slot = 0;
	uint64_t rec_cap = 0;
	size_t o_ck = 0, o_ob = 0, o_bits = 0, o_rp = 0, o_rec = 0, o_pn = 0, o_pp = 0;
	if (mode != 0)
	{
		const size_t nck = (size_t)maxT / CFG::CK + 2;
		rec_cap = (uint64_t)std::min<double>((double)maxT * A.recs_per_row, (double)maxT * 32.0) + 64;
		size_t o = 0;
		o_ck = o; o = align_up(o + nck * CFG::CKF * 4, 256);
		o_ob = o; o = align_up(o + nck * 32 * 8, 256);
		o_bits = o; o = align_up(o + ((size_t)maxT + 32) * 64, 256);
		o_rp = o; o = align_up(o + ((size_t)maxT + 2) * 4, 256);
		o_rec = o; o = align_up(o + rec_cap * sizeof(LaneRec<CFG::CPL>), 256);
		o_pn = o; o = align_up(o + ((size_t)maxT + 1) * 4, 256);
		o_pp = o; o = align_up(o + ((size_t)maxT + 1) * 4, 256);
		per_slot = o;
		size_t budget = (size_t)((double)(rt.free_bytes() + A.d_scratch.cap) * A.mem_fraction);
		size_t fit = std::max<size_t>(1, budget / per_slot);
		if (fit < (grid + 3) / 4)
		{
			// long reads handed on by the ribbon kernels: their full-band scratch needs the room of the (shared) ribbon pool
			// — taken only when the reads would otherwise need more than four rounds: giving the pool back drains the root's
			// stream (the other lanes' ribbon kernels) and costs its re-allocation, a second or third round of a few warps
			// runs beside those kernels.
			// Nothing enqueued may still use it: hold compute_mu (no lane can enqueue) and drain the root's stream.
			dyn_aligner& R = A.root ? *A.root : A;
			std::lock_guard<std::mutex> cl(R.compute_mu);
			if (R.d_rib_scratch.cap)
			{
				R.rt.sync();
				rt.sync();
				R.d_rib_scratch.release(R.rt);
				budget = (size_t)((double)(rt.free_bytes() + A.d_scratch.cap) * A.mem_fraction);
				fit = std::max<size_t>(1, budget / per_slot);
			}
		}
		if (tm.on)
			fprintf(stderr, "[dyn timing] scratch: %.1f MB per resident warp x %u wanted, budget %.1f GB -> %zu fit\n",
				per_slot / 1048576.0, grid, budget / 1073741824.0, fit);
		grid = (unsigned)std::min<size_t>(grid, fit);
	}
	if (WPC > 1 && lin_first) grid = std::max<unsigned>(grid / WPC, 1u) * WPC;  // whole CTAs of WPC warps
	std::vector<SlotScratch> slots(grid);
	memset(slots.data(), 0, slots.size() * sizeof(SlotScratch));
	if (mode != 0)
	{
		unsigned char* base = (unsigned char*)A.d_scratch.get(rt, per_slot * grid);
		for (unsigned s = 0; s < grid; ++s)
		{
			unsigned char* b = base + per_slot * s;
			slots[s].ckpt = (float*)(b + o_ck);
			slots[s].ckpt_ob = (double*)(b + o_ob);
			slots[s].bits = (uint16_t*)(b + o_bits);
			slots[s].rowptr = (uint32_t*)(b + o_rp);
			slots[s].recs = (void*)(b + o_rec);
			slots[s].pn = (uint32_t*)(b + o_pn);
			slots[s].pp = (float*)(b + o_pp);
		}
	}
	SlotScratch* d_slots = (SlotScratch*)A.d_slots.get(rt, (size_t)grid * sizeof(SlotScratch));
	A.h2d_staged(d_slots, slots.data(), (size_t)grid * sizeof(SlotScratch));
	ba.slots = d_slots;
	ba.n_slots = grid;
	ba.rec_cap = rec_cap;

	// ---- K2..K5: the full-band DP kernel ----------------------------------------------------------------------------
	const bool lin = (A.arith == 0);
	tm.lap("enqueue");
	rt.mark(2);
	if (!order.empty())
	{
		launch_align<CFG, MINB, CFGLIN, MINB_FB, WPC>(rt, ba, grid, mode, lin);
		++launches;
	}
	rt.mark(3);
	if (lin)
	{
		// reads the FP32 linear arithmetic could not represent (ST_LIN_FAULT) are re-run in the log2 domain
		rt.d2h(h_out_stage, d_out, (size_t)n * sizeof(ReadOut));
		rt.sync();
		std::memcpy(res.out.data(), h_out_stage, (size_t)n * sizeof(ReadOut));
		tm.lap("kernels");
		std::vector<uint32_t> again;
		for (uint32_t r : order)
			if (res.out[r].status == ST_LIN_FAULT) again.push_back(r);
		if constexpr (!std::is_void<CFGLIN2>::value)
		{
			if (!again.empty())
			{
				// second tier: the linear-domain kernels that renormalise twice as often
				A.n_retry_lin = again.size();
				A.h2d_staged(d_order, again.data(), again.size() * 4);
				rt.zero(d_queue, 64);
				ba.n_reads = (uint32_t)again.size();
				rt.mark(4);
				launch_align_t<CFGLIN2, MINB_FB, true>(rt, ba, (unsigned)std::min<size_t>(grid, again.size()), mode);
				rt.mark(5);
				rt.d2h(h_out_stage, d_out, (size_t)n * sizeof(ReadOut));
				rt.sync();
				std::memcpy(res.out.data(), h_out_stage, (size_t)n * sizeof(ReadOut));
				fallback_ms += rt.elapsed(4, 5);
				++launches;
				std::vector<uint32_t> still;
				for (uint32_t r : again)
					if (res.out[r].status == ST_LIN_FAULT) still.push_back(r);
				again.swap(still);
			}
		}
		again.insert(again.end(), direct_log2.begin(), direct_log2.end());
		if (!again.empty())
		{
			A.n_fallback = again.size();
			A.h2d_staged(d_order, again.data(), again.size() * 4);
			rt.zero(d_queue, 64);
			ba.n_reads = (uint32_t)again.size();
			rt.mark(4);
			launch_align<CFG, MINB, CFGLIN, MINB_FB, WPC>(rt, ba, (unsigned)std::min<size_t>(grid, again.size()), mode, false);
			rt.mark(5);
			rt.sync();
			fallback_ms += rt.elapsed(4, 5);
			++launches;
		}
	}
	rt.sync();
	main_ms = rt.elapsed(2, 3);
	}
	if (mode == 2)
	{
		FoldArgs fa;
		fa.reads = d_desc; fa.out = d_out; fa.n_reads = n; fa.kmers = d_kmers;
		fa.read_w = ba.read_w; fa.read_x = ba.read_x; fa.read_xx = ba.read_xx;
		fa.stat_w = ba.stat_w; fa.stat_x = ba.stat_x; fa.stat_xx = ba.stat_xx;
		launch_fold(rt, fa);
		++launches;
		if (A.ext_stats)
		{
			SumOutArgs sa;
			sa.out = d_out; sa.n_reads = n; sa.tail = A.ext_stats + 3 * A.K;
#ifndef DYN_HOST_EMU
			k_sum_out<<<std::min<unsigned>((n + 31) / 32, 256u), 32, 0, rt.stream>>>(sa);
			CK_CUDA(cudaGetLastError());
#else
			simt::launch(1, 0, [&]() { sum_out_lane(sa, 0u, 32u, threadIdx.x); });
#endif
			++launches;
		}
	}

	// ---- results ---------------------------------------------------------------------------------------------------
	rt.d2h(h_out_stage, d_out, (size_t)n * sizeof(ReadOut));
	rt.d2h(h_bad_stage, d_bad, (size_t)n * 4);
	if (mode == 1 && seg_total)
	{
		rt.d2h(sigpos_h, d_sigpos, seg_total * 4);
		rt.d2h(prob_h, d_prob, seg_total * 8);
	}
	if (mode == 2 && pooled)
	{
		rt.d2h(pooled, ba.stat_w, A.K * 8);
		rt.d2h(pooled + A.K, ba.stat_x, A.K * 8);
		rt.d2h(pooled + 2 * A.K, ba.stat_xx, A.K * 8);
	}
	if (mode == 2 && per_read_w)
	{
		rt.d2h(per_read_w, ba.read_w, pc_total * 8);
		rt.d2h(per_read_w + pc_total, ba.read_x, pc_total * 8);
		rt.d2h(per_read_w + 2 * pc_total, ba.read_xx, pc_total * 8);
	}
	rt.sync();
	std::memcpy(res.out.data(), h_out_stage, (size_t)n * sizeof(ReadOut));
	std::memcpy(res.bad_pos.data(), h_bad_stage, (size_t)n * 4);
	tm.lap("results_d2h");
	A.timing[0] = rt.elapsed(0, 1);
	A.timing[1] = A.ribbon_ms + main_ms + fallback_ms;
	A.timing[2] = launches;
	for (uint32_t r = 0; r < n; ++r)
	{
		if (res.desc[r].status != ST_OK && res.desc[r].status != ST_INVALID_NT) res.out[r].status = res.desc[r].status;
		if (res.bad_pos[r] != 0xffffffffu) res.out[r].status = ST_INVALID_NT;
	}
}

void run_batch_variant(dyn_aligner& A, const BatchIO& io, int mode, BatchResult& res, uint32_t* sigpos_h, double* prob_h,
	double* pooled, double* per_read_w);

void run_batch(dyn_aligner& A, const BatchIO& io, int mode, BatchResult& res, uint32_t* sigpos_h, double* prob_h,
	double* pooled, double* per_read_w)
{
#ifndef DYN_HOST_EMU
	try
	{
		run_batch_variant(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w);
	}
	catch (const DevOom&)
	{
		// The ribbon scratch pool is sized from the memory that is free when a batch needs it (92 %); buffers allocated
		// later (another lane's inputs, a larger batch) can then fail.  Give the pool back — nothing enqueued may still use
		// it: compute_mu keeps the lanes from enqueueing, the root's stream is drained — and run the batch once more; the
		// pool is re-sized from what is free then.
		dyn_aligner& R = A.root ? *A.root : A;
		{
			std::lock_guard<std::mutex> cl(R.compute_mu);
			R.rt.sync();
			A.rt.sync();
			R.d_rib_scratch.release(R.rt);
			A.d_scratch.release(A.rt);
		}
		run_batch_variant(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w);
	}
#else
	run_batch_variant(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w);
#endif
}

void run_batch_variant(dyn_aligner& A, const BatchIO& io, int mode, BatchResult& res, uint32_t* sigpos_h, double* prob_h,
	double* pooled, double* per_read_w)
{
	int v = A.variant;
	if (v < 0) v = A.uniform ? DEFAULT_VARIANT_UNI : DEFAULT_VARIANT;
	if (v >= 4 && v <= 8 && !A.uniform)
	{
		static const int general[5] = {3, 9, 1, 1, 2};
		v = general[v - 4];
	}
	if (v == 10 && !A.uniform) v = 11;
	if (v == 12 && !A.uniform) v = 13;
	A.last_variant = v;
	switch (v)
	{
#if DYN_HAS(4)
	case 4: run_batch_t<Cfg8, 8, Cfg8U, 8>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(5)
	case 5: run_batch_t<Cfg8, 9, Cfg8U, 8>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(6)
	case 6: run_batch_t<Cfg8, 10, Cfg8U, 8>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(7)
	case 7: run_batch_t<Cfg8, 11, Cfg8U, 8>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(8)
	case 8: run_batch_t<Cfg8, 12, Cfg8U, 8>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(10)
	case 10: run_batch_t<Cfg8, DYN_V10_MINB, Cfg8UR8, 8, Cfg8U>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(12)
	case 12: run_batch_t<Cfg8, DYN_V12_WPC, Cfg8UR8, 8, Cfg8U, DYN_V12_WPC>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(13)
	case 13: run_batch_t<Cfg8, DYN_V12_WPC, Cfg8R8, 8, Cfg8, DYN_V12_WPC>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(11)
	case 11: run_batch_t<Cfg8, 8, Cfg8R8, 8, Cfg8>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(9)
	case 9: run_batch_t<Cfg8, 9, Cfg8, 8>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(1)
	case 1: run_batch_t<Cfg8, 10>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(2)
	case 2: run_batch_t<Cfg8, 12>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(3)
	case 3: run_batch_t<Cfg8, 8>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#endif
#if DYN_HAS(0)
	default: run_batch_t<Cfg16, 7>(A, io, mode, res, sigpos_h, prob_h, pooled, per_read_w); break;
#else
	default: throw std::runtime_error("dynamont_b200: kernel variant not built (DYN_ONLY_VARIANT)");
#endif
	}
}

} // namespace

// ---------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------
extern "C"
{

dyn_aligner* dyn_create(const char* model_path, const char* pore, const char* mode, int threads, int band,
	int device, char* err, size_t errlen, int* err_kind)
{
	(void)threads;  // accepted and unused, like the reference
	if (err_kind) *err_kind = 0;
	dyn_aligner* A = nullptr;
	try
	{
		const PoreInfo* pi = nullptr;
		for (const PoreInfo& p : PORES)
			if (std::string(p.name) == pore) pi = &p;
		if (!pi)
		{
			if (err_kind) *err_kind = 1;
			throw std::invalid_argument(std::string("Unknown pore type: ") + pore);  // aligner_bindings.cpp:31
		}
		const std::string m(mode ? mode : "basic");
		const bool ntk = (m == "resquiggle" || m == "ntk");
		if (!ntk && m != "basic" && m != "nt")
		{
			if (err_kind) *err_kind = 1;
			throw std::invalid_argument("Unknown aligner mode: " + m);  // aligner_bindings.cpp:50
		}
		A = new dyn_aligner();
		A->ntk = ntk;
		A->model_path_ = model_path;
		A->pore_ = pore;
		A->mode_ = m;
		{
			// NTKAligner::initializeTransitions (NTK_aligner_api.cpp:35-104)
			static const double RNA002_T[14] = {0.019326040280789637, 0.19725479693713352, 0.1979799841413514,
				0.0006135538271005425, 0.7669801909288386, 0.27034500789657623, 0.00032463686748883153,
				0.02916688206070035, 1.0, 0.7296549921055607, 0.8020200158564497, 0.9797333838008437,
				2.3852272324574183e-06, 0.006598130068516047};
			static const double RNA004_T[14] = {0.029709838889618322, 0.2837864344979079, 0.15353628902814298,
				0.0041495012884881655, 0.47456322874771467, 0.05012685122100474, 0.0006112333189296363,
				0.13506593503589423, 1.0, 0.949873148779652, 0.8464637109688202, 0.9654529072452087,
				7.651926003806137e-05, 0.10658440170772512};
			static const double ONES_T[14] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
			const std::string pn(pi->name);
			const double* tv = (pn == "rna002") ? RNA002_T : (pn == "dna_r9") ? ONES_T : RNA004_T;
			for (int i = 0; i < 14; ++i) A->ntk_trans[i] = std::log(tv[i]);
			A->ntk_trans[14] = A->ntk_trans[16] = std::log(pi->m1);
			A->ntk_trans[15] = A->ntk_trans[17] = std::log(pi->e2);
		}
		A->rna = pi->rna;
		A->k = pi->k;
		A->band = band;
		A->trans[0] = std::log(pi->m1);
		A->trans[1] = std::log(pi->e1);
		A->trans[2] = std::log(pi->e2);
		load_model(*A, model_path);
		A->rt.init(device);
		A->device_ = A->rt.device;
		A->upload_table();
		return A;
	}
	catch (const std::exception& e)
	{
		if (err && errlen)
		{
			std::strncpy(err, e.what(), errlen - 1);
			err[errlen - 1] = 0;
		}
		delete A;
		return nullptr;
	}
}

void dyn_destroy(dyn_aligner* A)
{
	if (!A) return;
	for (dyn_aligner::Job* j : A->jobs)
	{
		if (!j->joined && j->th.joinable()) j->th.join();
		delete j;
	}
	A->jobs.clear();
	for (dyn_aligner*& l : A->lane)
	{
		if (l) dyn_destroy(l);
		l = nullptr;
	}
	try
	{
		A->rt.bind();
		for (DevBuf* b : {&A->d_table, &A->d_sig, &A->d_seq, &A->d_seqoff, &A->d_desc, &A->d_order, &A->d_pc, &A->d_kmers,
				 &A->d_bad, &A->d_out, &A->d_sigpos, &A->d_prob, &A->d_scratch, &A->d_slots, &A->d_queue, &A->d_rw,
				 &A->d_rx, &A->d_rxx, &A->d_sw, &A->d_sx, &A->d_sxx, &A->d_rib_scratch})
			b->release(A->rt);
		A->h_sigpos.release(A->rt);
		A->h_prob.release(A->rt);
#ifndef DYN_HOST_EMU
		if (A->ntk_pool_used)
		{
			cudaMemPool_t pool_h;
			if (cudaDeviceGetDefaultMemPool(&pool_h, A->rt.device) == cudaSuccess)
			{
				cudaDeviceSynchronize();
				cudaMemPoolTrimTo(pool_h, 0);
			}
		}
#endif
		A->rt.fini();
	}
	catch (...)
	{
	}
	delete A;
}

int dyn_kmer_size(const dyn_aligner* A) { return A->k; }
uint64_t dyn_num_kmers(const dyn_aligner* A) { return A->K; }
int dyn_is_rna(const dyn_aligner* A) { return A->rna ? 1 : 0; }

void dyn_model(const dyn_aligner* A, double* mean, double* stdev)
{
	std::memcpy(mean, A->mean.data(), A->K * 8);
	std::memcpy(stdev, A->stdev.data(), A->K * 8);
}

int dyn_set_model(dyn_aligner* A, const double* mean, const double* stdev)
{
	std::lock_guard<std::mutex> g(A->mu);
	A->mean.assign(mean, mean + A->K);
	A->stdev.assign(stdev, stdev + A->K);
	A->table_dirty = true;
	for (dyn_aligner* l : A->lane)
		if (l) dyn_set_model(l, mean, stdev);
	return 0;
}

void dyn_transitions(const dyn_aligner* A, double* log3)
{
	log3[0] = A->trans[0];
	log3[1] = A->trans[1];
	log3[2] = A->trans[2];
}

uint64_t dyn_count_segments(const dyn_aligner* A, const uint64_t* seq_off, uint32_t n_reads)
{
	uint64_t s = 0;
	for (uint32_t r = 0; r < n_reads; ++r)
	{
		const uint64_t L = seq_off[r + 1] - seq_off[r];
		if (L >= (uint64_t)A->k) s += L - A->k + 1;
	}
	return s;
}

// smallest row t in [0, T] with (size_t)(t * ratio) >= m   (band centre of row t, NT:100)
static uint64_t first_row_with_mid(uint64_t m, double ratio, uint64_t T)
{
	if (m == 0) return 0;
	uint64_t t = (uint64_t)std::ceil((double)m / ratio);
	if (t > T) t = T;
	while (t > 0 && (uint64_t)((double)(t - 1) * ratio) >= m) --t;
	while (t < T && (uint64_t)((double)t * ratio) < m) ++t;
	return t;
}

uint64_t dyn_read_cells(const dyn_aligner* A, uint64_t S, uint64_t L)
{
	if (S < 1 || L < (uint64_t)A->k) return 0;
	const uint64_t T = S + 1, N = L - A->k + 2;
	const uint64_t bw = std::min<uint64_t>((uint64_t)A->band / 2, N / 2);
	const double ratio = (double)N / (double)T;
	// rows 1..T-1 grouped by their band centre m: width(m) = min(m+bw+1, N) - max(m-bw, 1)   (NT:122-141)
	uint64_t cells = 0;
	uint64_t t0 = std::max<uint64_t>(first_row_with_mid(0, ratio, T), 1);
	for (uint64_t m = 0; m < N && t0 < T; ++m)
	{
		const uint64_t t1 = std::max<uint64_t>(first_row_with_mid(m + 1, ratio, T), 1);
		const uint64_t lo = std::max<uint64_t>(m >= bw ? m - bw : 0, 1);
		const uint64_t hi = std::min<uint64_t>(m + bw + 1, N);
		if (t1 > t0 && hi > lo) cells += (t1 - t0) * (hi - lo);
		t0 = t1;
	}
	return cells;
}

uint64_t dyn_batch_cells(const dyn_aligner* A, const uint64_t* sig_off, const uint64_t* seq_off, uint32_t n_reads,
	uint64_t* per_read)
{
	uint64_t total = 0;
	for (uint32_t r = 0; r < n_reads; ++r)
	{
		const uint64_t S = sig_off[r + 1] - sig_off[r], L = seq_off[r + 1] - seq_off[r];
		uint64_t c = 0;
		if (S >= 1 && L >= (uint64_t)A->k && S >= 2 * (L - A->k + 1)) c = dyn_read_cells(A, S, L);
		if (per_read) per_read[r] = c;
		total += c;
	}
	return total;
}

static const char* NTK_PARTIAL = "dynamont_b200: resquiggle (NTK) mode handles are served by dyn_ntk_align (one read per call); the batched entry points and training are basic mode only";

static int align_common(dyn_aligner* A, const BatchIO& io, int calc_probabilities, dyn_read_result* results,
	uint64_t* sequence_positions, uint64_t* signal_positions, double* probabilities)
{
	std::lock_guard<std::mutex> g(A->mu);
	try
	{
		if (A->ntk) throw std::runtime_error(NTK_PARTIAL);
		HostTimer tm;
		BatchResult res;
		const uint64_t seg_total = dyn_count_segments(A, io.seq_off, io.n);
		// device results land in pinned staging buffers and are fanned out into the caller's arrays below
		A->rt.bind();
		uint32_t* sigpos = (uint32_t*)A->h_sigpos.get(A->rt, (calc_probabilities ? seg_total : 0) * 4 + 4);
		double* prob_st = (double*)A->h_prob.get(A->rt, (calc_probabilities ? seg_total : 0) * 8 + 8);
		tm.lap("setup");
		run_batch(*A, io, calc_probabilities ? 1 : 0, res, sigpos, prob_st, nullptr, nullptr);
		tm.lap("run_batch");

		// retry reads whose sparse posterior buffer overflowed, one by one with a full-size buffer
		std::vector<uint32_t> retry;
		for (uint32_t r = 0; r < io.n; ++r)
			if (res.out[r].status == ST_REC_OVERFLOW) retry.push_back(r);
		if (!retry.empty())
		{
			// restores the record budget and the per-batch counters however the single-read retries end
			struct Restore
			{
				dyn_aligner* A;
				double recs;
				uint64_t fb, rl, nr, nf;
				double t0, t1, t2;
				int lv;
				~Restore()
				{
					A->recs_per_row = recs;
					A->n_fallback = fb; A->n_retry_lin = rl; A->n_ribbon = nr; A->n_rib_fault = nf;
					A->timing[0] = t0; A->timing[1] = t1; A->timing[2] = t2;
					A->last_variant = lv;
				}
			} restore{A, A->recs_per_row, A->n_fallback, A->n_retry_lin, A->n_ribbon, A->n_rib_fault, A->timing[0], A->timing[1],
				A->timing[2], A->last_variant};
			A->recs_per_row = 1e9;  // clamped to the band width inside run_batch
			for (uint32_t r : retry)
			{
				BatchIO one = io;
				one.n = 1;
				uint64_t so[2] = {io.sig_off[r], io.sig_off[r + 1]};
				uint64_t qo[2] = {io.seq_off[r], io.seq_off[r + 1]};
				// offsets are absolute into the caller's arrays, which run_batch indexes from sig_off[0]
				one.sig_off = so;
				one.seq_off = qo;
				BatchResult r1;
				const uint64_t kc = res.seg_off[r + 1] - res.seg_off[r];
				std::vector<uint32_t> sp(kc);
				std::vector<double> pr(kc);
				// single-read batches upload [0, sig_off[1]) — shift the base pointers instead
				BatchIO sh = one;
				uint64_t so0[2] = {0, so[1] - so[0]};
				uint64_t qo0[2] = {0, qo[1] - qo[0]};
				sh.sig_off = so0;
				sh.seq_off = qo0;
				if (io.sig_host) sh.sig_host = io.sig_host + so[0];
				if (io.sig_host64) sh.sig_host64 = io.sig_host64 + so[0];
				if (io.sig_dev) sh.sig_dev = io.sig_dev + so[0];
				if (io.seq_host) sh.seq_host = io.seq_host + qo[0];
				if (io.seq_dev) sh.seq_dev = io.seq_dev + qo[0];
				run_batch(*A, sh, 1, r1, sp.data(), pr.data(), nullptr, nullptr);
				res.out[r] = r1.out[0];
				std::copy(sp.begin(), sp.end(), sigpos + res.seg_off[r]);
				std::copy(pr.begin(), pr.end(), prob_st + res.seg_off[r]);
			}
		}

		for (uint32_t r = 0; r < io.n; ++r)
		{
			dyn_read_result& o = results[r];
			std::memset(&o, 0, sizeof(o));
			o.status = res.out[r].status == ST_REC_OVERFLOW ? (int32_t)ST_INTERNAL : res.out[r].status;
			o.seg_offset = res.seg_off[r];
			o.Z = res.out[r].Z;
			if (o.status == ST_INVALID_NT)
			{
				const uint64_t pos = io.seq_off[r] + res.bad_pos[r];
				if (io.seq_host) o.bad_char = io.seq_host[pos];
				else
				{
					A->rt.d2h(&o.bad_char, io.seq_dev + pos, 1);
					A->rt.sync();
				}
			}
			if (o.status == ST_OK && calc_probabilities) o.n_segments = res.seg_off[r + 1] - res.seg_off[r];
		}
		if (calc_probabilities)
		{
			const uint64_t half_k = (uint64_t)A->k / 2;
			parallel_ranges(io.n, seg_total, [&](uint32_t r_lo, uint32_t r_hi) {
				for (uint32_t r = r_lo; r < r_hi; ++r)
				{
					const dyn_read_result& o = results[r];
					if (o.status != ST_OK) continue;
					const uint64_t off = o.seg_offset, kc = o.n_segments;
					for (uint64_t i = 0; i < kc; ++i)
					{
						sequence_positions[off + i] = i + half_k;  // n - 1 + k/2 (NT:421)
						signal_positions[off + i] = sigpos[off + i];
					}
					std::memcpy(probabilities + off, prob_st + off, kc * sizeof(double));
				}
			});
		}
		tm.lap("fan_out");
		return 0;
	}
	catch (const std::exception& e)
	{
		A->last_error = e.what();
		return -1;
	}
}

int dyn_align_batch(dyn_aligner* A, const float* signal, const uint64_t* sig_off, const char* seq,
	const uint64_t* seq_off, uint32_t n_reads, int calc_probabilities, dyn_read_result* results,
	uint64_t* sequence_positions, uint64_t* signal_positions, double* probabilities)
{
	BatchIO io;
	io.sig_host = signal; io.seq_host = seq; io.sig_off = sig_off; io.seq_off = seq_off; io.n = n_reads;
	return align_common(A, io, calc_probabilities, results, sequence_positions, signal_positions, probabilities);
}

int dyn_align_batch_f64(dyn_aligner* A, const double* signal, const uint64_t* sig_off, const char* seq,
	const uint64_t* seq_off, uint32_t n_reads, int calc_probabilities, dyn_read_result* results,
	uint64_t* sequence_positions, uint64_t* signal_positions, double* probabilities)
{
	BatchIO io;
	io.sig_host64 = signal; io.seq_host = seq; io.sig_off = sig_off; io.seq_off = seq_off; io.n = n_reads;
	return align_common(A, io, calc_probabilities, results, sequence_positions, signal_positions, probabilities);
}

int dyn_align_batch_device(dyn_aligner* A, const float* d_signal, const uint64_t* sig_off, const char* d_seq,
	const uint64_t* seq_off, uint32_t n_reads, int calc_probabilities, dyn_read_result* results,
	uint64_t* sequence_positions, uint64_t* signal_positions, double* probabilities)
{
	BatchIO io;
	io.sig_dev = d_signal; io.seq_dev = d_seq; io.sig_off = sig_off; io.seq_off = seq_off; io.n = n_reads;
	return align_common(A, io, calc_probabilities, results, sequence_positions, signal_positions, probabilities);
}

// ---- asynchronous batches ---------------------------------------------------------------------------------------
static dyn_aligner* lane_of(dyn_aligner* A, int64_t ticket)
{
	const int li = (int)(ticket % dyn_aligner::LANES);
	if (!A->lane[li])
	{
		char err[512];
		int kind = 0;
		dyn_aligner* l = dyn_create(A->model_path_.c_str(), A->pore_.c_str(), A->mode_.c_str(), 1, A->band, A->device_, err, sizeof err, &kind);
		if (!l) throw std::runtime_error(std::string("dyn_align_submit: cannot create lane: ") + err);
		dyn_set_model(l, A->mean.data(), A->stdev.data());
		for (const auto& kv : A->options) dyn_set_option(l, kv.first.c_str(), kv.second);
		l->root = A;
		A->lane[li] = l;
	}
	return A->lane[li];
}

static int64_t submit_impl(dyn_aligner* A, bool on_device, const float* signal, const uint64_t* sig_off, const char* seq,
	const uint64_t* seq_off, uint32_t n_reads, int calc_probabilities, dyn_read_result* results, uint64_t* sequence_positions,
	uint64_t* signal_positions, double* probabilities)
{
	try
	{
		std::lock_guard<std::mutex> g(A->jobs_mu);
		const int64_t ticket = (int64_t)A->jobs.size();
		dyn_aligner* l = lane_of(A, ticket);
		dyn_aligner::Job* j = new dyn_aligner::Job();
		A->jobs.push_back(j);
		// the lane's own mutex serialises the calls that share it; the caller's buffers must stay valid until dyn_align_wait
		j->th = std::thread([=]() {
			j->rc = on_device
				? dyn_align_batch_device(l, signal, sig_off, seq, seq_off, n_reads, calc_probabilities, results, sequence_positions,
					  signal_positions, probabilities)
				: dyn_align_batch(l, signal, sig_off, seq, seq_off, n_reads, calc_probabilities, results, sequence_positions,
					  signal_positions, probabilities);
		});
		return ticket;
	}
	catch (const std::exception& e)
	{
		A->last_error = e.what();
		return -1;
	}
}

int64_t dyn_align_submit(dyn_aligner* A, const float* signal, const uint64_t* sig_off, const char* seq, const uint64_t* seq_off,
	uint32_t n_reads, int calc_probabilities, dyn_read_result* results, uint64_t* sequence_positions,
	uint64_t* signal_positions, double* probabilities)
{
	return submit_impl(A, false, signal, sig_off, seq, seq_off, n_reads, calc_probabilities, results, sequence_positions,
		signal_positions, probabilities);
}

int64_t dyn_align_submit_device(dyn_aligner* A, const float* d_signal, const uint64_t* sig_off, const char* d_seq,
	const uint64_t* seq_off, uint32_t n_reads, int calc_probabilities, dyn_read_result* results, uint64_t* sequence_positions,
	uint64_t* signal_positions, double* probabilities)
{
	return submit_impl(A, true, d_signal, sig_off, d_seq, seq_off, n_reads, calc_probabilities, results, sequence_positions,
		signal_positions, probabilities);
}

int dyn_align_wait(dyn_aligner* A, int64_t ticket)
{
	dyn_aligner::Job* j = nullptr;
	{
		std::lock_guard<std::mutex> g(A->jobs_mu);
		if (ticket < 0 || ticket >= (int64_t)A->jobs.size()) return -1;
		j = A->jobs[(size_t)ticket];
	}
	if (!j->joined)
	{
		if (j->th.joinable()) j->th.join();
		j->joined = true;
	}
	dyn_aligner* l = A->lane[ticket % dyn_aligner::LANES];
	if (j->rc != 0) A->last_error = l ? l->last_error : "dyn_align_wait: job failed";
	else if (l)
	{
		// what dyn_last_timing / dyn_last_ribbon / dyn_last_fallbacks report after a wait: the lane's last batch
		std::lock_guard<std::mutex> g(l->mu);
		A->timing[0] = l->timing[0]; A->timing[1] = l->timing[1]; A->timing[2] = l->timing[2];
		A->n_ribbon = l->n_ribbon; A->n_rib_fault = l->n_rib_fault; A->n_fallback = l->n_fallback; A->n_retry_lin = l->n_retry_lin;
		A->last_variant = l->last_variant;
		for (int i = 0; i < 16; ++i)
		{
			A->rib_reason[i] += l->rib_reason[i];
			l->rib_reason[i] = 0;
		}
		A->rib_recs_used = l->rib_recs_used;
		A->rib_last_two_level = l->rib_last_two_level;
	}
	return j->rc;
}

int dyn_train_batch(dyn_aligner* A, const float* signal, const uint64_t* sig_off, const char* seq,
	const uint64_t* seq_off, uint32_t n_reads, dyn_train_result* results, double* pooled_w, double* pooled_x,
	double* pooled_xx, double* pooled_xi, double* per_read_mean, double* per_read_stdev)
{
	std::lock_guard<std::mutex> g(A->mu);
	try
	{
		if (A->ntk) throw std::runtime_error(NTK_PARTIAL);
		BatchIO io;
		io.sig_host = signal; io.seq_host = seq; io.sig_off = sig_off; io.seq_off = seq_off; io.n = n_reads;
		BatchResult res;
		const uint64_t K = A->K;
		std::vector<double> pooled(3 * K, 0.0);
		uint64_t pc_total = 0;
		for (uint32_t r = 0; r < n_reads; ++r)
		{
			const uint64_t S = sig_off[r + 1] - sig_off[r], L = seq_off[r + 1] - seq_off[r];
			if (S >= 1 && L >= (uint64_t)A->k && S >= 2 * (L - A->k + 1)) pc_total += L - A->k + 2;
		}
		std::vector<double> cols;
		const bool want_cols = per_read_mean && per_read_stdev;
		if (want_cols) cols.assign(3 * pc_total, 0.0);
		// first pass with the usual sparse-record budget (the scratch of a resident warp then stays small enough for every
		// warp to be resident); a read whose records overflow has accumulated nothing (the statistics pass runs after the
		// overflow check) and is re-run alone with a full-size buffer, its statistics added to the batch's
		run_batch(*A, io, 2, res, nullptr, nullptr, pooled.data(), want_cols ? cols.data() : nullptr);
		std::vector<uint32_t> retry;
		for (uint32_t r = 0; r < n_reads; ++r)
			if (res.out[r].status == ST_REC_OVERFLOW) retry.push_back(r);
		if (!retry.empty())
		{
			const double saved = A->recs_per_row;
			A->recs_per_row = 1e9;  // clamped to the band width inside run_batch
			try
			{
				for (uint32_t r : retry)
				{
					BatchIO sh;
					uint64_t so0[2] = {0, sig_off[r + 1] - sig_off[r]};
					uint64_t qo0[2] = {0, seq_off[r + 1] - seq_off[r]};
					sh.sig_host = signal + sig_off[r];
					sh.seq_host = seq + seq_off[r];
					sh.sig_off = so0;
					sh.seq_off = qo0;
					sh.n = 1;
					BatchResult r1;
					std::vector<double> pooled1(3 * K, 0.0), cols1;
					const uint64_t ncol = res.desc[r].N;
					if (want_cols) cols1.assign(3 * ncol, 0.0);
					run_batch(*A, sh, 2, r1, nullptr, nullptr, pooled1.data(), want_cols ? cols1.data() : nullptr);
					res.out[r] = r1.out[0];
					for (uint64_t q = 0; q < 3 * K; ++q) pooled[q] += pooled1[q];
					if (want_cols)
						for (int a = 0; a < 3; ++a)
							std::copy(cols1.begin() + a * ncol, cols1.begin() + (a + 1) * ncol,
								cols.begin() + a * pc_total + res.desc[r].pc_off);
				}
			}
			catch (...)
			{
				A->recs_per_row = saved;
				throw;
			}
			A->recs_per_row = saved;
		}
		for (uint64_t q = 0; q < K; ++q)
		{
			if (pooled_w) pooled_w[q] += pooled[q];
			if (pooled_x) pooled_x[q] += pooled[K + q];
			if (pooled_xx) pooled_xx[q] += pooled[2 * K + q];
		}
		// kmer ids of every column are needed for the per-read M-step: recompute on the host (cheap, L ints)
		for (uint32_t r = 0; r < n_reads; ++r)
		{
			dyn_train_result& o = results[r];
			std::memset(&o, 0, sizeof(o));
			o.status = res.out[r].status;
			o.Z = res.out[r].Z;
			if (o.status == ST_INVALID_NT) o.bad_char = seq[seq_off[r] + res.bad_pos[r]];
			if (o.status != ST_OK) continue;
			// NT:703-722: normalised re-estimates of m1 and e2; e1 = exp(log 1)
			const double xm = res.out[r].xi_m, xe = res.out[r].xi_e;
			const double norm = xm + xe;
			o.m1 = norm > 0 ? xm / norm : 0.0;
			o.e2 = norm > 0 ? xe / norm : 0.0;
			o.e1 = std::exp(A->trans[1]);
			if (pooled_xi)
			{
				pooled_xi[0] += xm;
				pooled_xi[1] += xe;
			}
			if (want_cols)
			{
				// NT:519-535 per-read M-step
				double* pm = per_read_mean + (uint64_t)r * K;
				double* ps = per_read_stdev + (uint64_t)r * K;
				std::vector<double> w(K, 0.0), sx(K, 0.0), sxx(K, 0.0);
				const ReadDesc& d = res.desc[r];
				const char* s = seq + seq_off[r];
				for (uint32_t n = 1; n < d.N; ++n)
				{
					uint64_t id = 0;
					for (int i = 0; i < A->k; ++i) id = id * 4 + (uint64_t)host_digit((unsigned char)s[n - 1 + i]);
					w[id] += cols[d.pc_off + n];
					sx[id] += cols[pc_total + d.pc_off + n];
					sxx[id] += cols[2 * pc_total + d.pc_off + n];
				}
				for (uint64_t q = 0; q < K; ++q)
				{
					if (w[q] > 0.0)
					{
						const double mu = sx[q] / w[q];
						double var = sxx[q] / w[q] - mu * mu;
						if (var < 1e-12) var = 1e-12;
						pm[q] = mu;
						ps[q] = std::sqrt(var);
					}
					else
					{
						pm[q] = A->mean[q];
						ps[q] = A->stdev[q];
					}
				}
			}
		}
		return 0;
	}
	catch (const std::exception& e)
	{
		A->last_error = e.what();
		return -1;
	}
}

int dyn_train_accumulate(dyn_aligner* A, const float* signal, const uint64_t* sig_off, const char* seq, const uint64_t* seq_off,
	uint32_t n_reads, int inputs_on_device, double* d_stats, int32_t* status)
{
	std::lock_guard<std::mutex> g(A->mu);
	try
	{
		if (A->ntk) throw std::runtime_error(NTK_PARTIAL);
		BatchIO io;
		if (inputs_on_device) { io.sig_dev = signal; io.seq_dev = seq; }
		else { io.sig_host = signal; io.seq_host = seq; }
		io.sig_off = sig_off; io.seq_off = seq_off; io.n = n_reads;
		struct Guard
		{
			dyn_aligner* A;
			double recs;
			~Guard() { A->ext_stats = nullptr; A->recs_per_row = recs; }
		} guard{A, A->recs_per_row};
		A->ext_stats = d_stats;
		BatchResult res;
		run_batch(*A, io, 2, res, nullptr, nullptr, nullptr, nullptr);
		// a read whose sparse records overflowed (full-band kernels only) has accumulated nothing: re-run it alone
		std::vector<uint32_t> retry;
		for (uint32_t r = 0; r < n_reads; ++r)
			if (res.out[r].status == ST_REC_OVERFLOW) retry.push_back(r);
		if (!retry.empty())
		{
			const uint64_t f0 = A->n_fallback, f1 = A->n_retry_lin, f2 = A->n_ribbon, f3 = A->n_rib_fault;
			const double t1 = A->timing[1];
			A->recs_per_row = 1e9;
			for (uint32_t r : retry)
			{
				BatchIO sh;
				uint64_t so0[2] = {0, sig_off[r + 1] - sig_off[r]};
				uint64_t qo0[2] = {0, seq_off[r + 1] - seq_off[r]};
				if (inputs_on_device) { sh.sig_dev = signal + sig_off[r]; sh.seq_dev = seq + seq_off[r]; }
				else { sh.sig_host = signal + sig_off[r]; sh.seq_host = seq + seq_off[r]; }
				sh.sig_off = so0; sh.seq_off = qo0; sh.n = 1;
				BatchResult r1;
				run_batch(*A, sh, 2, r1, nullptr, nullptr, nullptr, nullptr);
				res.out[r] = r1.out[0];
			}
			A->n_fallback = f0; A->n_retry_lin = f1; A->n_ribbon = f2; A->n_rib_fault = f3;
			A->timing[1] = t1;
		}
		if (status)
			for (uint32_t r = 0; r < n_reads; ++r) status[r] = res.out[r].status == ST_REC_OVERFLOW ? (int32_t)ST_INTERNAL : res.out[r].status;
		return 0;
	}
	catch (const std::exception& e)
	{
		A->last_error = e.what();
		return -1;
	}
}

int dyn_train_mstep_device(dyn_aligner* A, const double* d_stats, double* transitions3)
{
	std::lock_guard<std::mutex> g(A->mu);
	try
	{
		Rt& rt = A->rt;
		rt.bind();
		const uint64_t K = A->K;
		// persistent device buffers and pinned staging: no allocation (= device synchronisation) per iteration
		double* d_mean = (double*)A->d_sx.get(rt, K * 8);
		double* d_sd = (double*)A->d_sxx.get(rt, K * 8);
		A->h_arena.get(rt, 2 * K * 8 + 4096);
		double* h_mean = (double*)A->h_arena.p;
		double* h_sd = h_mean + K;
		double* h_tail = h_sd + K;
		std::memcpy(h_mean, A->mean.data(), K * 8);
		std::memcpy(h_sd, A->stdev.data(), K * 8);
		rt.h2d(d_mean, h_mean, K * 8);
		rt.h2d(d_sd, h_sd, K * 8);
		MStepArgs ma;
		ma.stats = d_stats; ma.K = K; ma.mean = d_mean; ma.stdev = d_sd;
#ifndef DYN_HOST_EMU
		k_mstep<<<(unsigned)std::min<uint64_t>((K + 255) / 256, 1024), 256, 0, rt.stream>>>(ma);
		CK_CUDA(cudaGetLastError());
#else
		for (uint64_t q = 0; q < K; ++q) mstep_kmer(ma, q);
#endif
		double tail[4] = {0, 0, 0, 0};
		rt.d2h(h_mean, d_mean, K * 8);
		rt.d2h(h_sd, d_sd, K * 8);
		rt.d2h(h_tail, d_stats + 3 * K, 32);
		rt.sync();
		std::memcpy(A->mean.data(), h_mean, K * 8);
		std::memcpy(A->stdev.data(), h_sd, K * 8);
		std::memcpy(tail, h_tail, 32);
		A->table_dirty = true;
		for (dyn_aligner* l : A->lane)
			if (l) dyn_set_model(l, A->mean.data(), A->stdev.data());
		if (transitions3)
		{
			// NT:703-722: normalised re-estimates of m1 and e2; e1 = exp(log 1)
			const double norm = tail[0] + tail[1];
			transitions3[0] = norm > 0 ? tail[0] / norm : 0.0;
			transitions3[1] = std::exp(A->trans[1]);
			transitions3[2] = norm > 0 ? tail[1] / norm : 0.0;
		}
		return 0;
	}
	catch (const std::exception& e)
	{
		A->last_error = e.what();
		return -1;
	}
}

int dyn_preprocess_batch(dyn_aligner* A, const float* raw, const uint64_t* sig_off, uint32_t n_reads, const double* shift,
	const double* scale, int window, double n_sigmas, float* out)
{
	std::lock_guard<std::mutex> g(A->mu);
	try
	{
		if (window < 1 || window > 15) throw std::runtime_error("dyn_preprocess_batch: window must be in [1, 15]");
		Rt& rt = A->rt;
		rt.bind();
		const uint64_t total = n_reads ? sig_off[n_reads] : 0;
		if (!total) return 0;
		DevBuf b_raw, b_off, b_sh, b_sc, b_out;
		PreprocArgs pa;
		float* d_raw = (float*)b_raw.get(rt, total * 4);
		uint64_t* d_off = (uint64_t*)b_off.get(rt, ((size_t)n_reads + 1) * 8);
		double* d_sh = (double*)b_sh.get(rt, (size_t)n_reads * 8);
		double* d_sc = (double*)b_sc.get(rt, (size_t)n_reads * 8);
		float* d_out = (float*)b_out.get(rt, total * 4);
		rt.h2d(d_raw, raw, total * 4);
		rt.h2d(d_off, sig_off, ((size_t)n_reads + 1) * 8);
		rt.h2d(d_sh, shift, (size_t)n_reads * 8);
		rt.h2d(d_sc, scale, (size_t)n_reads * 8);
		pa.raw = d_raw; pa.sig_off = d_off; pa.n_reads = n_reads; pa.shift = d_sh; pa.scale = d_sc;
		pa.window = window; pa.n_sigmas = n_sigmas; pa.out = d_out;
#ifndef DYN_HOST_EMU
		const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)rt.sms * 32);
		k_preprocess<<<grid, 256, 0, rt.stream>>>(pa, total);
		CK_CUDA(cudaGetLastError());
#else
		for (uint64_t i = 0; i < total; ++i) preprocess_sample(pa, i);
#endif
		rt.d2h(out, d_out, total * 4);
		rt.sync();
		for (DevBuf* b : {&b_raw, &b_off, &b_sh, &b_sc, &b_out}) b->release(rt);
		return 0;
	}
	catch (const std::exception& e)
	{
		A->last_error = e.what();
		return -1;
	}
}

// utils.segmentation_to_string (utils.py:193-232): one CSV line per segment,
//   readid,signalid,start,end,basepos,base,motif,state,probability(.6f),polish
// start = signal_position + sigOffset, end = next segment's start (lastIndex for the last one); for RNA reads the motif
// is reversed and basepos counted from the other end; polish "NA" when empty.  Host-side formatting.
int64_t dyn_format_segments(const char* readid, const char* signalid, int64_t sig_offset, int64_t last_index, const char* read,
	int kmer_size, int rna, uint64_t n_segments, const uint64_t* sequence_positions, const uint64_t* signal_positions,
	const double* probabilities, const char* states, const char* const* polishes, char* out, uint64_t out_cap)
{
	const std::string rd(read);
	std::string buf;
	buf.reserve(n_segments * 96);
	char num[64];
	for (uint64_t i = 0; i < n_segments; ++i)
	{
		int64_t basepos = (int64_t)sequence_positions[i];
		const int64_t start = (int64_t)signal_positions[i] + sig_offset;
		const int64_t end = (i + 1 < n_segments) ? (int64_t)signal_positions[i + 1] + sig_offset : last_index;
		const int64_t m0 = std::max<int64_t>(0, basepos - kmer_size / 2);
		const int64_t m1 = std::min<int64_t>((int64_t)rd.size(), basepos + kmer_size / 2 + 1);
		std::string motif = m1 > m0 ? rd.substr((size_t)m0, (size_t)(m1 - m0)) : std::string();
		const char base = rd[(size_t)basepos];
		if (rna)
		{
			std::reverse(motif.begin(), motif.end());
			basepos = (int64_t)rd.size() - basepos - 1;
		}
		buf += readid; buf += ','; buf += signalid; buf += ',';
		buf += std::to_string(start); buf += ','; buf += std::to_string(end); buf += ','; buf += std::to_string(basepos); buf += ',';
		buf += base; buf += ','; buf += motif; buf += ','; buf += states[i]; buf += ',';
		snprintf(num, sizeof num, "%.6f", probabilities[i]);
		buf += num; buf += ',';
		const char* pol = polishes ? polishes[i] : nullptr;
		buf += (pol && pol[0]) ? pol : "NA";
		buf += '\n';
	}
	if (out && buf.size() <= out_cap) std::memcpy(out, buf.data(), buf.size());
	return (int64_t)buf.size();
}

void dyn_ntk_transitions(const dyn_aligner* A, double* out18) { std::memcpy(out18, A->ntk_trans, sizeof(A->ntk_trans)); }

#ifndef DYN_HOST_EMU
namespace
{

// device state of one resquiggle-mode read between the pre-pass and the sparse stages
struct NtkRun
{
	DevBuf b_model, b_sig, b_kmers, b_lat, b_rows, b_z, b_tn, b_tk, b_cnt, b_keys, b_sparse, b_seg, b_wide;
	uint32_t T = 0, N = 0, K = 0, hp = 1, wn = 0, wk = 0;
	uint64_t Kc = 0, total = 0;
	double z[4] = {0, 0, 0, 0};
	std::vector<uint64_t> rowptr;  // [T+1]
	dyn::ntk::Consts consts;
	void release(Rt& rt)
	{
		for (DevBuf* b : {&b_model, &b_sig, &b_kmers, &b_lat, &b_rows, &b_z, &b_tn, &b_tk, &b_cnt, &b_keys, &b_sparse, &b_seg, &b_wide}) b->release(rt);
	}
};

// validation + kmer encoding + TN / TK pre-passes + row masks + keys (NTK_aligner_api.cpp:197-441), all on the device.
// Returns a dyn_status; throws on CUDA errors.
int ntk_prepass_device(dyn_aligner* A, Rt& rt, const float* signal, uint64_t S, const char* seq, uint64_t L, NtkRun& R)
{
	using namespace dyn::ntk;
	rt.bind();
	// Aligner::validateInput (aligner.cpp:145-164) and sequenceToKmers (:166-205)
	if (S < 1) return DYN_SIGNAL_EMPTY;
	if (L < (uint64_t)A->k) return DYN_SEQ_SHORT;
	const uint64_t Kc = L - A->k + 1;
	if (S < 2 * Kc) return DYN_SIGNAL_SHORT;
	std::vector<int32_t> kmers(Kc);
	for (uint64_t c = 0; c < Kc; ++c)
	{
		uint64_t id = 0;
		for (int i = 0; i < A->k; ++i)
		{
			const int d = host_digit((unsigned char)seq[c + i]);
			if (d < 0 || d > 3) return DYN_INVALID_NT;
			id = id * 4 + (uint64_t)d;
		}
		kmers[c] = (int32_t)id;
	}
	const uint32_t T = (uint32_t)(S + 1), N = (uint32_t)(Kc + 1), K = (uint32_t)A->K;
	const uint32_t wn = (N + 31) / 32, wk = (K + 31) / 32;
	uint32_t hp = 1;
	for (int i = 1; i < A->k; ++i) hp *= 4;
	R.T = T; R.N = N; R.K = K; R.hp = hp; R.wn = wn; R.wk = wk; R.Kc = Kc;
	const size_t C = std::max<size_t>(N, K);
	std::vector<KmerModel> km(K);
	for (uint32_t q = 0; q < K; ++q)
	{
		km[q].mean = A->mean[q];
		km[q].stdev = A->stdev[q];
		km[q].log_stdev = std::log(A->stdev[q]);
	}
	std::vector<double> sig(S);
	for (uint64_t i = 0; i < S; ++i) sig[i] = (double)signal[i];
	KmerModel* d_model = (KmerModel*)R.b_model.get(rt, K * sizeof(KmerModel));
	double* d_sig = (double*)R.b_sig.get(rt, S * 8);
	int32_t* d_kmers = (int32_t*)R.b_kmers.get(rt, Kc * 4);
	// forward lattice (M, E) + log posteriors: 3 x T x C doubles; two backward rows; per-row maxima and scales
	double* d_lat = (double*)R.b_lat.get(rt, ((size_t)3 * T * C + 4 * C) * 8);
	unsigned char* d_rows = (unsigned char*)R.b_rows.get(rt, (size_t)T * 24 + 64);
	double* d_z = (double*)R.b_z.get(rt, 4 * 8);
	uint32_t* d_tn = (uint32_t*)R.b_tn.get(rt, (size_t)T * wn * 4);
	uint32_t* d_tk = (uint32_t*)R.b_tk.get(rt, (size_t)T * wk * 4);
	uint64_t* d_cnt = (uint64_t*)R.b_cnt.get(rt, ((size_t)T + 1) * 8);
	rt.h2d(d_model, km.data(), K * sizeof(KmerModel));
	rt.h2d(d_sig, sig.data(), S * 8);
	rt.h2d(d_kmers, kmers.data(), Kc * 4);
	Consts cc;
	cc.model = d_model;
	cc.half_log_2pi = 0.5 * std::log(2.0 * M_PI);
	cc.m = A->ntk_trans[16];
	cc.e = A->ntk_trans[17];
	// SPARSETHRESHOLD (NTK_aligner_api.cpp:17): the literal the reference ships, commented there as log(0.95)
	// but numerically log10(0.95), i.e. a posterior mass of 0.97797
	const double threshold = -0.02227639471;
	const double EPS = 1e-8;
	double* z = R.z;
	// one dense pre-pass (ntk_prepass.cuh): forward lattice, backward rows + log posteriors, Zf / Zb
	auto prepass = [&](int tk, uint32_t cols, double log_m, double log_e, double* zout) {
		PreArgs pa;
		pa.signal = d_sig; pa.kmers = d_kmers; pa.model = d_model; pa.half_log_2pi = cc.half_log_2pi;
		pa.m = std::exp(log_m); pa.e = std::exp(log_e);
		pa.T = T; pa.C = cols; pa.hp = hp; pa.tk = tk;
		pa.fM = d_lat; pa.fE = d_lat + (size_t)T * cols; pa.LP = d_lat + (size_t)2 * T * cols; pa.brow = d_lat + (size_t)3 * T * cols;
		pa.fmax = (unsigned long long*)d_rows; pa.bmax = pa.fmax + T;
		pa.fexp = (int*)(pa.bmax + T); pa.bexp = pa.fexp + T;
		pa.z = zout;
		rt.zero(d_rows, (size_t)T * 24);
		const uint32_t work = tk ? hp : cols;  // independent work items per row
		if (work < 4096)
		{
			// small lattice: one CTA per direction walks the rows (reads are batched across streams)
			k_pre_forward_cta<<<1, 1024, 0, rt.stream>>>(pa);
			k_pre_backward_cta<<<1, 1024, 0, rt.stream>>>(pa);
		}
		else
		{
			// large lattice (9-mers: 65 536 work items, 262 144 cells per row): one launch per row over the whole GPU
			const unsigned grid = (unsigned)std::min<uint32_t>((work + 255) / 256, (uint32_t)rt.sms * 8);
			k_pre_init<<<grid, 256, 0, rt.stream>>>(pa, 1);
			for (uint32_t t = 1; t < T; ++t) k_pre_forward_row<<<grid, 256, 0, rt.stream>>>(pa, t);
			k_pre_init<<<grid, 256, 0, rt.stream>>>(pa, 0);
			for (uint32_t t = T - 1; t-- > 0;) k_pre_backward_row<<<grid, 256, 0, rt.stream>>>(pa, t);
		}
		k_pre_z<<<2, 1024, 0, rt.stream>>>(pa);
		CK_CUDA(cudaGetLastError());
		return pa.LP;
	};
	// ---- TN (preProcTN, NTK:315-354)
	{
		const double* lp = prepass(0, N, A->ntk_trans[14], A->ntk_trans[15], d_z);
		rt.d2h(z, d_z, 16);
		rt.sync();
		if (std::abs(z[0] - z[1]) / (double)((size_t)T * N) > EPS || std::isinf(z[0]) || std::isinf(z[1]) || std::isnan(z[0]) || std::isnan(z[1])) return DYN_NTK_TN_FAILED;
		// LP = logP - Zf (NTK:336): folded into the threshold of the selection
		k_row_mask<<<T, 256, 0, rt.stream>>>(lp, N, wn, d_tn, threshold + z[0]);
		CK_CUDA(cudaGetLastError());
		rt.sync();
	}
	// ---- TK (preProcTK, NTK:356-400)
	{
		const double* lp = prepass(1, K, A->ntk_trans[16], A->ntk_trans[17], d_z + 2);
		rt.d2h(z + 2, d_z + 2, 16);
		rt.sync();
		if (std::abs(z[2] - z[3]) / (double)((size_t)T * K) > EPS || std::isinf(z[2]) || std::isinf(z[3]) || std::isnan(z[2]) || std::isnan(z[3])) return DYN_NTK_TK_FAILED;
		if (K >= 4 * RM_CHUNK)
		{
			// wide rows (9-mers): row maximum and candidate compaction over (row, chunk) CTAs, one CTA per row for the sort +
			// sequential log-sum-exp; k_row_mask only for the rows that path leaves undecided
			const size_t o_max = 0, o_cnt = align_up(o_max + (size_t)T * 8, 256), o_todo = align_up(o_cnt + (size_t)T * 4, 256),
						 o_val = align_up(o_todo + (size_t)T * 4, 256), o_idx = align_up(o_val + (size_t)T * RM_WIDE_CAP * 8, 256),
						 o_end = o_idx + (size_t)T * RM_WIDE_CAP * 4;
			unsigned char* d_w = (unsigned char*)R.b_wide.get(rt, o_end);
			rt.zero(d_w, o_val);
			rt.zero(d_tk, (size_t)T * wk * 4);
			WideMaskArgs wa;
			wa.LP = lp; wa.C = K; wa.T = T; wa.words = wk; wa.mask = d_tk; wa.threshold = threshold + z[3];  // Zb (NTK:382)
			wa.rowmax = (unsigned long long*)(d_w + o_max); wa.cnt = (uint32_t*)(d_w + o_cnt); wa.todo = (uint32_t*)(d_w + o_todo);
			wa.cap = RM_WIDE_CAP; wa.cval = (double*)(d_w + o_val); wa.cidx = (uint32_t*)(d_w + o_idx);
			const dim3 g2((K + RM_CHUNK - 1) / RM_CHUNK, T);
			k_wide_rowmax<<<g2, 256, 0, rt.stream>>>(wa);
			k_wide_candidates<<<g2, 256, 0, rt.stream>>>(wa);
			k_wide_select<<<T, 1024, 0, rt.stream>>>(wa);
			k_row_mask<<<T, 256, 0, rt.stream>>>(lp, K, wk, d_tk, wa.threshold, wa.todo);
		}
		else k_row_mask<<<T, 256, 0, rt.stream>>>(lp, K, wk, d_tk, threshold + z[3]);  // Zb (NTK:382)
		CK_CUDA(cudaGetLastError());
	}
	// ---- keys (preProcTNK, NTK:402-441)
	KeyArgs ka;
	ka.tn = d_tn; ka.tk = d_tk; ka.kmers = d_kmers; ka.T = T; ka.N = N; ka.K = K; ka.wn = wn; ka.wk = wk;
	ka.count = d_cnt; ka.keys = nullptr;
	k_keys<false><<<(T + 127) / 128, 128, 0, rt.stream>>>(ka);
	CK_CUDA(cudaGetLastError());
	R.rowptr.assign((size_t)T + 1, 0);
	rt.d2h(R.rowptr.data(), d_cnt, (size_t)T * 8);
	rt.sync();
	uint64_t total = 0;
	for (uint32_t t = 0; t < T; ++t)
	{
		const uint64_t c = R.rowptr[t];
		R.rowptr[t] = total;
		total += c;
	}
	R.rowptr[T] = total;
	R.total = total;
	uint64_t* d_keys = (uint64_t*)R.b_keys.get(rt, std::max<uint64_t>(total, 1) * 8);
	rt.h2d(d_cnt, R.rowptr.data(), ((size_t)T + 1) * 8);
	ka.keys = d_keys;
	if (total) k_keys<true><<<(T + 127) / 128, 128, 0, rt.stream>>>(ka);
	CK_CUDA(cudaGetLastError());
	rt.sync();
	R.consts = cc;
	return 0;
}

} // namespace
#endif

int dyn_ntk_prepass(dyn_aligner* A, const float* signal, uint64_t S, const char* seq, uint64_t L, uint32_t* tn_mask,
	uint32_t* tk_mask, uint64_t* keys, uint64_t keys_cap, uint64_t* n_keys, double* z4)
{
#ifdef DYN_HOST_EMU
	(void)signal; (void)S; (void)seq; (void)L; (void)tn_mask; (void)tk_mask; (void)keys; (void)keys_cap; (void)n_keys; (void)z4;
	A->last_error = "dyn_ntk_prepass: not available in the emulator build";
	return -1;
#else
	std::lock_guard<std::mutex> g(A->mu);
	NtkRun R;
	try
	{
		const int st = ntk_prepass_device(A, A->rt, signal, S, seq, L, R);
		if (st == 0)
		{
			Rt& rt = A->rt;
			if (n_keys) *n_keys = R.total;
			if (keys && R.total && R.total <= keys_cap) rt.d2h(keys, R.b_keys.p, R.total * 8);
			if (tn_mask) rt.d2h(tn_mask, R.b_tn.p, (size_t)R.T * R.wn * 4);
			if (tk_mask) rt.d2h(tk_mask, R.b_tk.p, (size_t)R.T * R.wk * 4);
			rt.sync();
			if (z4) std::memcpy(z4, R.z, sizeof(R.z));
		}
		R.release(A->rt);
		return st;
	}
	catch (const std::exception& e)
	{
		R.release(A->rt);
		A->last_error = e.what();
		return -1;
	}
#endif
}

#ifndef DYN_HOST_EMU
namespace
{

// NTKAligner::align for one read on the stream of `rt`, with the (reusable) device buffers of R.  Returns a dyn_status;
// throws on CUDA errors.
int ntk_align_one(dyn_aligner* A, Rt& rt, NtkRun& R, const float* signal, uint64_t S, const char* seq, uint64_t L,
	int calc_probabilities, double* Z, uint64_t* n_segments, char* states, uint64_t* sequence_positions,
	uint64_t* signal_positions, double* probabilities, uint32_t* polish_kmers, uint64_t cap)
{
	using namespace dyn::ntk;
	if (n_segments) *n_segments = 0;
	static const bool trace = getenv("DYN_NTK_TRACE") != nullptr;
	const auto c0 = std::chrono::steady_clock::now();
	int st = ntk_prepass_device(A, rt, signal, S, seq, L, R);
	if (st != 0) return st;
	const auto c1 = std::chrono::steady_clock::now();
	const uint64_t nk = std::max<uint64_t>(R.total, 1);
	const uint32_t T = R.T, N = R.N;
	// F, B, LP, V [nk][5] doubles; Z[2]; status; segment arrays [T + N]
	const size_t segcap = (size_t)T + N + 8;
	const size_t sparse_bytes = nk * 5 * 8 * 4;
	double* d_sparse = (double*)R.b_sparse.get(rt, sparse_bytes);
	size_t o = 0;
	auto carve = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes, 16); return at; };
	const size_t o_z = carve(16), o_st = carve(4), o_ns = carve(4), o_state = carve(segcap), o_seq = carve(segcap * 8),
				 o_sig = carve(segcap * 8), o_pr = carve(segcap * 8), o_km = carve(segcap * 4), o_buf = carve(segcap * 8);
	unsigned char* d_seg = (unsigned char*)R.b_seg.get(rt, o);
	rt.zero(d_seg, o);
	SparseArgs sa;
	sa.signal = (const double*)R.b_sig.p; sa.kmers = (const int32_t*)R.b_kmers.p;
	sa.keys = (const uint64_t*)R.b_keys.p; sa.rowptr = (const uint64_t*)R.b_cnt.p; sa.nk = R.total;
	sa.T = T; sa.N = N; sa.K = R.K; sa.hp = R.hp; sa.k = (uint32_t)A->k;
	sa.c = R.consts;
	for (int i = 0; i < 14; ++i) sa.tr[i] = A->ntk_trans[i];
	sa.F = d_sparse; sa.B = d_sparse + nk * 5; sa.LP = d_sparse + nk * 10; sa.V = d_sparse + nk * 15;
	sa.out_z = (double*)(d_seg + o_z); sa.out_status = (int32_t*)(d_seg + o_st); sa.seg_n = (uint32_t*)(d_seg + o_ns);
	sa.seg_state = (char*)(d_seg + o_state); sa.seg_seqpos = (uint64_t*)(d_seg + o_seq); sa.seg_sigpos = (uint64_t*)(d_seg + o_sig);
	sa.seg_prob = (double*)(d_seg + o_pr); sa.seg_kmer = (uint32_t*)(d_seg + o_km); sa.prob_buf = (double*)(d_seg + o_buf);
	sa.calc_prob = calc_probabilities;
	k_ntk_sparse_fb<<<2, 32, 0, rt.stream>>>(sa);
	k_ntk_sparse<<<1, 32, 0, rt.stream>>>(sa);
	CK_CUDA(cudaGetLastError());
	std::vector<unsigned char> h(o);
	rt.d2h(h.data(), d_seg, o);
	rt.sync();
	if (trace)
	{
		const auto c2 = std::chrono::steady_clock::now();
		fprintf(stderr, "ntk read T=%u keys=%llu: prepass %.1f ms, sparse %.1f ms\n", T, (unsigned long long)R.total,
			std::chrono::duration<double, std::milli>(c1 - c0).count(), std::chrono::duration<double, std::milli>(c2 - c1).count());
	}
	const double* hz = (const double*)(h.data() + o_z);
	const int32_t kst = *(const int32_t*)(h.data() + o_st);
	if (kst == 1) return DYN_NTK_ALIGN_FAILED;  // NTK:913-918
	if (kst != 0) return DYN_INTERNAL;
	if (Z) *Z = hz[1];  // result.Z = Zb (NTK:920)
	const uint32_t ns = *(const uint32_t*)(h.data() + o_ns);
	if (calc_probabilities)
	{
		if (ns > cap) return DYN_INTERNAL;
		// the traceback emits segments from the end of the read: reverse (NTK:800)
		for (uint32_t i = 0; i < ns; ++i)
		{
			const uint32_t j = ns - 1 - i;
			states[i] = ((const char*)(h.data() + o_state))[j];
			sequence_positions[i] = ((const uint64_t*)(h.data() + o_seq))[j];
			signal_positions[i] = ((const uint64_t*)(h.data() + o_sig))[j];
			probabilities[i] = ((const double*)(h.data() + o_pr))[j];
			polish_kmers[i] = ((const uint32_t*)(h.data() + o_km))[j];
		}
		if (n_segments) *n_segments = ns;
	}
	return 0;
}

} // namespace
#endif

int dyn_ntk_align(dyn_aligner* A, const float* signal, uint64_t S, const char* seq, uint64_t L, int calc_probabilities,
	double* Z, uint64_t* n_segments, char* states, uint64_t* sequence_positions, uint64_t* signal_positions,
	double* probabilities, uint32_t* polish_kmers, uint64_t cap)
{
#ifdef DYN_HOST_EMU
	(void)signal; (void)S; (void)seq; (void)L; (void)calc_probabilities; (void)Z; (void)n_segments; (void)states;
	(void)sequence_positions; (void)signal_positions; (void)probabilities; (void)polish_kmers; (void)cap;
	A->last_error = "dyn_ntk_align: not available in the emulator build";
	return -1;
#else
	std::lock_guard<std::mutex> g(A->mu);
	NtkRun R;
	try
	{
		const int st = ntk_align_one(A, A->rt, R, signal, S, seq, L, calc_probabilities, Z, n_segments, states,
			sequence_positions, signal_positions, probabilities, polish_kmers, cap);
		R.release(A->rt);
		return st;
	}
	catch (const std::exception& e)
	{
		R.release(A->rt);
		A->last_error = e.what();
		return -1;
	}
#endif
}

int dyn_ntk_align_batch(dyn_aligner* A, const float* signal, const uint64_t* sig_off, const char* seq, const uint64_t* seq_off,
	uint32_t n_reads, int calc_probabilities, int32_t* status, double* Z, uint64_t* n_segments, const uint64_t* out_off,
	char* states, uint64_t* sequence_positions, uint64_t* signal_positions, double* probabilities, uint32_t* polish_kmers,
	int concurrency)
{
#ifdef DYN_HOST_EMU
	(void)signal; (void)sig_off; (void)seq; (void)seq_off; (void)n_reads; (void)calc_probabilities; (void)status; (void)Z;
	(void)n_segments; (void)out_off; (void)states; (void)sequence_positions; (void)signal_positions; (void)probabilities;
	(void)polish_kmers; (void)concurrency;
	A->last_error = "dyn_ntk_align_batch: not available in the emulator build";
	return -1;
#else
	std::lock_guard<std::mutex> g(A->mu);
	try
	{
		// Reads are independent: a pool of host threads, each with its own CUDA stream and its own (reused) device
		// buffers, pulls reads from a queue, so that the small per-read grids of many reads overlap on the device.
		int workers = (int)std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)(concurrency > 0 ? concurrency : 64), n_reads));
		{
			// every read in flight holds its dense pre-pass lattice (3 x T x max(N, K) doubles): 9-mers need 6.3 GB per 1000
			// samples, so the pool is as wide as the free HBM allows
			uint64_t maxS = 0, maxL = 0;
			for (uint32_t r = 0; r < n_reads; ++r)
			{
				maxS = std::max<uint64_t>(maxS, sig_off[r + 1] - sig_off[r]);
				maxL = std::max<uint64_t>(maxL, seq_off[r + 1] - seq_off[r]);
			}
			const double per_read = 24.0 * (double)(maxS + 1) * (double)std::max<uint64_t>(A->K, maxL) * 1.15 + 64e6;
			A->rt.bind();
			const double fit = 0.6 * (double)A->rt.free_bytes() / per_read;
			workers = (int)std::max(1.0, std::min((double)workers, fit));
		}
		{
			// the workers' buffers are stream-ordered allocations: keep what they free in the device's pool between calls (by
			// default the pool gives everything back at the next synchronisation, and every call would map tens of GB again —
			// that, not the kernels, was 90 % of a 9-mer batch); dyn_destroy trims the pool
			cudaMemPool_t pool_h;
			if (cudaDeviceGetDefaultMemPool(&pool_h, A->rt.device) == cudaSuccess)
			{
				uint64_t keep = ~0ull;
				cudaMemPoolSetAttribute(pool_h, cudaMemPoolAttrReleaseThreshold, &keep);
				A->ntk_pool_used = true;
			}
		}
		std::atomic<uint32_t> next(0);
		std::vector<std::string> errors(workers);
		auto work = [&](int wi) {
			Rt rt = A->rt;  // same device; own stream, own events unused
			NtkRun R;
			try
			{
				CK_CUDA(cudaSetDevice(rt.device));
				CK_CUDA(cudaStreamCreateWithFlags(&rt.stream, cudaStreamNonBlocking));
				rt.async_alloc = true;
				while (true)
				{
					const uint32_t r = next.fetch_add(1);
					if (r >= n_reads) break;
					const uint64_t S = sig_off[r + 1] - sig_off[r], L = seq_off[r + 1] - seq_off[r];
					const uint64_t o = out_off[r], cap = out_off[r + 1] - out_off[r];
					status[r] = ntk_align_one(A, rt, R, signal + sig_off[r], S, seq + seq_off[r], L, calc_probabilities, Z + r,
						n_segments + r, states + o, sequence_positions + o, signal_positions + o, probabilities + o,
						polish_kmers + o, cap);
				}
			}
			catch (const std::exception& e)
			{
				errors[wi] = e.what();
			}
			R.release(rt);
			if (rt.stream)
			{
				cudaStreamSynchronize(rt.stream);
				cudaStreamDestroy(rt.stream);
			}
		};
		std::vector<std::thread> pool;
		for (int wi = 0; wi < workers; ++wi) pool.emplace_back(work, wi);
		for (auto& th : pool) th.join();
		for (const std::string& e : errors)
			if (!e.empty()) throw std::runtime_error(e);
		return 0;
	}
	catch (const std::exception& e)
	{
		A->last_error = e.what();
		return -1;
	}
#endif
}

const char* dyn_status_message(int status)
{
	switch (status)
	{
	case DYN_OK: return "";
	case DYN_SIGNAL_EMPTY: return "Signal is empty";
	case DYN_SEQ_SHORT: return "Sequence shorter than model kmer size";
	case DYN_SIGNAL_SHORT: return "Signal too short compared to sequence";
	case DYN_INVALID_NT: return "Invalid nucleotide: ";
	case DYN_ALIGN_FAILED: return "Alignment failed: alignment scores do not match";
	case DYN_TRAIN_FAILED: return "Training failed: alignment scores do not match";
	case DYN_NTK_TN_FAILED: return "NTK preprocessing TN failed: alignment scores do not match";
	case DYN_NTK_TK_FAILED: return "NTK preprocessing TK failed: alignment scores do not match";
	case DYN_NTK_ALIGN_FAILED: return "NTK alignment failed: alignment scores do not match";
	case DYN_BAND_UNSUPPORTED: return "dynamont_b200: band too wide for this build (band/2 must be <= 207)";
	default: return "dynamont_b200: internal error";
	}
}

const char* dyn_last_error(const dyn_aligner* A) { return A->last_error.c_str(); }

uint64_t dyn_last_fallbacks(const dyn_aligner* A) { return A->n_fallback; }
uint64_t dyn_last_lin_retries(const dyn_aligner* A) { return A->n_retry_lin; }
int dyn_last_variant(const dyn_aligner* A) { return A->last_variant; }
void dyn_last_ribbon(const dyn_aligner* A, uint64_t* out2)
{
	out2[0] = A->n_ribbon;
	out2[1] = A->n_rib_fault;
}
void dyn_ribbon_fault_reasons(const dyn_aligner* A, uint64_t* out16)
{
	for (int i = 0; i < 16; ++i) out16[i] = A->rib_reason[i];
	// [13]: records per row of the last batch x 1000, [14]: the last batch ran with two-level checkpoints,
	// [15]: cumulative reads the log2-domain ribbon kept (of those counted under their reason codes)
	out16[13] = (uint64_t)(A->rib_recs_used * 1000.0);
	out16[14] = (uint64_t)A->rib_last_two_level;
}

void dyn_last_timing(const dyn_aligner* A, double* out3)
{
	out3[0] = A->timing[0];
	out3[1] = A->timing[1];
	out3[2] = A->timing[2];
}

#ifdef DYN_DEBUG_ROWS
void dyn_debug_rows(double* p) { dyn::g_dbg_rows = p; }
#endif

int dyn_set_stream(dyn_aligner* A, void* cuda_stream)
{
	std::lock_guard<std::mutex> g(A->mu);
	A->rt.use_stream(cuda_stream);
	return 0;
}

int dyn_set_option(dyn_aligner* A, const char* key, double value)
{
	std::lock_guard<std::mutex> g(A->mu);
	const std::string k(key);
	A->options.emplace_back(k, value);
	for (dyn_aligner* l : A->lane)
		if (l) dyn_set_option(l, key, value);
	if (k == "warps_per_sm") A->warps_per_sm = std::max(0, (int)value);
	else if (k == "arith") A->arith = std::min(1, std::max(0, (int)value));
	else if (k == "variant") A->variant = std::min(N_VARIANTS - 1, std::max(-1, (int)value));
	else if (k == "fwd_fast") A->fwd_fast = value != 0.0 ? 1 : 0;
	else if (k == "thr2") A->thr2 = value;
	else if (k == "recs_per_row") A->recs_per_row = value;
	else if (k == "mem_fraction") A->mem_fraction = value;
	else if (k == "sms") A->rt.sms = std::max(1, (int)value);
	else if (k == "ribbon") A->ribbon = (value == 2.0 || value == 4.0) ? (int)value : 0;
	else if (k == "rib_guard") A->rib_guard = std::max(8, (int)value);
	else if (k == "thr_rib") A->thr_rib = value;
	else if (k == "rib_recs_per_row") A->rib_recs_per_row = value;
	else if (k == "rib_bps") A->rib_bps = (int)value;
	else if (k == "rib_two_level") A->rib_two_level = (int)value;
	else if (k == "rib_gather") A->rib_gather = (int)value;
	else if (k == "rib_log") A->rib_log = (int)value;
	else if (k == "rib_min_bw") A->rib_min_bw = (int)value;
	else return -1;
	return 0;
}

} // extern "C"

#ifdef DYN_HOST_EMU
#include "ribbon.cu"  // the emulator build is one translation unit
#endif
