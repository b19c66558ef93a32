// TEST INFRASTRUCTURE — a tiny SIMT emulator so that the CUDA kernels in dynamont_b200/csrc can be unit-tested
// on a machine without a GPU (the dev container).  It is NOT a product path: the shipped library is built by
// nvcc for sm_100a only and has no CPU fallback; this header is force-included (-include) only by
// tests/emu/build_emu.py when it compiles the same .cu sources with g++ into tests/emu/_build/.
//
// Model: a CTA is one warp of 32 lanes; every lane is a ucontext fibre; warp collectives (__shfl_sync,
// __ballot_sync, __any_sync, __syncwarp) are rendezvous points at which the fibres of a warp are switched
// round-robin.  CTAs of a grid run one after the other.  MUFU approximations map to libm.
#pragma once
#ifndef DYN_HOST_EMU
#define DYN_HOST_EMU 1
#endif

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__

using std::max;
using std::min;

struct uint4
{
	unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct uint2
{
	unsigned x, y;
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct dim3e
{
	unsigned x = 1, y = 1, z = 1;
};

namespace simt
{

constexpr int WARP = 32;
constexpr size_t STACK_BYTES = 256 * 1024;

struct WarpRt
{
	ucontext_t sched;
	ucontext_t ctx[WARP];
	bool done[WARP];
	int cur = 0;
	// rendezvous
	int arrived = 0;
	unsigned gen = 0;
	uint64_t xch[WARP];
	unsigned vote = 0;
	unsigned vote_out = 0;
	std::function<void()> body;
	unsigned block = 0, grid = 1;
	std::vector<char> stacks;
	unsigned char* smem = nullptr;
};

inline WarpRt*& rt()
{
	static thread_local WarpRt* p = nullptr;
	return p;
}

inline void yield_to_sched()
{
	WarpRt* w = rt();
	swapcontext(&w->ctx[w->cur], &w->sched);
}

// all 32 lanes must call; returns after everyone arrived
inline void rendezvous()
{
	WarpRt* w = rt();
	const unsigned g = w->gen;
	if (++w->arrived == WARP)
	{
		w->arrived = 0;
		++w->gen;
		return;
	}
	while (w->gen == g) yield_to_sched();
}

inline void fibre_main()
{
	WarpRt* w = rt();
	w->body();
	w->done[w->cur] = true;
	swapcontext(&w->ctx[w->cur], &w->sched);
}

inline void run_warp(WarpRt& w)
{
	rt() = &w;
	w.stacks.resize(STACK_BYTES * WARP);
	w.arrived = 0;
	for (int l = 0; l < WARP; ++l)
	{
		w.done[l] = false;
		getcontext(&w.ctx[l]);
		w.ctx[l].uc_stack.ss_sp = w.stacks.data() + STACK_BYTES * l;
		w.ctx[l].uc_stack.ss_size = STACK_BYTES;
		w.ctx[l].uc_link = &w.sched;
		makecontext(&w.ctx[l], (void (*)())fibre_main, 0);
	}
	int remaining = WARP;
	while (remaining)
	{
		remaining = 0;
		for (int l = 0; l < WARP; ++l)
		{
			if (w.done[l]) continue;
			w.cur = l;
			swapcontext(&w.sched, &w.ctx[l]);
			if (!w.done[l]) ++remaining;
		}
	}
	rt() = nullptr;
}

struct Idx
{
	unsigned x, y = 0, z = 0;
};

inline Idx thread_idx() { return Idx{(unsigned)rt()->cur}; }
inline Idx block_idx() { return Idx{rt()->block}; }
inline Idx grid_dim() { return Idx{rt()->grid}; }
inline Idx block_dim() { return Idx{32u}; }

// launch a "kernel": grid CTAs of exactly one warp each, smem bytes of dynamic shared memory per CTA
template <typename F>
inline void launch(unsigned grid, size_t smem_bytes, F f)
{
	for (unsigned b = 0; b < grid; ++b)
	{
		WarpRt w;
		std::vector<unsigned char> sm(smem_bytes + 64);
		w.smem = sm.data() + (64 - ((uintptr_t)sm.data() & 63)) % 64;
		w.block = b;
		w.grid = grid;
		w.body = f;
		run_warp(w);
	}
}

inline unsigned char* dyn_smem() { return rt()->smem; }

} // namespace simt

#define threadIdx (simt::thread_idx())
#define blockIdx (simt::block_idx())
#define gridDim (simt::grid_dim())
#define blockDim (simt::block_dim())

// ---- warp collectives (full mask only) ----------------------------------------------------------------
template <typename T>
inline T __shfl_sync(unsigned, T v, int src)
{
	static_assert(sizeof(T) <= 8, "shuffle width");
	simt::WarpRt* w = simt::rt();
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	w->xch[w->cur] = raw;
	simt::rendezvous();
	raw = w->xch[src & 31];
	simt::rendezvous();
	T out;
	memcpy(&out, &raw, sizeof(T));
	return out;
}

inline unsigned __ballot_sync(unsigned, int pred)
{
	simt::WarpRt* w = simt::rt();
	if (w->arrived == 0) w->vote = 0;
	if (pred) w->vote |= 1u << w->cur;
	// the last arriver publishes
	if (w->arrived == simt::WARP - 1) w->vote_out = w->vote;
	simt::rendezvous();
	const unsigned r = w->vote_out;
	simt::rendezvous();
	return r;
}

inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
	simt::WarpRt* w = simt::rt();
	w->xch[w->cur] = v;
	simt::rendezvous();
	unsigned s = 0;
	for (int l = 0; l < simt::WARP; ++l) s += (unsigned)w->xch[l];
	simt::rendezvous();
	return s;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline void __syncwarp(unsigned = 0xffffffffu) { simt::rendezvous(); }
inline void __syncthreads() { simt::rendezvous(); }
inline void __threadfence_block() {}
inline void __threadfence() {}

// ---- intrinsics --------------------------------------------------------------------------------------------
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline float __int_as_float(int i)
{
	float f;
	memcpy(&f, &i, 4);
	return f;
}
inline int __float_as_int(float f)
{
	int i;
	memcpy(&i, &f, 4);
	return i;
}
struct float2
{
	float x, y;
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct float4
{
	float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift)
{
	shift &= 31;
	return shift ? (hi << shift) | (lo >> (32 - shift)) : hi;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{
	shift &= 31;
	return shift ? (lo >> shift) | (hi << (32 - shift)) : lo;
}
inline unsigned __float_as_uint(float f)
{
	unsigned u;
	memcpy(&u, &f, 4);
	return u;
}
inline float __uint_as_float(unsigned u)
{
	float f;
	memcpy(&f, &u, 4);
	return f;
}
inline int __double2loint(double d)
{
	uint64_t u;
	memcpy(&u, &d, 8);
	return (int)(uint32_t)u;
}
inline int __double2hiint(double d)
{
	uint64_t u;
	memcpy(&u, &d, 8);
	return (int)(uint32_t)(u >> 32);
}
inline double __hiloint2double(int hi, int lo)
{
	const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
	double d;
	memcpy(&d, &u, 8);
	return d;
}
inline unsigned long long __double2ull_rz(double d) { return (unsigned long long)d; }
inline double __dmul_rn(double a, double b) { return a * b; }

template <typename T>
inline T atomicAdd(T* p, T v)
{
	const T old = *p;
	*p = old + v;
	return old;
}
template <typename T>
inline T atomicMin(T* p, T v)
{
	const T old = *p;
	if (v < old) *p = v;
	return old;
}
template <typename T>
inline T __ldg(const T* p) { return *p; }
