// Micro-benchmark: packed FP32 (FFMA2 / FMUL2), 3-input FMNMX3 and a MUFU : FFMA2 : ALU mix shaped like one pair of
// lattice cells of the linear-domain forward-backward kernel (6 MUFU, 23 FFMA2, 8 ALU).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f32x2 tools/ubench_f32x2.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c)
{
	u64 r;
	asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b)
{
	u64 r;
	asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
	return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c)
{
	float r;
	asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
	return r;
}
__device__ __forceinline__ float ex2(float x)
{
	float y;
	asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
	return y;
}
__device__ __forceinline__ u64 pack(float lo, float hi)
{
	return ((u64)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ float lo_of(u64 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi_of(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc)
{
	u64 a[8];
	float f[8];
#pragma unroll
	for (int i = 0; i < 8; ++i)
	{
		a[i] = pack(threadIdx.x * 0.001f + i, 0.5f + i);
		f[i] = threadIdx.x * 0.002f + i;
	}
	const float c = out[0];
	const u64 cc = pack(c + 0.999f, c + 1.001f);
	unsigned acc = 0;
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < iters; ++it)
	{
		if (MODE == 0)  // 64 FFMA2
		{
#pragma unroll
			for (int r = 0; r < 8; ++r)
#pragma unroll
				for (int i = 0; i < 8; ++i) a[i] = ffma2(a[i], cc, cc);
		}
		if (MODE == 1)  // 64 FMUL2
		{
#pragma unroll
			for (int r = 0; r < 8; ++r)
#pragma unroll
				for (int i = 0; i < 8; ++i) a[i] = fmul2(a[i], cc);
		}
		if (MODE == 2)  // 64 FMNMX3
		{
#pragma unroll
			for (int r = 0; r < 8; ++r)
#pragma unroll
				for (int i = 0; i < 8; ++i) f[i] = fmax3(f[i], c + i, f[(i + 1) & 7]);
		}
		if (MODE == 3)  // 32 FFMA2 + 32 FFMA interleaved
		{
#pragma unroll
			for (int r = 0; r < 4; ++r)
#pragma unroll
				for (int i = 0; i < 8; ++i)
				{
					a[i] = ffma2(a[i], cc, cc);
					f[i] = fmaf(f[i], c, c);
				}
		}
		if (MODE == 4)  // 4 cell pairs: each 6 MUFU + 23 FFMA2 + 8 ALU (4 FMNMX, 2 FADD, 2 SHF) = 37 instr
		{
#pragma unroll
			for (int p = 0; p < 4; ++p)
			{
				u64& x = a[2 * p];
				u64& y = a[2 * p + 1];
				float& g = f[2 * p];
				float& h = f[2 * p + 1];
				// three emission-like blocks: 2 FFMA2 + 2 MUFU each
#pragma unroll
				for (int e = 0; e < 3; ++e)
				{
					u64 z = ffma2(x, cc, y);
					u64 s = ffma2(z, z, cc);
					x = pack(ex2(lo_of(s)), ex2(hi_of(s)));
					// recurrence-like: 4 FFMA2 (5 on the last block -> 6 + 17 = 23 in total)
					y = fmul2(x, y);
					u64 q = fmul2(y, cc);
					y = ffma2(q, cc, x);
					x = fmul2(x, q);
					if (e == 2)
					{
						y = fmul2(y, x);
						x = ffma2(x, cc, y);
						y = fmul2(y, cc);
						x = fmul2(x, cc);
						y = ffma2(y, cc, cc);
					}
				}
				// Viterbi-like ALU: 4 FMNMX, 2 FADD, 2 SHF
				g = fmaxf(g, lo_of(x));
				h = fmaxf(h, hi_of(x));
				const float d0 = g - lo_of(y), d1 = h - hi_of(y);
				acc = __funnelshift_l(__float_as_uint(d0), acc, 1);
				acc = __funnelshift_l(__float_as_uint(d1), acc, 1);
				g = fmaxf(g, lo_of(y));
				h = fmaxf(h, hi_of(y));
			}
		}
		if (MODE == 5)  // same mix with scalar FP32 instead of packed: 6 MUFU + 46 FFMA/FMUL + 8 ALU = 60 instr
		{
#pragma unroll
			for (int p = 0; p < 4; ++p)
			{
				float x0 = lo_of(a[2 * p]), x1 = hi_of(a[2 * p]), y0 = lo_of(a[2 * p + 1]), y1 = hi_of(a[2 * p + 1]);
				float& g = f[2 * p];
				float& h = f[2 * p + 1];
#pragma unroll
				for (int e = 0; e < 3; ++e)
				{
					float z0 = fmaf(x0, c, y0), z1 = fmaf(x1, c, y1);
					float s0 = fmaf(z0, z0, c), s1 = fmaf(z1, z1, c);
					x0 = ex2(s0); x1 = ex2(s1);
					y0 = x0 * y0; y1 = x1 * y1;
					float q0 = y0 * c, q1 = y1 * c;
					y0 = fmaf(q0, c, x0); y1 = fmaf(q1, c, x1);
					x0 = x0 * q0; x1 = x1 * q1;
					if (e == 2)
					{
						y0 *= x0; y1 *= x1;
						x0 = fmaf(x0, c, y0); x1 = fmaf(x1, c, y1);
						y0 *= c; y1 *= c;
						x0 *= c; x1 *= c;
						y0 = fmaf(y0, c, c); y1 = fmaf(y1, c, c);
					}
				}
				g = fmaxf(g, x0);
				h = fmaxf(h, x1);
				const float d0 = g - y0, d1 = h - y1;
				acc = __funnelshift_l(__float_as_uint(d0), acc, 1);
				acc = __funnelshift_l(__float_as_uint(d1), acc, 1);
				g = fmaxf(g, y0);
				h = fmaxf(h, y1);
				a[2 * p] = pack(x0, x1);
				a[2 * p + 1] = pack(y0, y1);
			}
		}
	}
	const long long t1 = clock64();
	float s = (float)acc;
#pragma unroll
	for (int i = 0; i < 8; ++i) s += lo_of(a[i]) + hi_of(a[i]) + f[i];
	out[blockIdx.x * blockDim.x + threadIdx.x + 1] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int n_instr_per_iter)
{
	float* out;
	long long* cyc;
	cudaMalloc(&out, 148 * 1024 * 4 + 4);
	cudaMemset(out, 0, 4);
	cudaMalloc(&cyc, 8);
	const int iters = 20000;
	k<MODE><<<148, threads>>>(out, 100, cyc);
	k<MODE><<<148, threads>>>(out, iters, cyc);
	long long h = 0;
	cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
	const double warps_per_smsp = threads / 32.0 / 4.0;
	const double instr = (double)iters * n_instr_per_iter * warps_per_smsp;
	printf("%-44s warps/SMSP %4.1f  cycles/warp-instr/SMSP %.3f  cycles/iter/warp-slot %.1f\n", name, warps_per_smsp,
		h / instr, (double)h / iters / warps_per_smsp);
	cudaFree(out);
	cudaFree(cyc);
}

int main()
{
	for (int threads : {128, 256, 512})
	{
		run<0>("FFMA2", threads, 64);
		run<1>("FMUL2", threads, 64);
		run<2>("FMNMX3", threads, 64);
		run<3>("FFMA2 + FFMA 1:1", threads, 64);
		run<4>("cell-pair mix packed (4 pairs: 148 instr)", threads, 148);
		run<5>("cell-pair mix scalar (4 pairs: 240 instr)", threads, 240);
	}
	return 0;
}
