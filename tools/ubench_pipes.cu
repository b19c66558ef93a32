// Micro-benchmark: issue cost of FADD / FFMA / FMNMX / MUFU.EX2 / MUFU.LG2 and of their mixes on one SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes tools/ubench_pipes.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc)
{
	float a[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
	const float c = out[0];
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < iters; ++it)
	{
#pragma unroll
		for (int r = 0; r < 8; ++r)
		{
#pragma unroll
			for (int i = 0; i < 8; ++i)
			{
				if (MODE == 0) a[i] = a[i] + c;                          // FADD
				if (MODE == 1) a[i] = fmaf(a[i], c, c);                  // FFMA
				if (MODE == 2) a[i] = fmaxf(a[i], c + i);                // FMNMX (+FADD folded? c+i is loop-invariant)
				if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));   // MUFU only
				if (MODE == 4)                                           // 1 MUFU : 4 FADD
				{
					if ((i & 3) == 0 && (r & 1) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
					else a[i] = a[i] + c;
				}
				if (MODE == 5)                                           // 1 MUFU : 7 FADD (like the DP row: 26 of ~190)
				{
					if (i == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
					else a[i] = a[i] + c;
				}
			}
		}
	}
	const long long t1 = clock64();
	float s = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) s += a[i];
	out[blockIdx.x * blockDim.x + threadIdx.x + 1] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads, int n_instr_per_iter, int n_mufu_per_iter)
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
	const double instr = (double)iters * n_instr_per_iter * warps_per_smsp;  // warp-instructions per SMSP
	printf("%-28s threads/SM %4d  cycles/warp-instr/SMSP %.3f   (MUFU share %.2f)\n", name, threads, h / instr,
		(double)n_mufu_per_iter / n_instr_per_iter);
	cudaFree(out);
	cudaFree(cyc);
}

int main()
{
	for (int threads : {128, 256, 512})
	{
		run<0>("FADD", threads, 64, 0);
		run<1>("FFMA", threads, 64, 0);
		run<2>("FMNMX", threads, 64, 0);
		run<3>("MUFU.EX2", threads, 64, 64);
		run<4>("1 MUFU : 7 FADD (8/64)", threads, 64, 8);
		run<5>("1 MUFU : 7 FADD (per group)", threads, 64, 8);
	}
	return 0;
}
