// Ribbon kernels (dp_ribbon.cuh): kernel entry points and launcher.  Built by nvcc for sm_100a; under the test emulator
// this file is included by engine.cu (one translation unit, g++).
#include "dp_ribbon.cuh"
#include "ribbon.h"

namespace dyn
{
namespace rib
{

namespace
{

constexpr int WPB = 4;  // warps per CTA; the warps of a CTA are independent (each pulls its own reads)

#ifndef DYN_RIB_GR
#define DYN_RIB_GR 16  // rows per group of the 2-columns-per-lane kernels (8: the first version of round 2)
#endif
using RC2 = RCfg<2, false, DYN_RIB_GR>;
using RC4 = RCfg<4>;
using RC2L = RCfg<2, true, LOG_GROUP_ROWS>;  // the log2-domain tier behind the linear-domain ribbon (align, records-free layout, two-level checkpoints)
constexpr int BPSL = 4;
constexpr int BPS2 = 5;  // 20 resident warps per SM: <= 96 registers per thread, no spills (measured: 6 -> 80 registers: -4 %, 8 -> 64: -13 %)
constexpr int BPS4 = 4;  // 16 resident warps per SM: <= 128 registers per thread

template <class RC, int MODE, bool TL>
DYN_DEV void worker(const BatchArgs& args, unsigned char* smem_raw, int lane, unsigned slot)
{
	if (slot >= args.n_slots) return;
	const SlotScratch sc = args.slots[slot];
	while (true)
	{
		uint32_t i = 0;
		if (lane == 0) i = atomicAdd(args.queue, 1u);
		i = __shfl_sync(FULL, i, 0);
		if (i >= args.n_reads) break;
		const uint32_t ridx = args.order[i];
		const ReadDesc rd = args.reads[ridx];
		if (rd.status != ST_OK)
		{
			if (lane == 0)
			{
				ReadOut o;
				o.Z = 0.0; o.dZ = 0.0; o.nrec = 0; o.status = rd.status; o.xi_m = 0.0; o.xi_e = 0.0;
				args.out[ridx] = o;
			}
			continue;
		}
		ribbon_read<RC, MODE, TL>(args, rd, ridx, sc, smem_raw, lane);
		__syncwarp();
	}
}

#ifndef DYN_HOST_EMU
template <class RC, int MODE, int BPS, bool TL>
__global__ void __launch_bounds__(32 * WPB, BPS) k_ribbon(BatchArgs args)
{
	extern __shared__ __align__(16) unsigned char smem_all[];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	worker<RC, MODE, TL>(args, smem_all + (size_t)wid * RC::SMEM_BYTES, lane, blockIdx.x * WPB + wid);
}

template <class RC, int BPS, bool TL>
int launch_t(cudaStream_t stream, const BatchArgs& args, unsigned n_warps, int mode)
{
	// (the dynamic shared memory of a CTA, 4 x 2.3 / 4.6 KB, is below the 48 KB that needs no opt-in)
	const size_t smem = RC::SMEM_BYTES * WPB;
	const unsigned ctas = (n_warps + WPB - 1) / WPB;
	if (mode == 0) k_ribbon<RC, 0, BPS, false><<<ctas, 32 * WPB, 0, stream>>>(args);
	else if (mode == 1) k_ribbon<RC, 1, BPS, TL><<<ctas, 32 * WPB, smem, stream>>>(args);
	else if (mode == 3)
	{
		if constexpr (TL) k_ribbon<RC, 3, BPS, true><<<ctas, 32 * WPB, smem, stream>>>(args);
		else return (int)cudaErrorInvalidValue;  // the records-free layout is built for two-level checkpoints only
	}
	else k_ribbon<RC, 2, BPS, TL><<<ctas, 32 * WPB, smem, stream>>>(args);
	return (int)cudaGetLastError();
}
#else
template <class RC, int BPS, bool TL>
int launch_t(void*, const BatchArgs& args, unsigned n_warps, int mode)
{
	const size_t smem = RC::SMEM_BYTES;
	if (mode == 0) simt::launch(n_warps, smem, [&]() { worker<RC, 0, false>(args, simt::dyn_smem(), threadIdx.x, blockIdx.x); });
	else if (mode == 1) simt::launch(n_warps, smem, [&]() { worker<RC, 1, TL>(args, simt::dyn_smem(), threadIdx.x, blockIdx.x); });
	else if (mode == 3)
	{
		if constexpr (TL) simt::launch(n_warps, smem, [&]() { worker<RC, 3, true>(args, simt::dyn_smem(), threadIdx.x, blockIdx.x); });
		else return -1;
	}
	else simt::launch(n_warps, smem, [&]() { worker<RC, 2, TL>(args, simt::dyn_smem(), threadIdx.x, blockIdx.x); });
	return 0;
}
#endif

template <class RC>
void fill(Geometry& g, int bps)
{
	g.cpl = RC::CPL;
	g.hw = RC::HW;
	g.ck = RC::GR;
	g.ckf = RC::CKF;
	g.hdrw = RC::HDRW;
	g.recf = RC::RECF;
	g.smem_per_warp = RC::SMEM_BYTES;
	g.warps_per_block = WPB;
	g.blocks_per_sm = bps;
}

} // namespace

bool geometry(int cpl, int bps, Geometry& g)
{
	if (cpl == 2) fill<RC2>(g, (bps == 6 || bps == 8) ? bps : BPS2);
	else if (cpl == 4) fill<RC4>(g, BPS4);
	else return false;
	return true;
}

int launch(void* stream, const BatchArgs& args, unsigned n_warps, int mode, int cpl, int bps, bool two_level, bool log_domain)
{
#ifndef DYN_HOST_EMU
	cudaStream_t s = (cudaStream_t)stream;
#else
	void* s = stream;
#endif
	if (log_domain)
	{
		if ((mode != 3 && mode != 2) || !two_level) return -1;
#ifndef DYN_HOST_EMU
		const unsigned ctas = (n_warps + WPB - 1) / WPB;
		if (mode == 3) k_ribbon<RC2L, 3, BPSL, true><<<ctas, 32 * WPB, RC2L::SMEM_BYTES * WPB, s>>>(args);
		else k_ribbon<RC2L, 2, BPSL, true><<<ctas, 32 * WPB, RC2L::SMEM_BYTES * WPB, s>>>(args);
		return (int)cudaGetLastError();
#else
		if (mode == 3) simt::launch(n_warps, RC2L::SMEM_BYTES, [&]() { worker<RC2L, 3, true>(args, simt::dyn_smem(), threadIdx.x, blockIdx.x); });
		else simt::launch(n_warps, RC2L::SMEM_BYTES, [&]() { worker<RC2L, 2, true>(args, simt::dyn_smem(), threadIdx.x, blockIdx.x); });
		return 0;
#endif
	}
	if (cpl == 2)
	{
		if (two_level) return launch_t<RC2, BPS2, true>(s, args, n_warps, mode);
#ifndef DYN_HOST_EMU
		if (bps == 6) return launch_t<RC2, 6, false>(s, args, n_warps, mode);
		if (bps == 8) return launch_t<RC2, 8, false>(s, args, n_warps, mode);
#endif
		return launch_t<RC2, BPS2, false>(s, args, n_warps, mode);
	}
	if (cpl == 4) return two_level ? launch_t<RC4, BPS4, true>(s, args, n_warps, mode) : launch_t<RC4, BPS4, false>(s, args, n_warps, mode);
	(void)bps;
	return -1;
}

} // namespace rib
} // namespace dyn
