// Launcher interface of the ribbon kernels (dp_ribbon.cuh); its own translation unit (ribbon.cu) so that it builds in
// parallel with engine.cu.
#pragma once

#include <stddef.h>
#include <stdint.h>

namespace dyn
{
struct BatchArgs;
namespace rib
{

constexpr int LOG_GROUP_ROWS = 8;  // rows per group of the log2-domain ribbon (its scratch is sized on the host from this)

struct Geometry
{
	int cpl;              // lattice columns per lane
	int hw;               // half window: live columns of a row are [mid - hw, mid + hw]
	int ck;               // checkpoint spacing (rows)
	int ckf;              // floats per checkpoint
	int hdrw;             // 32-bit words per row header
	int recf;             // floats per lane record
	size_t smem_per_warp;
	int warps_per_block;
	int blocks_per_sm;    // resident CTAs per SM the kernels are built for (mode 1 / 2)
};

// cpl: 2 or 4.  Returns false for an unsupported width.
bool geometry(int cpl, int bps, Geometry& g);  // bps: resident CTAs per SM (0 = default)

// n_warps resident warps (one scratch slot each, args.slots[0 .. n_warps)); mode 0: Z only, 1: align, 2: train,
// 3: align with the records-free scratch layout (row header = decision words only; the path posteriors are evaluated by a
// second forward sweep after the traceback; needs two_level).
// Returns 0 or the cudaError_t of the launch.
// two_level: checkpoints of every 8th group only (dp_ribbon.cuh SG), the rest replayed into a per-warp ring in pass 2
// log_domain: the log2-domain build of the same passes (mode 3 or 2, two_level only; 2 columns per lane whatever cpl says)
int launch(void* stream, const BatchArgs& args, unsigned n_warps, int mode, int cpl, int bps, bool two_level, bool log_domain = false);

} // namespace rib
} // namespace dyn
