// Shared shape of the two blend kernels.  Work unit: one 16x8 HALF tile, handled by ONE warp (two units per
// 16x16 tile, so the units are short); each thread owns four pixels — the pixel at (lane&7, lane>>3) of each
// of the four 8x4 blocks of the half tile.  A "slot" q is one of those blocks, so the 32
// pixels of a slot are spatially compact.  The kernels are PERSISTENT: CTAs of kBlendWarps warps — one per SM
// sub-partition, so the four warp schedulers of an SM always hold the same number of warps — fill the GPU once
// and every warp pulls units from an atomic cursor over the longest-list-first tile order (binning.cu:
// tile_meta_kernel) until none is left; the warps of a CTA share nothing but the launch (no CTA-wide barrier).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr_common.cuh"

// minimum resident CTAs per SM the compiler must allow (bounds registers per thread)
#ifndef GSR_BLEND_FWD_MINBLOCKS
#define GSR_BLEND_FWD_MINBLOCKS 8   // 8 CTAs x 4 warps: 64 registers per thread
#endif
#ifndef GSR_BLEND_BWD_MINBLOCKS
#define GSR_BLEND_BWD_MINBLOCKS 7   // 7 CTAs x 4 warps: 72 registers per thread (8 would spill)
#endif
#ifndef GSR_BLEND_BATCH
#define GSR_BLEND_BATCH 64
#endif

namespace gsr {

constexpr int kBlendThreads = 32;   // one warp = one 16x8 half tile
constexpr int kBlendWarps = 4;      // warps per persistent CTA (one per SM sub-partition)
constexpr int kBlendCtaThreads = kBlendThreads * kBlendWarps;
constexpr int kSlots = 4;
constexpr int kBlendBatch = GSR_BLEND_BATCH;    // splat records staged per bulk copy (64: 3 x 1 KB)

__device__ __forceinline__ uint32_t slot_px(uint32_t tile_x, int lane, int q) {
  return tile_x * GSR_TILE_X + (q & 1) * 8 + (lane & 7);
}
__device__ __forceinline__ uint32_t slot_py(uint32_t tile_y, int warp, int lane, int q) {
  return tile_y * GSR_TILE_Y + warp * 8 + (q >> 1) * 4 + (lane >> 3);
}

// Grid of a persistent blend kernel: every CTA slot of the device (occupancy x SM count, queried once per
// device), but no more CTAs than there are units of work.
template <typename Kernel>
inline unsigned blend_persistent_grid(unsigned units, Kernel kernel) {
  constexpr int kMaxDev = 64;
  static int slots[kMaxDev] = {0};   // benign race: every writer stores the same value
  int dev = 0;
  cudaGetDevice(&dev);
  int n = (dev >= 0 && dev < kMaxDev) ? slots[dev] : 0;
  if (n == 0) {
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlendCtaThreads, 0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    n = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
    if (dev >= 0 && dev < kMaxDev) slots[dev] = n;
  }
  const unsigned want = (units + kBlendWarps - 1) / kBlendWarps;
  return want < (unsigned)n ? (want > 0 ? want : 1u) : (unsigned)n;
}

}  // namespace gsr
