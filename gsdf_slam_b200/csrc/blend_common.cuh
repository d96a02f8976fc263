// Shared shape of the two blend kernels: one single-warp CTA per 16x8 HALF tile (two CTAs per
// 16x16 tile, so the work units are short and the last wave of the launch stays small; no
// CTA-wide barrier exists in the kernels), each thread owns four pixels — the pixel at (lane&7, lane>>3) of each
// of the four 8x4 blocks of the half tile.  A "slot" q is one of those blocks, so the 32
// pixels of a slot are spatially compact: a warp-wide vote per slot tells cheaply whether a
// splat touches that block at all.
// Measured alternatives (round 2): persistent grids pulling half tiles from an atomic cursor — with
// 4-warp CTAs (+5 % / +12 % time: the compiler loses the knowledge that the CTA is ONE warp, which keeps the
// per-tile address arithmetic on the uniform datapath) and with single-warp CTAs (+2 %) — and fewer resident CTAs
// per SM (24: +10 %, 16: +28 %).  The hardware's CTA scheduler over the longest-list-first order stays.
// Also measured: the per-pixel arithmetic of pixel PAIRS issued as packed fp32x2 instructions (FFMA2 / FMUL2,
// sm_100; bit-identical results incl. a packed restatement of expf): 17 % / 12 % fewer SASS instructions per
// loop iteration, but forward 180 vs 179 us and backward 285 vs 265 us at cfg2 — the selects that replace the
// per-pixel branches load the half-rate ALU pipe and FFMA2 occupies the FMA pipe for two cycles
// (tools/microbench/issue_bench.cu) — so the scalar kernels stay.
#pragma once
#include <stdint.h>

#include "gsr_common.cuh"

// minimum resident CTAs per SM the compiler must allow (bounds registers per thread)
#ifndef GSR_BLEND_FWD_MINBLOCKS
#define GSR_BLEND_FWD_MINBLOCKS 1
#endif
#ifndef GSR_BLEND_BWD_MINBLOCKS
#define GSR_BLEND_BWD_MINBLOCKS 1
#endif

namespace gsr {

constexpr int kBlendThreads = 32;   // one warp = one 16x8 half tile = one CTA
constexpr int kSlots = 4;
constexpr int kBlendBatch = 64;    // splat records staged per bulk copy (3 x 1 KB)

__device__ __forceinline__ uint32_t slot_px(uint32_t tile_x, int lane, int q) {
  return tile_x * GSR_TILE_X + (q & 1) * 8 + (lane & 7);
}
__device__ __forceinline__ uint32_t slot_py(uint32_t tile_y, int warp, int lane, int q) {
  return tile_y * GSR_TILE_Y + warp * 8 + (q >> 1) * 4 + (lane >> 3);
}

}  // namespace gsr
