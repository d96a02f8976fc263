// Forward alpha-blend.  Replaces the reference's FORWARD::renderCUDA (forward.cu:317-477).
//
// Work unit = one 16x8 HALF tile handled by ONE warp, four pixels per thread (blend_common.cuh).  The
// kernel is PERSISTENT: a grid of 4-warp CTAs (one warp per SM sub-partition, so the four schedulers of
// an SM always hold the same number of warps) pulls units from an atomic cursor over the
// longest-list-first tile order until none is left.
//
// What is new relative to the reference:
//  * the tile's depth-ordered splats are read as CONTIGUOUS packed records
//    (binning.cu) with cp.async.bulk (TMA 1-D bulk copies) completing on an
//    mbarrier, double-buffered so batch k+1 lands while batch k is blended —
//    no id -> xy/conic dependent gather and no per-pixel global colour read;
//  * each thread blends four pixels (one from each 8x4 block of its warp's 16x8
//    half tile): one shared-memory read of a splat feeds four independent blend
//    chains (4x fewer LDS, 4-way ILP);
//  * no per-32-splat checkpoint stream (4 KB per bucket in the reference): the
//    backward kernel of this implementation replays the tile front-to-back;
//  * a conservative per-splat power threshold skips the exp for pixels far from
//    the splat (cannot change a result, see preprocess.cu);
//  * a finished pixel costs nothing: its upper power bound `hi` drops from 0 to -inf, so the test
//    `power <= hi` the loop makes anyway rejects every later splat (no per-pixel flag to test).
// The per-pixel arithmetic keeps the reference's operation order so n_contrib,
// final_T and colour replay bit-for-bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "blend_common.cuh"
#include "gsr_common.cuh"
#include "tma.cuh"

namespace gsr {

// kCount: instrumented instantiation behind gsr_export_pair_counts (bench.py's roofline numerators): counts the
// (splat, pixel) pairs visited / passing the power gate / blended.  The production instantiation carries none of it.
template <bool kCount>
__global__ void __launch_bounds__(kBlendCtaThreads, GSR_BLEND_FWD_MINBLOCKS)
blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, const float4* __restrict__ rec,
                 int W, int H,
                 const float* __restrict__ bg, float* __restrict__ out_color,
                 float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                 float* __restrict__ accum_c, uint32_t* __restrict__ max_contrib, uint32_t capacity,
                 uint32_t* __restrict__ cursor, uint32_t num_units, unsigned long long* __restrict__ pair_counts) {
  __shared__ __align__(128) float4 sR_all[kBlendWarps][2][3 * kBlendBatch];  // 48-byte records, one bulk copy per batch
  __shared__ __align__(8) uint64_t s_full_all[kBlendWarps][2];

  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  float4 (*sR)[3 * kBlendBatch] = sR_all[wid];
  uint64_t* s_full = s_full_all[wid];
  const uint32_t grid_x = (uint32_t)(W + GSR_TILE_X - 1) / GSR_TILE_X;
  const size_t N = (size_t)W * H;
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];

  if (lane == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  uint32_t phase = 0;  // bit b: parity the next wait on buffer b expects
  unsigned long long n_visit = 0, n_gate = 0, n_blend = 0;  // kCount only

  for (;;) {
    // next unit: half tile `unit & 1` of the tile at position `unit >> 1` of the longest-list-first order
    uint32_t unit = 0;
    if (lane == 0) unit = atomicAdd(cursor, 1u);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= num_units) break;
    const uint32_t tile_id = tile_order[unit >> 1];
    const int half = (int)(unit & 1u);  // upper / lower half of the tile
    const uint32_t tile_x = tile_id % grid_x, tile_y = tile_id / grid_x;

    float pxf[kSlots], pyf[kSlots];
    uint32_t pix_id[kSlots];
    bool inside[kSlots];
#pragma unroll
    for (int q = 0; q < kSlots; ++q) {
      const uint32_t px = slot_px(tile_x, lane, q);
      const uint32_t py = slot_py(tile_y, half, lane, q);
      inside[q] = px < (uint32_t)W && py < (uint32_t)H;
      pix_id[q] = (uint32_t)W * py + px;
      pxf[q] = (float)px;
      pyf[q] = (float)py;
    }

    const uint2 range = ranges[tile_id];
    // capacity: records the (optimistically sized) buffer holds; a list that did not fit is not
    // touched — gsr_forward re-runs the stage with the exact size in that case
    const int total = range.y <= capacity ? (int)(range.y - range.x) : 0;
    const int rounds = (total + kBlendBatch - 1) / kBlendBatch;

    auto issue = [&](int round) {
      const int buf = round & 1;
      const int start = range.x + round * kBlendBatch;
      const int cnt = min(kBlendBatch, total - round * kBlendBatch);
      const uint32_t bytes = (uint32_t)cnt * 48u;
      mbar_expect_tx(&s_full[buf], bytes);
      bulk_g2s(&sR[buf][0], rec + 3 * (size_t)start, bytes, &s_full[buf]);
    };
    if (lane == 0 && rounds > 0) issue(0);

    // hi[q]: largest power a splat may have at pixel q and still be evaluated: 0 while the pixel is
    // live (the reference skips power > 0), -inf once it is finished or lies outside the image
    float hi[kSlots];
    float T[kSlots];
    uint32_t last_contributor[kSlots];
    float C[kSlots][3];
#pragma unroll
    for (int q = 0; q < kSlots; ++q) {
      hi[q] = inside[q] ? 0.0f : -INFINITY;
      T[q] = 1.0f;
      last_contributor[q] = 0;
      C[q][0] = C[q][1] = C[q][2] = 0.f;
    }
    // every pixel of this thread finished (re-evaluated only when one of them finishes)
    bool all_done = fmaxf(fmaxf(hi[0], hi[1]), fmaxf(hi[2], hi[3])) < 0.0f;

    for (int i = 0; i < rounds; ++i) {
      const int buf = i & 1;
      mbar_wait(&s_full[buf], (phase >> buf) & 1u);
      phase ^= 1u << buf;
      // every lane is past the previous batch here; stop when every pixel is saturated
      if (__all_sync(0xffffffffu, all_done)) break;
      if (lane == 0 && i + 1 < rounds) issue(i + 1);

      const int cnt = min(kBlendBatch, total - i * kBlendBatch);
      const uint32_t pos0 = (uint32_t)i * kBlendBatch;
      for (int j = 0; !all_done && j < cnt; j++) {
        const float4 A = sR[buf][3 * j];
        const float4 B = sR[buf][3 * j + 1];
        const float2 xy = {A.x, A.y};
        const float4 con_o = {A.z, A.w, B.x, B.y};
        // gate of all four pixels first (independent chains, one branch for the common miss)
        float pw[kSlots];
        bool cand[kSlots];
        bool any_cand = false;
#pragma unroll
        for (int q = 0; q < kSlots; ++q) {
          float2 d = {xy.x - pxf[q & 1], xy.y - pyf[q & 2]};
          float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
          pw[q] = power;
          // power < B.z: certainly alpha < 1/255 (conservative, see preprocess.cu)
          cand[q] = !(power > hi[q]) && !(power < B.z);
          any_cand = any_cand || cand[q];
          if (kCount) {
            n_visit += hi[q] == 0.0f ? 1 : 0;   // the reference visits a pair while the pixel is not done
            n_gate += cand[q] ? 1 : 0;
          }
        }
        if (!any_cand) continue;
        // the four exp / alpha chains are evaluated together (4-way ILP, no branches); a chain
        // of a pixel that is not a candidate computes a value nobody uses
        float alpha[kSlots];
        bool live[kSlots];
#pragma unroll
        for (int q = 0; q < kSlots; ++q) {
          alpha[q] = min(0.99f, con_o.w * exp(pw[q]));
          live[q] = cand[q] && !(alpha[q] < 1.0f / 255.0f);
        }
        if (!(live[0] || live[1] || live[2] || live[3])) continue;
        const float4 col = sR[buf][3 * j + 2];
#pragma unroll
        for (int q = 0; q < kSlots; ++q) {
          if (!live[q]) continue;
          const float test_T = T[q] * (1 - alpha[q]);
          if (test_T < 0.0001f) {
            hi[q] = -INFINITY;  // done
            all_done = fmaxf(fmaxf(hi[0], hi[1]), fmaxf(hi[2], hi[3])) < 0.0f;
            continue;
          }
          C[q][0] += col.x * alpha[q] * T[q];
          C[q][1] += col.y * alpha[q] * T[q];
          C[q][2] += col.z * alpha[q] * T[q];
          T[q] = test_T;
          last_contributor[q] = pos0 + j + 1;
          if (kCount) ++n_blend;
        }
      }
    }

    uint32_t my_max = 0;
#pragma unroll
    for (int q = 0; q < kSlots; ++q) {
      if (inside[q]) {
        final_T[pix_id[q]] = T[q];
        n_contrib[pix_id[q]] = last_contributor[q];
        out_color[0 * N + pix_id[q]] = C[q][0] + T[q] * bg0;
        out_color[1 * N + pix_id[q]] = C[q][1] + T[q] * bg1;
        out_color[2 * N + pix_id[q]] = C[q][2] + T[q] * bg2;
        accum_c[0 * N + pix_id[q]] = C[q][0];
        accum_c[1 * N + pix_id[q]] = C[q][1];
        accum_c[2 * N + pix_id[q]] = C[q][2];
      }
      my_max = max(my_max, last_contributor[q]);
    }
    // deepest list position any pixel of the tile used (bounds the backward replay)
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, my_max);
    if (lane == 0) atomicMax(&max_contrib[tile_id], wmax);  // zeroed by tile_meta_kernel
  }
  if (kCount) {
    atomicAdd(&pair_counts[0], n_visit);
    atomicAdd(&pair_counts[1], n_gate);
    atomicAdd(&pair_counts[2], n_blend);
  }
}

int launch_blend_fwd(int W, int H, const uint2* ranges, const uint32_t* tile_order, const float4* rec,
                     const float* bg, float* out_color, float* final_T,
                     uint32_t* n_contrib, float* accum_c, uint32_t* max_contrib, uint32_t capacity,
                     uint32_t* cursor, unsigned long long* pair_counts, cudaStream_t s) {
  const unsigned tiles = (unsigned)(((W + GSR_TILE_X - 1) / GSR_TILE_X) * ((H + GSR_TILE_Y - 1) / GSR_TILE_Y));
  const unsigned units = 2 * tiles;
  if (pair_counts != nullptr) {
    const unsigned grid = blend_persistent_grid(units, blend_fwd_kernel<true>);
    blend_fwd_kernel<true><<<grid, kBlendCtaThreads, 0, s>>>(ranges, tile_order, rec, W, H, bg, out_color, final_T,
                                                           n_contrib, accum_c, max_contrib, capacity, cursor, units,
                                                           pair_counts);
  } else {
    const unsigned grid = blend_persistent_grid(units, blend_fwd_kernel<false>);
    blend_fwd_kernel<false><<<grid, kBlendCtaThreads, 0, s>>>(ranges, tile_order, rec, W, H, bg, out_color, final_T,
                                                            n_contrib, accum_c, max_contrib, capacity, cursor, units,
                                                            nullptr);
  }
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
