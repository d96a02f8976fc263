// Forward alpha-blend: one single-warp CTA per 16x8 half tile, FOUR pixels per thread.
// Replaces the reference's FORWARD::renderCUDA (forward.cu:317-477).
//
// What is new relative to the reference:
//  * the tile's depth-ordered splats are read as CONTIGUOUS packed records
//    (binning.cu) with cp.async.bulk (TMA 1-D bulk copies) completing on an
//    mbarrier, double-buffered so batch k+1 lands while batch k is blended —
//    no id -> xy/conic dependent gather and no per-pixel global colour read;
//  * each thread blends four pixels (one from each 8x4 block of its warp's 16x8
//    half tile): one shared-memory read of a splat feeds four independent blend
//    chains (4x fewer LDS, 4-way ILP), and small CTAs fill the 148 SMs in one wave;
//  * no per-32-splat checkpoint stream (4 KB per bucket in the reference): the
//    backward kernel of this implementation replays the tile front-to-back;
//  * a conservative per-splat power threshold skips the exp for pixels far from
//    the splat (cannot change a result, see finalize_kernel).
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
// kRedo: the grid runs over the tiles of redo_list (fully sorted by then) instead of all tiles.
template <bool kCount, bool kRedo>
__global__ void __launch_bounds__(kBlendThreads, GSR_BLEND_FWD_MINBLOCKS)
blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, const float4* __restrict__ rec,
                 int W, int H,
                 const float* __restrict__ bg, float* __restrict__ out_color,
                 float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                 float* __restrict__ accum_c, uint32_t* __restrict__ max_contrib, uint32_t capacity,
                 unsigned long long* __restrict__ pair_counts, const uint32_t* __restrict__ sorted_len,
                 uint32_t* __restrict__ redo_flag, uint32_t* __restrict__ redo_list, uint32_t* __restrict__ redo_count) {
  __shared__ __align__(128) float4 sR[2][3 * kBlendBatch];  // 48-byte records, one bulk copy per batch
  __shared__ __align__(8) uint64_t s_full[2];

  const int lane = threadIdx.x;
  if (kRedo && (blockIdx.x >> 1) >= *redo_count) return;
  // 1-D grid in longest-list-first order (binning.cu: tile_order_kernel); two CTAs per tile
  const uint32_t tile_id = kRedo ? redo_list[blockIdx.x >> 1] : tile_order[blockIdx.x >> 1];
  const int warp = blockIdx.x & 1;  // upper / lower half of the tile
  // second pass: only the half tiles that stopped at the end of their sorted prefix continue (bit 0 / 1 of the flag)
  if (kRedo && !(redo_flag[tile_id] & (1u << warp))) return;
  const uint32_t grid_x = (uint32_t)(W + GSR_TILE_X - 1) / GSR_TILE_X;
  const uint32_t tile_x = tile_id % grid_x, tile_y = tile_id / grid_x;

  float pxf[kSlots], pyf[kSlots];
  uint32_t pix_id[kSlots];
  bool inside[kSlots];
#pragma unroll
  for (int q = 0; q < kSlots; ++q) {
    const uint32_t px = slot_px(tile_x, lane, q);
    const uint32_t py = slot_py(tile_y, warp, lane, q);
    inside[q] = px < (uint32_t)W && py < (uint32_t)H;
    pix_id[q] = (uint32_t)W * py + px;
    pxf[q] = (float)px;
    pyf[q] = (float)py;
  }

  const uint2 range = ranges[tile_id];
  // capacity: records the (optimistically sized) buffer holds; a list that did not fit is not
  // touched — gsr_forward re-runs the stage with the exact size in that case
  const int full = range.y <= capacity ? (int)(range.y - range.x) : 0;
  // lists of kSortSmallCap+ entries are sorted and packed only up to sorted_len[tile] (binning.cu: tile_select_kernel)
  // unless the tile has been through the full sort (kRedo; the instrumented replay: flagged tiles)
  // unless the tile has been through the full sort (the instrumented replay: flagged tiles).  The second pass
  // (kRedo, after that full sort) CONTINUES behind the prefix: list positions first .. full - 1, with the pixel states
  // the first pass left in final_T / accum_c / n_contrib (bit 31 of n_contrib: the pixel was saturated).
  int first = 0, total = full;
  if (sorted_len != nullptr && full >= kSortSmallCap) {
    const int prefix = min(full, (int)sorted_len[tile_id]);
    if (kRedo) {
      first = prefix;
      total = full - prefix;
    } else if (!(kCount && redo_flag[tile_id] != 0)) {
      total = prefix;
    }
  }
  const int rounds = (total + kBlendBatch - 1) / kBlendBatch;

  if (lane == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    fence_mbar_init();
  }
  __syncwarp();

  auto issue = [&](int round) {
    const int buf = round & 1;
    const int start = range.x + first + round * kBlendBatch;
    const int cnt = min(kBlendBatch, total - round * kBlendBatch);
    const uint32_t bytes = (uint32_t)cnt * 48u;
    mbar_expect_tx(&s_full[buf], bytes);
    bulk_g2s(&sR[buf][0], rec + 3 * (size_t)start, bytes, &s_full[buf]);
  };
  if (lane == 0 && rounds > 0) issue(0);

  bool done[kSlots];
  float T[kSlots];
  uint32_t last_contributor[kSlots];
  float C[kSlots][3];
  const size_t N = (size_t)W * H;
#pragma unroll
  for (int q = 0; q < kSlots; ++q) {
    done[q] = !inside[q];
    T[q] = 1.0f;
    last_contributor[q] = 0;
    C[q][0] = C[q][1] = C[q][2] = 0.f;
    if (kRedo && inside[q]) {
      const uint32_t nc = n_contrib[pix_id[q]];
      done[q] = (nc >> 31) != 0;
      last_contributor[q] = nc & 0x7fffffffu;
      T[q] = final_T[pix_id[q]];
      C[q][0] = accum_c[0 * N + pix_id[q]];
      C[q][1] = accum_c[1 * N + pix_id[q]];
      C[q][2] = accum_c[2 * N + pix_id[q]];
    }
  }

  unsigned long long n_visit = 0, n_gate = 0, n_blend = 0;  // kCount only
  for (int i = 0; i < rounds; ++i) {
    const int buf = i & 1;
    mbar_wait(&s_full[buf], (uint32_t)((i >> 1) & 1));
    bool all_done = true;
#pragma unroll
    for (int q = 0; q < kSlots; ++q) all_done = all_done && done[q];
    // every lane is past the previous batch here; stop when every pixel is saturated
    if (__all_sync(0xffffffffu, all_done)) break;
    if (lane == 0 && i + 1 < rounds) issue(i + 1);

    const int cnt = min(kBlendBatch, total - i * kBlendBatch);
    const uint32_t pos0 = (uint32_t)first + (uint32_t)i * kBlendBatch;
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
        // power < B.z: certainly alpha < 1/255 (conservative, see finalize_kernel)
        cand[q] = !done[q] && !(power > 0.0f) && !(power < B.z);
        any_cand = any_cand || cand[q];
        if (kCount) {
          n_visit += done[q] ? 0 : 1;   // the reference visits a pair while the pixel is not done
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
          done[q] = true;
          continue;
        }
        C[q][0] += col.x * alpha[q] * T[q];
        C[q][1] += col.y * alpha[q] * T[q];
        C[q][2] += col.z * alpha[q] * T[q];
        T[q] = test_T;
        last_contributor[q] = pos0 + j + 1;
        if (kCount) ++n_blend;
      }
      all_done = done[0] && done[1] && done[2] && done[3];
    }
  }

  bool to_be_continued = false;
  if (!kRedo && !kCount && total < full) {
    // the sorted prefix ended before every pixel was saturated: the tile takes the full sort and this half tile
    // continues from here in the second pass; its pixel states are parked in the outputs
    bool all_done = true;
#pragma unroll
    for (int q = 0; q < kSlots; ++q) all_done = all_done && done[q];
    if (!__all_sync(0xffffffffu, all_done)) {
      to_be_continued = true;
      if (lane == 0) {
        const uint32_t old = atomicOr(&redo_flag[tile_id], (1u << warp) | 4u);   // bit 2: the tile is listed
        if (!(old & 4u)) redo_list[atomicAdd(redo_count, 1u)] = tile_id;
      }
    }
  }

  uint32_t my_max = 0;
#pragma unroll
  for (int q = 0; q < kSlots; ++q) {
    if (inside[q]) {
      final_T[pix_id[q]] = T[q];
      n_contrib[pix_id[q]] = last_contributor[q] | ((to_be_continued && done[q]) ? 0x80000000u : 0u);
      out_color[0 * N + pix_id[q]] = C[q][0] + T[q] * bg[0];
      out_color[1 * N + pix_id[q]] = C[q][1] + T[q] * bg[1];
      out_color[2 * N + pix_id[q]] = C[q][2] + T[q] * bg[2];
      accum_c[0 * N + pix_id[q]] = C[q][0];
      accum_c[1 * N + pix_id[q]] = C[q][1];
      accum_c[2 * N + pix_id[q]] = C[q][2];
    }
    my_max = max(my_max, last_contributor[q]);
  }
  // deepest list position any pixel of the tile used (bounds the backward replay)
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, my_max);
  if (lane == 0) atomicMax(&max_contrib[tile_id], wmax);  // zeroed by tile_meta_kernel
  if (kCount) {
    atomicAdd(&pair_counts[0], n_visit);
    atomicAdd(&pair_counts[1], n_gate);
    atomicAdd(&pair_counts[2], n_blend);
  }
}

int launch_blend_fwd(int W, int H, const uint2* ranges, const uint32_t* tile_order, const float4* rec,
                     const float* bg, float* out_color, float* final_T,
                     uint32_t* n_contrib, float* accum_c, uint32_t* max_contrib, uint32_t capacity,
                     unsigned long long* pair_counts, const uint32_t* sorted_len, uint32_t* redo_flag,
                     uint32_t* redo_list, uint32_t* redo_count, bool redo, cudaStream_t s) {
  const unsigned grid = (unsigned)(((W + GSR_TILE_X - 1) / GSR_TILE_X) * ((H + GSR_TILE_Y - 1) / GSR_TILE_Y));
#define GSR_BLEND_FWD_ARGS                                                                                          \
  ranges, tile_order, rec, W, H, bg, out_color, final_T, n_contrib, accum_c, max_contrib, capacity, pair_counts,   \
      sorted_len, redo_flag, redo_list, redo_count
  if (redo)   // (the number of flagged tiles is only known on the device: CTAs beyond it leave at once)
    blend_fwd_kernel<false, true><<<2 * grid, kBlendThreads, 0, s>>>(GSR_BLEND_FWD_ARGS);
  else if (pair_counts != nullptr)
    blend_fwd_kernel<true, false><<<2 * grid, kBlendThreads, 0, s>>>(GSR_BLEND_FWD_ARGS);
  else
    blend_fwd_kernel<false, false><<<2 * grid, kBlendThreads, 0, s>>>(GSR_BLEND_FWD_ARGS);
#undef GSR_BLEND_FWD_ARGS
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
