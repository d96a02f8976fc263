// Binning: turn the depth-sorted Gaussians into per-tile, depth-ordered lists.
//
// The reference builds one 64-bit key (tile<<32 | depth bits) per (Gaussian,
// tile) overlap and radix-sorts all R of them on 32+log2(T) bits
// (duplicateWithKeys + cub SortPairs + identifyTileRanges,
// rasterizer_impl.cu:139-246, 476-500).  Here the same total order is produced
// with far less traffic by splitting the LSD sort at bit 32:
//   * the low 32 key bits (depth) are identical for every duplicate of a
//     Gaussian, so those four digit passes run on the P un-duplicated Gaussians
//     (preprocess emits (depth<<32 | idx); launch_sort_u64 orders them);
//   * duplicates are then emitted in that order (emit_kernel) and only the tile
//     digits (bits 32..32+log2 T) are sorted on the R duplicates — LSD radix sort
//     is stable, so the result is exactly the order of sorting the full key,
//     including ties (equal tile and depth keep ascending Gaussian index).
// finalize_kernel then derives the tile ranges and writes the packed,
// tile-contiguous splat records that the blend kernels stream with bulk copies.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr_capi.h"
#include "gsr_common.cuh"

namespace gsr {

// One thread per depth-sorted Gaussian: re-run the opacity-aware tile test and
// write (tile<<32 | idx) for every surviving tile, row-major like the reference
// (rasterizer_impl.cu:177-200).
__global__ void __launch_bounds__(256)
emit_kernel(int P, const uint64_t* __restrict__ sorted, const uint32_t* __restrict__ tiles_touched,
            const uint32_t* __restrict__ dup_offsets, const float2* __restrict__ means2D,
            const float4* __restrict__ conic_opacity, const int* __restrict__ radii, int grid_x,
            int grid_y, uint64_t* __restrict__ items, uint32_t* __restrict__ large_count,
            uint32_t* __restrict__ large_list, uint32_t large_capacity,
            const uint32_t* __restrict__ num_valid) {
  // Only the first num_valid depth-sorted positions own tiles.  In depth order the near
  // (= large) splats are neighbours, so lane l of a warp takes its Gaussian from the l-th of
  // 32 depth strata: every warp then walks a similar mix of near and far rectangles.
  const int lane = threadIdx.x & 31;
  const uint32_t nv = *num_valid;
  const uint32_t per_lane = (nv + 31u) / 32u;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp_global >= per_lane) return;  // whole warp
  const uint32_t jj = (uint32_t)lane * per_lane + warp_global;
  const int j = jj < nv ? (int)jj : P;
  uint32_t g = 0, cnt = 0, off = 0;
  float2 xy = {0.f, 0.f};
  float4 co = {0.f, 0.f, 0.f, 0.f};
  uint2 rect_min = {0, 0}, rect_max = {0, 0};
  float thr = 0.f;
  if (j < P) {
    g = (uint32_t)(sorted[j] & 0xffffffffu);
    cnt = tiles_touched[g];
  }
  if (cnt != 0) {
    off = dup_offsets[j];
    xy = means2D[g];
    co = conic_opacity[g];
    tile_rect(xy, radii[g], grid_x, grid_y, rect_min, rect_max);
    thr = opacity_power_threshold(co.w);
  }
  const int area = (int)((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y));
  // Very large rectangles are not walked here: they are cut into chunks of kQueueRectTiles
  // tiles, queued as (chunk << 24 | sorted position) and spread over the whole GPU by
  // emit_large_kernel (a single warp walking a 3000-tile rectangle would be the critical path).
  bool large = cnt != 0 && area > kQueueRectTiles;
  if (large) {
    const uint32_t nchunks = ((uint32_t)area + kQueueRectTiles - 1) / kQueueRectTiles;
    const uint32_t q0 = atomicAdd(large_count, nchunks);
    if (q0 + nchunks <= large_capacity && nchunks <= 256u) {
      for (uint32_t c = 0; c < nchunks; ++c) large_list[q0 + c] = (c << 24) | (uint32_t)j;
    } else {
      // queue full (pathological input): poison the reserved entries and walk it right here
      for (uint32_t c = 0; c < nchunks && q0 + c < large_capacity; ++c) large_list[q0 + c] = 0xffffffffu;
      large = false;
    }
  }
  // All (Gaussian, tile) pairs of the warp are tested 32 per step; a hit claims the next slot
  // of its owner's reserved range [off, off + cnt).  (The order inside one Gaussian's range is
  // irrelevant: its tiles are distinct, and the tile sort only needs the ranges of different
  // Gaussians to stay in depth order.)
  __shared__ uint32_t s_slot[256 / 32][32];
  const int warp = threadIdx.x >> 5;
  s_slot[warp][lane] = 0;
  __syncwarp();
  uint32_t* slot = s_slot[warp];
  warp_walk_rects_flat(cnt != 0 && !large, co, xy, rect_min, rect_max, thr, off, g,
                       [&](int owner, int x, int y, uint32_t owner_off, uint32_t owner_g) {
                         const uint32_t k = atomicAdd(&slot[owner], 1u);
                         const uint64_t tile = (uint64_t)(y * grid_x + x);
                         items[owner_off + k] = (tile << 32) | owner_g;
                       });
}

// One warp per queued chunk (kQueueRectTiles tiles of one large rectangle), grid-stride over
// the queue.  32 tiles are tested per step; the hits claim slots of the owner's reserved range
// by advancing dup_offsets[j] atomically (that array is not needed after the emit stage, and
// the order inside one Gaussian's range is irrelevant).
__global__ void __launch_bounds__(256)
emit_large_kernel(const uint32_t* __restrict__ large_count, const uint32_t* __restrict__ large_list,
                  const uint64_t* __restrict__ sorted, uint32_t* __restrict__ dup_offsets,
                  const float2* __restrict__ means2D, const float4* __restrict__ conic_opacity,
                  const int* __restrict__ radii, int grid_x, int grid_y,
                  uint64_t* __restrict__ items, uint32_t large_capacity) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t num_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t count = min(*large_count, large_capacity);
  const unsigned lanemask_lt = (1u << lane) - 1u;
  for (uint32_t k = warp_global; k < count; k += num_warps) {
    const uint32_t item = large_list[k];
    if (item == 0xffffffffu) continue;
    const uint32_t j = item & ((1u << 24) - 1u);
    const uint32_t chunk = item >> 24;
    const uint32_t g = (uint32_t)(sorted[j] & 0xffffffffu);
    const float2 xy = means2D[g];
    const float4 co = conic_opacity[g];
    uint2 rect_min, rect_max;
    tile_rect(xy, radii[g], grid_x, grid_y, rect_min, rect_max);
    const float thr = opacity_power_threshold(co.w);
    const int w = (int)(rect_max.x - rect_min.x), h = (int)(rect_max.y - rect_min.y);
    const int area = w * h;
    const int t_begin = (int)chunk * kQueueRectTiles;
    const int t_end = min(area, t_begin + kQueueRectTiles);
    for (int t0 = t_begin; t0 < t_end; t0 += 32) {
      const int t = t0 + lane;
      bool pass = false;
      int x = 0, y = 0;
      if (t < t_end) {
        y = (int)rect_min.y + t / w;
        x = (int)rect_min.x + t - (t / w) * w;
        pass = tile_min_power(co, xy, x, y) <= thr;
      }
      const unsigned pm = __ballot_sync(0xffffffffu, pass);
      if (pm == 0) continue;
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&dup_offsets[j], (uint32_t)__popc(pm));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (pass) {
        const uint64_t tile = (uint64_t)(y * grid_x + x);
        items[base + __popc(pm & lanemask_lt)] = (tile << 32) | g;
      }
    }
  }
}

int launch_emit(int P, const uint64_t* sorted, const uint32_t* tiles_touched,
                uint32_t* dup_offsets, const float2* means2D, const float4* conic_opacity,
                const int* radii, int grid_x, int grid_y, uint64_t* items, uint32_t* large_count,
                uint32_t* large_list, const uint32_t* num_valid, cudaStream_t s) {
  if (P <= 0) return 0;
  if (P >= (1 << 24)) return GSR_ERR_BAD_ARG;  // queue entries hold 24-bit sorted positions
  const uint32_t large_capacity = 2u * (uint32_t)P;  // the list lives in a P x u64 scratch array
  emit_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, sorted, tiles_touched, dup_offsets, means2D,
                                              conic_opacity, radii, grid_x, grid_y, items,
                                              large_count, large_list, large_capacity, num_valid);
  GSR_LAUNCHED();
  // 148 SMs x 2 CTAs x 8 warps walk the queued large rectangles
  emit_large_kernel<<<148 * 2, 256, 0, s>>>(large_count, large_list, sorted, dup_offsets, means2D,
                                            conic_opacity, radii, grid_x, grid_y, items,
                                            large_capacity);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// One thread per sorted duplicate: tile boundaries -> ranges (the job of the
// reference's identifyTileRanges, rasterizer_impl.cu:217-246) and the packed
// per-duplicate record {xy, conic, opacity, reject threshold, idx, rgb}.
__global__ void __launch_bounds__(256)
finalize_kernel(uint32_t R, const uint64_t* __restrict__ sorted, const float2* __restrict__ means2D,
                const float4* __restrict__ conic_opacity, const float* __restrict__ colors,
                uint2* __restrict__ ranges, float4* __restrict__ rec_a, float4* __restrict__ rec_b,
                float4* __restrict__ rec_c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const uint64_t it = sorted[i];
  const uint32_t tile = (uint32_t)(it >> 32);
  const uint32_t g = (uint32_t)(it & 0xffffffffu);
  if (i == 0) {
    ranges[tile].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(sorted[i - 1] >> 32);
    if (prev != tile) {
      ranges[prev].y = i;
      ranges[tile].x = i;
    }
  }
  if (i == R - 1) ranges[tile].y = R;

  const float2 xy = means2D[g];
  const float4 co = conic_opacity[g];
  // power below which opacity*exp(power) is certainly < 1/255 (1 % safety margin):
  // lets the blend kernels skip the exp for far-away pixels without changing any result.
  const float reject = -logf(255.0f * co.w) - 0.01f;
  rec_a[i] = make_float4(xy.x, xy.y, co.x, co.y);
  rec_b[i] = make_float4(co.z, co.w, reject, __uint_as_float(g));
  rec_c[i] = make_float4(colors[3 * (size_t)g], colors[3 * (size_t)g + 1], colors[3 * (size_t)g + 2], 0.f);
}

int launch_finalize(uint32_t R, const uint64_t* sorted, const float2* means2D,
                    const float4* conic_opacity, const float* colors, uint2* ranges, uint32_t T,
                    float4* rec_a, float4* rec_b, float4* rec_c, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(ranges, 0, (size_t)T * sizeof(uint2), s);
  if (e != cudaSuccess) return (int)e;
  if (R == 0) return 0;
  finalize_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, sorted, means2D, conic_opacity, colors, ranges,
                                                  rec_a, rec_b, rec_c);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// Longest-processing-time-first launch order for the blend kernels: a counting sort of the
// tiles by list length (1024 buckets of 4 entries, longest first).  The blend kernels run
// one CTA per tile for the whole length of its list, so starting the long lists first keeps
// the last wave short (matters most for heavy-tailed scenes such as cfg4's ground splats).
__global__ void __launch_bounds__(1024)
tile_order_kernel(const uint2* __restrict__ ranges, uint32_t T, uint32_t* __restrict__ order) {
  __shared__ uint32_t s_cnt[1024];
  __shared__ uint32_t s_warp[32];
  const int tid = threadIdx.x;
  s_cnt[tid] = 0;
  __syncthreads();
  for (uint32_t t = tid; t < T; t += 1024) {
    const uint2 r = ranges[t];
    const uint32_t b = 1023u - min(1023u, (r.y - r.x) >> 2);  // bucket 0 = longest lists
    atomicAdd(&s_cnt[b], 1u);
  }
  __syncthreads();
  // exclusive scan of the 1024 bucket counts
  const uint32_t c = s_cnt[tid];
  uint32_t inc = c;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += x;
    }
    s_warp[lane] = w;
  }
  __syncthreads();
  s_cnt[tid] = inc - c + (warp > 0 ? s_warp[warp - 1] : 0u);
  __syncthreads();
  for (uint32_t t = tid; t < T; t += 1024) {
    const uint2 r = ranges[t];
    const uint32_t b = 1023u - min(1023u, (r.y - r.x) >> 2);
    order[atomicAdd(&s_cnt[b], 1u)] = t;
  }
}

int launch_tile_order(const uint2* ranges, uint32_t T, uint32_t* order, cudaStream_t s) {
  tile_order_kernel<<<1, 1024, 0, s>>>(ranges, T, order);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// Debug/test export: rebuild the reference's 64-bit sorted keys (tile<<32 | depth bits)
// and the sorted Gaussian-id list from the packed records.
__global__ void export_binning_kernel(uint32_t R, const uint64_t* __restrict__ sorted,
                                      const float* __restrict__ depths, uint64_t* keys,
                                      uint32_t* point_list) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const uint64_t it = sorted[i];
  const uint32_t g = (uint32_t)(it & 0xffffffffu);
  if (keys) keys[i] = (it & 0xffffffff00000000ull) | (uint64_t)__float_as_uint(depths[g]);
  if (point_list) point_list[i] = g;
}

int launch_export_binning(uint32_t R, const uint64_t* sorted, const float* depths, uint64_t* keys,
                          uint32_t* point_list, cudaStream_t s) {
  if (R == 0) return 0;
  export_binning_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, sorted, depths, keys, point_list);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
