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
    // test all tiles of the chunk first (bit `st` = this lane's tile of step `st` passed), reserve
    // the chunk's slots with ONE atomic (a returning atomic per 32 tiles was the critical path),
    // then write
    static_assert(kQueueRectTiles <= 32 * 32, "one pass bit per step");
    uint32_t passbits = 0, total = 0;
    int st = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += 32, ++st) {
      const int t = t0 + lane;
      bool pass = false;
      if (t < t_end) {
        const int y = (int)rect_min.y + t / w;
        const int x = (int)rect_min.x + t - (t / w) * w;
        pass = tile_min_power(co, xy, x, y) <= thr;
      }
      passbits |= pass ? (1u << st) : 0u;
      total += __popc(__ballot_sync(0xffffffffu, pass));
    }
    if (total == 0) continue;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&dup_offsets[j], total);
    base = __shfl_sync(0xffffffffu, base, 0);
    st = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += 32, ++st) {
      const bool pass = (passbits >> st) & 1u;
      const unsigned pm = __ballot_sync(0xffffffffu, pass);
      if (pass) {
        const int t = t0 + lane;
        const int y = (int)rect_min.y + t / w;
        const int x = (int)rect_min.x + t - (t / w) * w;
        const uint64_t tile = (uint64_t)(y * grid_x + x);
        items[base + __popc(pm & lanemask_lt)] = (tile << 32) | g;
      }
      base += __popc(pm);
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

// One thread per sorted duplicate: the packed per-duplicate record {xy, conic, opacity, reject
// threshold, idx, rgb}.  (The tile ranges — the job of the reference's identifyTileRanges,
// rasterizer_impl.cu:217-246 — already exist: tile_meta_kernel scans the exact per-tile counts.)
__global__ void __launch_bounds__(256)
finalize_kernel(uint32_t R, const uint64_t* __restrict__ sorted, const float2* __restrict__ means2D,
                const float4* __restrict__ conic_opacity, const float* __restrict__ colors,
                float4* __restrict__ rec_a, float4* __restrict__ rec_b, float4* __restrict__ rec_c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const uint64_t it = sorted[i];
  const uint32_t g = (uint32_t)(it & 0xffffffffu);

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
                    const float4* conic_opacity, const float* colors, float4* rec_a, float4* rec_b,
                    float4* rec_c, cudaStream_t s) {
  if (R == 0) return 0;
  finalize_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, sorted, means2D, conic_opacity, colors, rec_a,
                                                  rec_b, rec_c);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// Everything that follows from the exact per-tile splat counts (accumulated by preprocess), in
// one single-CTA kernel right after preprocess:
//   * ranges[t]   = exclusive scan of the counts (the reference's identifyTileRanges result),
//   * *num_dup    = R, the total number of duplicates (read back by the host to size the
//                   binning buffer while the depth sort is already running),
//   * tile_ghist  = digit histograms of the tile sort (no histogram pass over the R duplicates),
//   * order       = longest-processing-time-first launch order of the blend kernels: a counting
//                   sort of the tiles by list length (1024 buckets of 4 entries, longest first) —
//                   the blend kernels run one CTA per half tile for the whole length of its
//                   list, so starting the long lists first keeps the last wave short,
//   * max_contrib = 0 (blend_fwd raises it with atomicMax).
__global__ void __launch_bounds__(1024)
tile_meta_kernel(const uint32_t* __restrict__ tile_count, uint32_t T, int npass,
                 uint2* __restrict__ ranges, uint32_t* __restrict__ num_dup,
                 uint32_t* __restrict__ tile_ghist, uint32_t* __restrict__ order,
                 uint32_t* __restrict__ max_contrib) {
  __shared__ uint32_t s_cnt[1024];
  __shared__ uint32_t s_hist[kTileHistPasses][256];
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  s_cnt[tid] = 0;
  (&s_hist[0][0])[tid] = 0;  // kTileHistPasses * 256 == 1024 entries
  if (tid == 0) s_carry = 0;
  __syncthreads();

  // block-wide inclusive scan of one value per thread (1024 threads); all threads must call
  auto block_incl_scan = [&](uint32_t v) -> uint32_t {
    uint32_t inc = v;
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
    const uint32_t r = inc + (warp > 0 ? s_warp[warp - 1] : 0u);
    __syncthreads();  // s_warp is reused by the next call
    return r;
  };

  for (uint32_t t0 = 0; t0 < T; t0 += 1024) {
    const uint32_t t = t0 + tid;
    const uint32_t c = t < T ? tile_count[t] : 0u;
    const uint32_t inc = block_incl_scan(c);
    const uint32_t carry = s_carry;
    if (t < T) {
      ranges[t] = make_uint2(carry + inc - c, carry + inc);
      max_contrib[t] = 0;
      if (c) {
        for (int p = 0; p < npass; ++p) atomicAdd(&s_hist[p][(t >> (8 * p)) & 0xffu], c);
      }
      atomicAdd(&s_cnt[1023u - min(1023u, c >> 2)], 1u);  // bucket 0 = longest lists
    }
    __syncthreads();
    if (tid == 1023) s_carry = carry + inc;
    __syncthreads();
  }
  if (tid == 0) *num_dup = s_carry;
  for (int i = tid; i < npass * 256; i += 1024) tile_ghist[i] = (&s_hist[0][0])[i];

  // exclusive scan of the 1024 bucket counts, then place the tiles
  const uint32_t c = s_cnt[tid];
  const uint32_t inc = block_incl_scan(c);
  s_cnt[tid] = inc - c;
  __syncthreads();
  for (uint32_t t = tid; t < T; t += 1024) {
    const uint32_t b = 1023u - min(1023u, tile_count[t] >> 2);
    order[atomicAdd(&s_cnt[b], 1u)] = t;
  }
}

int launch_tile_meta(const uint32_t* tile_count, uint32_t T, int tile_bits, uint2* ranges,
                     uint32_t* num_dup, uint32_t* tile_ghist, uint32_t* order, uint32_t* max_contrib,
                     cudaStream_t s) {
  const int npass = (tile_bits + 7) / 8;
  if (npass > kTileHistPasses) return GSR_ERR_BAD_ARG;
  tile_meta_kernel<<<1, 1024, 0, s>>>(tile_count, T, npass, ranges, num_dup, tile_ghist, order, max_contrib);
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
