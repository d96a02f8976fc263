// Binning: turn the rendered Gaussians into per-tile, depth-ordered lists of packed splat
// records.
//
// The reference builds one 64-bit key (tile<<32 | depth bits) per (Gaussian, tile) overlap,
// radix-sorts all R of them on 32+log2(T) bits with cub and then finds the tile boundaries
// (duplicateWithKeys + SortPairs + identifyTileRanges, rasterizer_impl.cu:139-246, 476-500):
// six passes over R keys and values.  That global sort is gone here:
//   * preprocess already counted, exactly, how many splats every tile receives, so
//     tile_meta_kernel gets the tile RANGES (and R) from a scan of T counters before a single
//     duplicate exists;
//   * emit_kernel drops every duplicate straight into its tile's range (slot order inside a
//     tile is arbitrary at this point: one atomic per duplicate);
//   * tile_sort_pack_*_kernel sort each tile's list by (depth bits, Gaussian index) in shared
//     memory — the 64-bit items are unique, so the result is THE order a stable sort of the
//     reference's keys produces, ties (equal depth -> ascending index) included — and write
//     the packed records the blend kernels stream with bulk copies.
// Per duplicate the pipeline moves 8 B (emit) + 16 B (sort in/out) + 36 B gathered + 48 B
// record, instead of 152 B of sort traffic plus the same gathers.
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "gsr_capi.h"
#include "gsr_common.cuh"
#include "gsr_params.cuh"

namespace gsr {

// ---------------------------------------------------------------------------------------------
// tile_meta: everything that follows from the exact per-tile splat counts, in one single-CTA
// kernel right after preprocess:
//   * ranges[t]      = exclusive scan of the counts (the reference's identifyTileRanges result),
//   * hdr->num_dup   = R, hdr->max_tile_len, hdr->num_long_tiles (lists longer than the small
//                      shared-memory sort handles),
//   * order          = longest-processing-time-first launch order for the per-tile kernels: a
//                      counting sort of the tiles by list length (1024 buckets of 4 entries,
//                      longest first) — one CTA works on a tile for the whole length of its
//                      list, so starting the long lists first keeps the last wave short,
//   * max_contrib    = 0 (blend_fwd raises it with atomicMax).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
tile_meta_kernel(const uint32_t* __restrict__ tile_count, uint32_t T, uint2* __restrict__ ranges,
                 GeomHeader* __restrict__ hdr, uint32_t* __restrict__ order,
                 uint32_t* __restrict__ max_contrib) {
  __shared__ uint32_t s_cnt[1024];
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_max, s_long, s_huge;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  s_cnt[tid] = 0;
  if (tid == 0) { s_max = 0; s_long = 0; s_huge = 0; }
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
  auto bucket_of = [](uint32_t c) { return 1023u - min(1023u, c >> 2); };  // bucket 0 = longest lists

  // each thread owns `ipt` consecutive tiles: one block scan for any T
  const uint32_t ipt = (T + 1023u) / 1024u;
  const uint32_t t_first = (uint32_t)tid * ipt;
  uint32_t sum = 0, mx = 0, nlong = 0, nhuge = 0;
  for (uint32_t k = 0; k < ipt; ++k) {
    const uint32_t t = t_first + k;
    if (t < T) {
      const uint32_t c = tile_count[t];
      sum += c;
      mx = max(mx, c);
      nlong += c >= (uint32_t)kSortSmallCap ? 1u : 0u;
      nhuge += c >= (uint32_t)kSortMidCap ? 1u : 0u;
    }
  }
  const uint32_t inc = block_incl_scan(sum);
  mx = __reduce_max_sync(0xffffffffu, mx);
  nlong = __reduce_add_sync(0xffffffffu, nlong);
  nhuge = __reduce_add_sync(0xffffffffu, nhuge);
  if (lane == 0) {
    if (mx) atomicMax(&s_max, mx);
    if (nlong) atomicAdd(&s_long, nlong);
    if (nhuge) atomicAdd(&s_huge, nhuge);
  }
  uint32_t run = inc - sum;
  for (uint32_t k = 0; k < ipt; ++k) {  // uniform trip count: the warp votes below need all lanes
    const uint32_t t = t_first + k;
    const bool in = t < T;
    const uint32_t c = in ? tile_count[t] : 0u;
    if (in) {
      // empty tiles keep (0, 0) like the reference's zero-initialised ranges array
      ranges[t] = c ? make_uint2(run, run + c) : make_uint2(0u, 0u);
      max_contrib[t] = 0;
    }
    run += c;
    // warp-aggregated: neighbouring tiles have similar lengths and would collide on a bucket
    const uint32_t b = in ? bucket_of(c) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    if (in && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&s_cnt[b], (uint32_t)__popc(peers));
  }
  __syncthreads();
  if (tid == 1023) {
    hdr->num_dup = inc;
    hdr->max_tile_len = s_max;
    hdr->num_long_tiles = s_long;
    hdr->num_huge_tiles = s_huge;
    hdr->num_large = 0;   // the preprocess stage's work queue is consumed; the emit stage builds its own
  }

  // exclusive scan of the 1024 bucket counts, then place the tiles
  const uint32_t c = s_cnt[tid];
  const uint32_t binc = block_incl_scan(c);
  s_cnt[tid] = binc - c;
  __syncthreads();
  const uint32_t t_end = (T + 31u) & ~31u;
  for (uint32_t t = tid; t < t_end; t += 1024) {
    const bool in = t < T;
    const uint32_t b = in ? bucket_of(tile_count[t]) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (in && lane == leader) base = atomicAdd(&s_cnt[b], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (in) order[base + __popc(peers & ((1u << lane) - 1u))] = t;
  }
}

int launch_tile_meta(const uint32_t* tile_count, uint32_t T, uint2* ranges, GeomHeader* hdr,
                     uint32_t* order, uint32_t* max_contrib, cudaStream_t s) {
  tile_meta_kernel<<<1, 1024, 0, s>>>(tile_count, T, ranges, hdr, order, max_contrib);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// emit: one thread per rendered Gaussian (compacted list): re-run the opacity-aware tile test
// and drop (depth_bits<<32 | idx) into the next free slot of every surviving tile's range.
// `capacity` guards the optimistically sized buffer (gsr_forward re-runs the stage if R
// exceeded it).
// ---------------------------------------------------------------------------------------------
#ifndef GSR_EMIT_OWNERS
#define GSR_EMIT_OWNERS 8
#endif
constexpr int kEmitOwners = GSR_EMIT_OWNERS;

__global__ void __launch_bounds__(256)
emit_kernel(const uint64_t* __restrict__ vis_items, const uint32_t* __restrict__ num_visible,
            const uint32_t* __restrict__ tiles_touched, const float2* __restrict__ means2D,
            const float4* __restrict__ conic_opacity, const int* __restrict__ radii, int grid_x,
            int grid_y, const uint2* __restrict__ ranges, uint32_t* __restrict__ tile_fill,
            uint64_t* __restrict__ items, uint32_t capacity, uint32_t* __restrict__ large_count,
            uint32_t* __restrict__ large_list, uint32_t large_capacity) {
  // kEmitOwners Gaussians per warp (lanes 0 .. kEmitOwners-1 own one each; all 32 lanes walk
  // their tiles): the stage is bound by the latency of the slot atomics, so the more warps share
  // the work the better.  One chunk per warp, a grid over all P list slots: warps striding over the list from a
  // machine-filling grid saved 7 us of empty CTAs at 3 M Gaussians (1/6 rendered) but lost 42 us at cfg4
  // (2 M rendered: the hardware's CTA scheduler balances uneven chunks, a static stride does not).
  const int lane = threadIdx.x & 31;
  const uint32_t nv = *num_visible;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp_global * kEmitOwners >= nv) return;  // whole warp
  const uint32_t j = lane < kEmitOwners ? warp_global * kEmitOwners + lane : 0xffffffffu;
  uint32_t g = 0, depth_bits = 0, cnt = 0;
  float2 xy = {0.f, 0.f};
  float4 co = {0.f, 0.f, 0.f, 0.f};
  uint2 rect_min = {0, 0}, rect_max = {0, 0};
  float thr = 0.f;
  if (j < nv) {
    const uint64_t it = vis_items[j];
    g = (uint32_t)(it & 0xffffffffu);
    depth_bits = (uint32_t)(it >> 32);
    cnt = tiles_touched[g];
  }
  if (cnt != 0) {
    xy = means2D[g];
    co = conic_opacity[g];
    tile_rect(xy, radii[g], grid_x, grid_y, rect_min, rect_max);
    thr = opacity_power_threshold(co.w);
  }
  const int area = (int)((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y));
  // Very large rectangles are not walked here: they are cut into chunks of about kQueueRectTiles
  // tiles, queued as (chunk << 24 | list position) and spread over the whole GPU by
  // emit_large_kernel (a single warp walking a 3000-tile rectangle would be the critical path).
  bool large = cnt != 0 && area > kQueueRectTiles;
  if (large) {
    const uint32_t nchunks = queue_chunks((uint32_t)area);
    const uint32_t q0 = atomicAdd(large_count, nchunks);
    if (q0 + nchunks <= large_capacity) {
      for (uint32_t c = 0; c < nchunks; ++c) large_list[q0 + c] = (c << 24) | j;
    } else {
      // queue full (pathological input): poison the reserved entries and walk it right here
      for (uint32_t c = 0; c < nchunks && q0 + c < large_capacity; ++c) large_list[q0 + c] = 0xffffffffu;
      large = false;
    }
  }
  // all (Gaussian, tile) pairs of the warp are tested 32 per step (gsr_common.cuh)
  warp_walk_rects_flat(cnt != 0 && !large, co, xy, rect_min, rect_max, thr, depth_bits, g,
                       [&](int, int x, int y, uint32_t owner_depth, uint32_t owner_g) {
                         const uint32_t tile = (uint32_t)(y * grid_x + x);
                         const uint2 rg = ranges[tile];
                         const uint32_t pos = rg.x + atomicAdd(&tile_fill[tile], 1u);
                         // pos < rg.y always holds (preprocess counted with the same test); the guard keeps a
                         // mismatch from ever spilling into the neighbouring tile's list
                         if (pos < rg.y && pos < capacity) items[pos] = ((uint64_t)owner_depth << 32) | owner_g;
                       });
}

// One warp per queued chunk (kQueueRectTiles tiles of one large rectangle), grid-stride over
// the queue, 32 tiles per step.
__global__ void __launch_bounds__(256)
emit_large_kernel(const uint32_t* __restrict__ large_count, const uint32_t* __restrict__ large_list,
                  const uint64_t* __restrict__ vis_items, const float2* __restrict__ means2D,
                  const float4* __restrict__ conic_opacity, const int* __restrict__ radii, int grid_x,
                  int grid_y, const uint2* __restrict__ ranges, uint32_t* __restrict__ tile_fill,
                  uint64_t* __restrict__ items, uint32_t capacity, uint32_t large_capacity) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t num_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t count = min(*large_count, large_capacity);
  for (uint32_t k = warp_global; k < count; k += num_warps) {
    const uint32_t item = large_list[k];
    if (item == 0xffffffffu) continue;
    const uint32_t j = item & ((1u << 24) - 1u);
    const uint32_t chunk = item >> 24;
    const uint64_t it = vis_items[j];
    const uint32_t g = (uint32_t)(it & 0xffffffffu);
    const float2 xy = means2D[g];
    const float4 co = conic_opacity[g];
    uint2 rect_min, rect_max;
    tile_rect(xy, radii[g], grid_x, grid_y, rect_min, rect_max);
    const float thr = opacity_power_threshold(co.w);
    const int w = (int)(rect_max.x - rect_min.x), h = (int)(rect_max.y - rect_min.y);
    const int area = w * h;
    const int per_chunk = (area + (int)queue_chunks((uint32_t)area) - 1) / (int)queue_chunks((uint32_t)area);
    const int t_begin = (int)chunk * per_chunk;
    const int t_end = min(area, t_begin + per_chunk);
    // per-rectangle constants out of the tile loop: the two reciprocals of the tile test (same operations on the
    // same inputs as tile_min_power) and t / w by multiplication (gsr_common.cuh)
    const float rcp_x = __frcp_rn(15 * 15 * co.x), rcp_z = __frcp_rn(15 * 15 * co.z);
    const uint32_t w_magic = w > 1 ? 0xffffffffu / (uint32_t)w + 1u : 0u;
    for (int t0 = t_begin; t0 < t_end; t0 += 32) {
      const int t = t0 + lane;
      if (t < t_end) {
        const int q = w > 1 ? (int)__umulhi((uint32_t)t, w_magic) : t;
        const int y = (int)rect_min.y + q;
        const int x = (int)rect_min.x + t - q * w;
        if (tile_min_power_pre(co, xy, x, y, rcp_x, rcp_z) <= thr) {
          const uint32_t tile = (uint32_t)(y * grid_x + x);
          const uint2 rg = ranges[tile];
          const uint32_t pos = rg.x + atomicAdd(&tile_fill[tile], 1u);
          if (pos < rg.y && pos < capacity) items[pos] = it;
        }
      }
    }
  }
}

int launch_emit(int P, const uint64_t* vis_items, const uint32_t* num_visible,
                const uint32_t* tiles_touched, const float2* means2D, const float4* conic_opacity,
                const int* radii, int grid_x, int grid_y, const uint2* ranges, uint32_t* tile_fill,
                uint64_t* items, uint32_t capacity, uint32_t* large_count, uint32_t* large_list,
                cudaStream_t s) {
  if (P <= 0) return 0;
  if (P >= (1 << 24)) return GSR_ERR_BAD_ARG;  // queue entries hold 24-bit list positions (gsr_forward checks first)
  const uint32_t large_capacity = 2u * (uint32_t)P;
  // the grid covers all P list slots; warps past the *num_visible rendered ones leave at once
  const long long warps = ((long long)P + kEmitOwners - 1) / kEmitOwners;
  emit_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(vis_items, num_visible, tiles_touched, means2D,
                                              conic_opacity, radii, grid_x, grid_y, ranges, tile_fill,
                                              items, capacity, large_count, large_list, large_capacity);
  GSR_LAUNCHED();
  // 148 SMs x 4 CTAs x 8 warps walk the queued large rectangles
  emit_large_kernel<<<148 * 4, 256, 0, s>>>(large_count, large_list, vis_items, means2D, conic_opacity,
                                            radii, grid_x, grid_y, ranges, tile_fill, items, capacity,
                                            large_capacity);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// tile_sort_pack: one CTA per tile.  Bitonic sort of the tile's 64-bit items (unique, so the
// sorted order is unique) and the packed per-duplicate records
//   {x, y, conic.x, conic.y | conic.z, opacity, reject_power, idx | r, g, b, 0}.
// Lists shorter than kSortSmallCap entries are sorted in registers by the "small" launch (one
// 4-warp CTA per tile, LPT order, bitonic network).  The "long" launch (persistent CTAs over the
// hdr->num_long_tiles longest tiles, which form a prefix of the LPT order) runs a stable LSD
// radix sort per tile, in shared memory up to kSortLongSmemCap entries and through global
// scratch beyond — any list length works.
// ---------------------------------------------------------------------------------------------
// ---- lists shorter than kSortSmallCap (= 512): one 4-warp CTA per tile, sorted in registers -----
// Thread t holds the kE consecutive network positions t*kE .. t*kE+kE-1 (n_pad = threads*kE).
// Compare-exchange stages with distance j < kE stay inside a thread; kE <= j < 32*kE are
// exchanges with lane t ^ (j / kE) through shuffles; the stages with j >= 32*kE (another
// warp) go through shared memory.  Several warps per tile shorten the dependent instruction chain of
// a tile — the kernel is one short wave whose duration is the latency of its longest lists.
// (A CTA-wide shared-memory network was bound by the shared-memory pipe: ncu mio_throttle.)
#ifndef GSR_TSP_THREADS
#define GSR_TSP_THREADS 128
#endif
constexpr int kSortSmallThreads = GSR_TSP_THREADS;

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int lane_mask) {
  const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, lane_mask);
  const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), lane_mask);
  return ((uint64_t)hi << 32) | lo;
}

// One stage (merge level K, distance J) — template parameters, so every register index is a
// compile-time constant and the whole network is straight-line code.
template <int kE, int kT, int K, int J>
__device__ __forceinline__ void bitonic_stage(uint64_t (&v)[kE], int t, uint64_t* s_x) {
  constexpr int kN = kT * kE;
  if constexpr (J >= kE) {
    constexpr int tm = J / kE;  // partner thread = t ^ tm
    const bool lower = (t & tm) == 0;
    const bool up = (K >= kN) ? true : ((t & (K / kE)) == 0);  // bit K of the position is clear
    const bool take_min = lower == up;
    if constexpr (tm < 32) {
#pragma unroll
      for (int e = 0; e < kE; ++e) {
        const uint64_t o = shfl_xor_u64(v[e], tm);
        v[e] = take_min ? (v[e] < o ? v[e] : o) : (v[e] < o ? o : v[e]);
      }
    } else {
      // the other warp: element-major staging, consecutive threads touch consecutive words
#pragma unroll
      for (int e = 0; e < kE; ++e) s_x[e * kT + t] = v[e];
      __syncthreads();
#pragma unroll
      for (int e = 0; e < kE; ++e) {
        const uint64_t o = s_x[e * kT + (t ^ tm)];
        v[e] = take_min ? (v[e] < o ? v[e] : o) : (v[e] < o ? o : v[e]);
      }
      __syncthreads();
    }
  } else {
#pragma unroll
    for (int e = 0; e < kE; ++e) {
      if ((e & J) == 0) {
        // bit K of (t*kE + e): inside the thread when K < kE, else a thread bit
        const bool up = (K < kE) ? ((e & K) == 0) : ((K >= kN) ? true : ((t & (K / kE)) == 0));
        const uint64_t a = v[e], b = v[e | J];
        const bool swap = (a > b) == up;
        v[e] = swap ? b : a;
        v[e | J] = swap ? a : b;
      }
    }
  }
}
template <int kE, int kT, int K, int J>
struct BitonicStages {
  static __device__ __forceinline__ void run(uint64_t (&v)[kE], int t, uint64_t* s_x) {
    bitonic_stage<kE, kT, K, J>(v, t, s_x);
    if constexpr (J > 1) BitonicStages<kE, kT, K, J / 2>::run(v, t, s_x);
  }
};
template <int kE, int kT, int K>
struct BitonicLevels {
  static __device__ __forceinline__ void run(uint64_t (&v)[kE], int t, uint64_t* s_x) {
    BitonicStages<kE, kT, K, K / 2>::run(v, t, s_x);
    if constexpr (K < kT * kE) BitonicLevels<kE, kT, K * 2>::run(v, t, s_x);
  }
};

// s_x: kT * kE 64-bit words of shared memory
template <int kE, int kT>
__device__ __forceinline__ void cta_sort_pack_tile(uint32_t n, uint32_t start, uint64_t* __restrict__ items,
                                                   const float4* __restrict__ splat_rec,
                                                   float4* __restrict__ rec, uint64_t* s_x) {
  const int t = threadIdx.x;
  uint64_t v[kE];
#pragma unroll
  for (int e = 0; e < kE; ++e) {
    const uint32_t i = (uint32_t)t * kE + e;
    v[e] = i < n ? items[start + i] : ~0ull;
  }
  BitonicLevels<kE, kT, 2>::run(v, t, s_x);
  // Sorted position t*kE + e sits in v[e].  The sorted Gaussian ids go through shared memory
  // so that each record can be copied by THREE threads (one float4 each), consecutive threads
  // covering consecutive 16-byte pieces: the gathers are 48-byte contiguous reads and the record
  // stream leaves in full, coalesced lines.
  uint32_t* s_g = reinterpret_cast<uint32_t*>(s_x);  // (the exchanges are over)
#pragma unroll
  for (int e = 0; e < kE; ++e) {
    const uint32_t i = (uint32_t)t * kE + e;
    if (i < n) items[start + i] = v[e];
    s_g[i] = (uint32_t)v[e];
  }
  __syncthreads();
  const uint32_t pieces = 3u * n;
  constexpr int kInFlight = 8;  // gathers in flight per thread (the copy is pure latency otherwise)
  for (uint32_t q0 = 0; q0 < pieces; q0 += kT * kInFlight) {
    float4 r4[kInFlight];
#pragma unroll
    for (int u = 0; u < kInFlight; ++u) {
      const uint32_t q = q0 + (uint32_t)kT * u + t;
      if (q < pieces) {
        const uint32_t pos = q / 3u, piece = q - 3u * pos;  // sorted position and float4 index
        r4[u] = __ldg(splat_rec + 3 * (size_t)s_g[pos] + piece);
      }
    }
#pragma unroll
    for (int u = 0; u < kInFlight; ++u) {
      const uint32_t q = q0 + (uint32_t)kT * u + t;
      if (q < pieces) rec[3 * (size_t)start + q] = r4[u];
    }
  }
}

__global__ void __launch_bounds__(kSortSmallThreads)
tile_sort_pack_small_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, uint32_t T,
                            BinHeader* __restrict__ bin_hdr, uint32_t capacity, uint64_t* __restrict__ items,
                            uint64_t* __restrict__ scratch, const uint32_t* __restrict__ sorted_len,
                            const float4* __restrict__ splat_rec, float4* __restrict__ rec) {
  __shared__ uint64_t s_x[kSortSmallCap];  // 4 KB: cross-warp exchange, then the sorted ids
  if (blockIdx.x == 0 && threadIdx.x == 0) bin_hdr->capacity = capacity;  // read by gsr_export_binning
  const uint32_t tile = order[blockIdx.x];
  const uint2 range = ranges[tile];
  uint32_t n = range.y - range.x;
  if (n == 0 || range.y > capacity) return;
  uint64_t* src = items;
  if (n >= (uint32_t)kSortSmallCap) {
    // a longer list: its nearest sorted_len entries were selected into `scratch` (tile_select_kernel); they are
    // ordered and packed here, the list itself stays as emitted (the full sort, if ever needed, starts from it)
    if (sorted_len == nullptr) return;
    n = sorted_len[tile];
    if (n == 0 || n >= (uint32_t)kSortSmallCap) return;
    src = scratch;
  }
  constexpr int kT = kSortSmallThreads;
  static_assert(kT * 8 >= kSortSmallCap, "the largest instantiation must hold kSortSmallCap entries");
  if (n <= kT) cta_sort_pack_tile<1, kT>(n, range.x, src, splat_rec, rec, s_x);
  else if (n <= 2 * kT) cta_sort_pack_tile<2, kT>(n, range.x, src, splat_rec, rec, s_x);
  else if (n <= 4 * kT || kT * 4 >= kSortSmallCap) cta_sort_pack_tile<4, kT>(n, range.x, src, splat_rec, rec, s_x);
  else cta_sort_pack_tile<8, kT>(n, range.x, src, splat_rec, rec, s_x);
}

// ---- lists of kSortSmallCap or more entries: select the nearest < kSortSmallCap -----------------------------
// The blend kernels stop at the deepest contributor of a tile — at most a few hundred list positions in an opaque
// scene (measured: 1 - 20 % of lists of 800 - 13 000 entries) — so ordering and packing the whole list is wasted.
// Per tile: min / max of the depth words, a 256-bucket histogram of the (monotone) linear map onto that range, the
// longest run of whole buckets holding fewer than kSortSmallCap entries -> those entries are the list's m nearest
// (equal depths share a bucket, so the cut never splits a tie) and are copied to `scratch`; the small kernel sorts
// and packs them.  blend_fwd reads the prefix only; a tile whose pixels are not saturated at its end — or whose
// prefix would be too short to be worth it — is flagged, fully sorted by the mid / long kernels below and blended
// again (rare: 1 tile in 3000 at cfg5, none at cfg1 / cfg4).
constexpr int kSelThreadsSort = 512;
constexpr uint32_t kSelMinPrefix = 256;   // fewer selectable entries: take the full sort straight away
// redo_flag bits: 1 / 2 = the upper / lower half tile continues behind its prefix in the second blend pass
// (set by blend_fwd), 4 = the tile is in redo_list
__device__ __forceinline__ void flag_redo(uint32_t tile, uint32_t* redo_flag, uint32_t* redo_list, uint32_t* redo_count) {
  if (!(atomicOr(&redo_flag[tile], 4u) & 4u)) redo_list[atomicAdd(redo_count, 1u)] = tile;
}
__global__ void __launch_bounds__(kSelThreadsSort)
tile_select_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ order, const GeomHeader* __restrict__ hdr,
                   uint32_t T, uint32_t capacity, const uint64_t* __restrict__ items, uint64_t* __restrict__ scratch,
                   uint32_t* __restrict__ sorted_len, uint32_t* __restrict__ redo_flag, uint32_t* __restrict__ redo_list,
                   uint32_t* __restrict__ redo_count) {
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_red[2][kSelThreadsSort / 32];
  __shared__ uint32_t s_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // lists of kSortSmallCap or more entries are the first num_long_tiles entries of the LPT order
  const uint32_t limit = min(T, hdr->num_long_tiles);
  for (uint32_t k = blockIdx.x; k < limit; k += gridDim.x) {
    const uint32_t tile = order[k];
    const uint2 range = ranges[tile];
    const uint32_t n = range.y - range.x, start = range.x;
    if (n < (uint32_t)kSortSmallCap || range.y > capacity) continue;
    // depth range of the list
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (uint32_t i = tid; i < n; i += kSelThreadsSort) {
      const uint32_t d = (uint32_t)(items[start + i] >> 32);
      lo = min(lo, d);
      hi = max(hi, d);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; }
    if (tid < 256) s_hist[tid] = 0;
    if (tid == 0) s_slot = 0;
    __syncthreads();
    lo = s_red[0][lane & (kSelThreadsSort / 32 - 1)];
    hi = s_red[1][lane & (kSelThreadsSort / 32 - 1)];
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    // bucket = floor((d - lo) * 256 / (hi - lo + 1)) in float arithmetic: every step is monotone in d
    const float scale = 256.0f / ((float)(hi - lo) + 1.0f);
    auto bucket = [lo, scale](uint32_t d) { return min(255u, (uint32_t)((float)(d - lo) * scale)); };
    for (uint32_t i = tid; i < n; i += kSelThreadsSort) atomicAdd(&s_hist[bucket((uint32_t)(items[start + i] >> 32))], 1u);
    __syncthreads();
    // inclusive scan of the 256 buckets by the first 8 warps; the cut = the buckets whose cumulative count stays
    // below kSortSmallCap (cumulative counts are non-decreasing, so these form a prefix of the buckets)
    uint32_t incl = 0;
    if (tid < 256) {
      incl = s_hist[tid];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
      }
      if (lane == 31) s_red[0][warp] = incl;
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t base = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w)
        if (w < warp) base += s_red[0][w];
      incl += base;
      const bool in = incl < (uint32_t)kSortSmallCap;
      const unsigned bal = __ballot_sync(0xffffffffu, in);
      if (lane == 0) s_red[1][warp] = (uint32_t)__popc(bal);
      // the entries inside the cut = the largest cumulative count still below the cap
      const uint32_t mine = in ? incl : 0u;
      const uint32_t wmax = __reduce_max_sync(0xffffffffu, mine);
      if (lane == 0) s_red[0][8 + warp] = wmax;
    }
    __syncthreads();
    uint32_t cut = 0, m = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      cut += s_red[1][w];
      m = max(m, s_red[0][8 + w]);
    }
    if (m < min(kSelMinPrefix, n)) {
      if (tid == 0) {
        sorted_len[tile] = 0;
        flag_redo(tile, redo_flag, redo_list, redo_count);
      }
      __syncthreads();
      continue;
    }
    for (uint32_t i = tid; i < n; i += kSelThreadsSort) {
      const uint64_t it = items[start + i];
      if (bucket((uint32_t)(it >> 32)) < cut) scratch[start + atomicAdd(&s_slot, 1u)] = it;
    }
    if (tid == 0) sorted_len[tile] = m;
    __syncthreads();   // shared state is reused by the next tile of this CTA
  }
}

// ---- lists of kSortSmallCap .. kSortMidCap-1 entries: the same network on a 512-thread CTA --------
// (2, 4 or 8 positions per thread; persistent CTAs over the slice of the LPT order that holds
// these lists)
constexpr int kSortMidThreads = 512;

__global__ void __launch_bounds__(kSortMidThreads)
tile_sort_pack_mid_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ order,
                          const GeomHeader* __restrict__ hdr, uint32_t T, uint32_t capacity,
                          uint64_t* __restrict__ items, const float4* __restrict__ splat_rec,
                          float4* __restrict__ rec, const uint32_t* __restrict__ redo_list,
                          const uint32_t* __restrict__ redo_count) {
  __shared__ uint64_t s_x[kSortMidThreads * 8];  // 32 KB
  // The FULL sort of a list: for the tiles of redo_list (flagged by tile_select_kernel / blend_fwd), or — without a
  // list (test export) — for every list of this length class.  In the LPT order the lists of >= kSortMidCap entries
  // come first, then those of kSortSmallCap .. kSortMidCap-1 entries (both caps are bucket boundaries of that order).
  const uint32_t first = redo_list ? 0u : min(T, hdr->num_huge_tiles);
  const uint32_t limit = redo_list ? min(T, *redo_count) : min(T, hdr->num_long_tiles);
  for (uint32_t k = first + blockIdx.x; k < limit; k += gridDim.x) {
    const uint2 range = ranges[redo_list ? redo_list[k] : order[k]];
    const uint32_t n = range.y - range.x;
    if (n < (uint32_t)kSortSmallCap || n >= (uint32_t)kSortMidCap || range.y > capacity) continue;
    if (n <= 2 * kSortMidThreads) cta_sort_pack_tile<2, kSortMidThreads>(n, range.x, items, splat_rec, rec, s_x);
    else if (n <= 4 * kSortMidThreads) cta_sort_pack_tile<4, kSortMidThreads>(n, range.x, items, splat_rec, rec, s_x);
    else cta_sort_pack_tile<8, kSortMidThreads>(n, range.x, items, splat_rec, rec, s_x);
    __syncthreads();  // s_x is reused by the next tile of this CTA
  }
}

// ---- lists of kSortMidCap or more entries: one CTA per tile, stable LSD radix sort -----------------
// 8-bit digits: the four digits of the depth bits, then a neighbour-compare fix-up of the (rare) runs of
// equal depth; the digits of the Gaussian index (low word, as many as P needs) are only sorted when a
// long run of equal depths exists (see the kernel).  Lists up to kSortLongSmemCap entries ping-pong
// between two shared-memory buffers, longer ones between the tile's own range and a scratch range
// in global memory (L2-resident) — any length works.  Per pass: per-warp digit histograms over
// contiguous segments (match_any ranking, no atomics), one scan, stable scatter.
constexpr int kSortLargeThreads = 512;
constexpr int kLongWarps = kSortLargeThreads / 32;

// lanes of the warp holding the same 8-bit digit (invalid lanes match nobody): eight ballots.
// (__match_any_sync does this in one instruction, but its latency grows with the number of
// distinct values and dominated this kernel: ncu r1, tile_sort_pack_long.)
__device__ __forceinline__ unsigned warp_peers_8bit(uint32_t d, bool valid) {
  unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned m = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? m : ~m;
  }
  return valid ? peers : 0u;
}

__device__ __forceinline__ void cta_radix_pass(const uint64_t* in, uint64_t* out, uint32_t n, int shift,
                                               uint32_t (*s_hist)[256], uint32_t* s_scan) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  constexpr int kBatch = 4;  // rounds whose loads are issued together
  // contiguous segment of this warp, a multiple of 32 items
  const uint32_t seg = ((n + kLongWarps * 32 - 1) / (kLongWarps * 32)) * 32;
  const uint32_t lo = min(n, (uint32_t)warp * seg), hi = min(n, lo + seg);
  for (int i = tid; i < kLongWarps * 256; i += kSortLargeThreads) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  // phase A: per-warp digit counts
  for (uint32_t i0 = lo; i0 < hi; i0 += 32 * kBatch) {
    uint32_t d[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const uint32_t i = i0 + 32 * u + lane;
      d[u] = i < hi ? (uint32_t)(in[i] >> shift) & 0xffu : 0u;
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (i0 + 32 * u >= hi) break;  // warp-uniform
      const bool valid = i0 + 32 * u + lane < hi;
      const unsigned peers = warp_peers_8bit(d[u], valid);
      if (valid && (peers & lt) == 0) s_hist[warp][d[u]] += (uint32_t)__popc(peers);
      __syncwarp();
    }
  }
  __syncthreads();
  // exclusive offsets: digit-major, then warp order inside a digit (keeps the pass stable)
  uint32_t total = 0;
  if (tid < 256) {
#pragma unroll
    for (int w = 0; w < kLongWarps; ++w) {
      const uint32_t c = s_hist[w][tid];
      s_hist[w][tid] = total;
      total += c;
    }
  }
  uint32_t inc = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) s_scan[warp] = inc;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int w = 0; w < kLongWarps; ++w)
    if (w < warp) wbase += s_scan[w];
  const uint32_t digit_base = wbase + inc - total;
  if (tid < 256) {
#pragma unroll
    for (int w = 0; w < kLongWarps; ++w) s_hist[w][tid] += digit_base;
  }
  __syncthreads();
  // phase B: stable scatter
  for (uint32_t i0 = lo; i0 < hi; i0 += 32 * kBatch) {
    uint64_t it[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const uint32_t i = i0 + 32 * u + lane;
      it[u] = i < hi ? in[i] : 0ull;
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (i0 + 32 * u >= hi) break;  // warp-uniform
      const bool valid = i0 + 32 * u + lane < hi;
      const uint32_t d = (uint32_t)(it[u] >> shift) & 0xffu;
      const unsigned peers = warp_peers_8bit(d, valid);
      uint32_t base = 0;
      if (valid) base = s_hist[warp][d];
      __syncwarp();
      if (valid && (peers & lt) == 0) s_hist[warp][d] = base + (uint32_t)__popc(peers);
      __syncwarp();
      if (valid) out[base + __popc(peers & lt)] = it[u];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSortLargeThreads)
tile_sort_pack_long_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ order,
                           const GeomHeader* __restrict__ hdr, uint32_t T, uint32_t capacity, int idx_passes,
                           uint64_t* __restrict__ items, uint64_t* __restrict__ scratch,
                           const float4* __restrict__ splat_rec, float4* __restrict__ rec,
                           const uint32_t* __restrict__ redo_list, const uint32_t* __restrict__ redo_count) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  uint32_t (*s_hist)[256] = reinterpret_cast<uint32_t (*)[256]>(s_raw);
  uint64_t* s_a = reinterpret_cast<uint64_t*>(s_raw + kLongWarps * 256 * sizeof(uint32_t));
  uint64_t* s_b = s_a + kSortLongSmemCap;
  __shared__ uint32_t s_scan[kLongWarps];
  const int tid = threadIdx.x;
  // The lists of kSortMidCap or more entries are exactly the first num_huge_tiles entries of
  // the LPT order (kSortMidCap is the lower edge of its last bucket).
  const uint32_t limit = redo_list ? min(T, *redo_count) : min(T, hdr->num_huge_tiles);
  for (uint32_t k = blockIdx.x; k < limit; k += gridDim.x) {
    const uint2 range = ranges[redo_list ? redo_list[k] : order[k]];
    const uint32_t n = range.y - range.x, start = range.x;
    if (n < (uint32_t)kSortMidCap || range.y > capacity) continue;
    const bool in_smem = n <= (uint32_t)kSortLongSmemCap;
    uint64_t* a = in_smem ? s_a : items + start;
    uint64_t* b = in_smem ? s_b : scratch + start;
    if (in_smem) {
      for (uint32_t i = tid; i < n; i += kSortLargeThreads) s_a[i] = items[start + i];
      __syncthreads();
    }
    // The order wanted is (depth bits, Gaussian index).  Equal depths are rare (a handful of pairs per
    // list), so the index digits are NOT sorted first: four passes over the depth bits, then every run
    // of equal depths — found by comparing neighbours — is put into index order by the thread that owns
    // its first element (insertion sort, runs of up to kTieRun entries).  Only when a longer run exists
    // (degenerate scenes: thousands of splats at one depth) does the list take the full LSD sequence,
    // index digits then depth digits, which is correct from any starting order.
    for (int p = 0; p < 4; ++p) {
      cta_radix_pass(a, b, n, 32 + 8 * p, s_hist, s_scan);
      uint64_t* t = a; a = b; b = t;
    }
    if (tid == 0) s_scan[0] = 0;
    __syncthreads();
    constexpr uint32_t kTieRun = 32;
    for (uint32_t i = tid; i + 1 < n; i += kSortLargeThreads) {
      const uint32_t d = (uint32_t)(a[i] >> 32);
      if ((uint32_t)(a[i + 1] >> 32) != d) continue;
      if (i > 0 && (uint32_t)(a[i - 1] >> 32) == d) continue;  // not the head of its run
      uint32_t len = 2;
      while (i + len < n && len <= kTieRun && (uint32_t)(a[i + len] >> 32) == d) ++len;
      if (len > kTieRun) {
        s_scan[0] = 1;  // benign race: every writer stores 1
        continue;
      }
      for (uint32_t j = 1; j < len; ++j) {
        const uint64_t x = a[i + j];
        uint32_t k = j;
        while (k > 0 && a[i + k - 1] > x) {
          a[i + k] = a[i + k - 1];
          --k;
        }
        a[i + k] = x;
      }
    }
    __syncthreads();
    const bool long_runs = s_scan[0] != 0;
    __syncthreads();  // s_scan is reused by the passes below
    if (long_runs) {
      for (int p = 0; p < idx_passes + 4; ++p) {
        const int shift = p < idx_passes ? 8 * p : 32 + 8 * (p - idx_passes);
        cta_radix_pass(a, b, n, shift, s_hist, s_scan);
        uint64_t* t = a; a = b; b = t;
      }
    }
    // sorted list in `a`
    if (a != items + start)
      for (uint32_t i = tid; i < n; i += kSortLargeThreads) items[start + i] = a[i];
    // three consecutive threads copy the three float4 of a record: coalesced on both sides
    const uint32_t pieces = 3u * n;
    constexpr int kInFlight = 4;
    for (uint32_t q0 = 0; q0 < pieces; q0 += kSortLargeThreads * kInFlight) {
      float4 t4[kInFlight];
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) {
        const uint32_t q = q0 + (uint32_t)kSortLargeThreads * u + tid;
        if (q < pieces) {
          const uint32_t pos = q / 3u, piece = q - 3u * pos;
          t4[u] = __ldg(splat_rec + 3 * (size_t)(uint32_t)(a[pos] & 0xffffffffu) + piece);
        }
      }
#pragma unroll
      for (int u = 0; u < kInFlight; ++u) {
        const uint32_t q = q0 + (uint32_t)kSortLargeThreads * u + tid;
        if (q < pieces) rec[3 * (size_t)start + q] = t4[u];
      }
    }
    __syncthreads();  // `a` may be a shared buffer the next tile overwrites
  }
}

// select (lists of kSortSmallCap+ entries, when expected) + sort / pack of the short lists and of the selected prefixes
int launch_tile_sort_pack(int P, uint32_t T, const uint2* ranges, const uint32_t* order, const GeomHeader* hdr,
                          BinHeader* bin_hdr, uint32_t capacity, uint64_t* items, uint64_t* scratch,
                          const float4* splat_rec, float4* rec, bool do_mid, bool do_long, uint32_t* sorted_len,
                          uint32_t* redo_flag, uint32_t* redo_list, uint32_t* redo_count, cudaStream_t s) {
  (void)P;
  if (T == 0) return 0;
  // The kernels for lists of 512+ entries are launched only when the caller expects such lists (gsr_forward: the
  // previous frames' counts; it re-runs the stage if the guess was wrong): an empty launch costs 3 us.
  const bool longer = do_mid || do_long;
  if (longer) {
    tile_select_kernel<<<148 * 2, kSelThreadsSort, 0, s>>>(ranges, order, hdr, T, capacity, items, scratch, sorted_len,
                                                           redo_flag, redo_list, redo_count);
    GSR_LAUNCHED();
  }
  tile_sort_pack_small_kernel<<<T, kSortSmallThreads, 0, s>>>(ranges, order, T, bin_hdr, capacity, items, scratch,
                                                                        longer ? sorted_len : nullptr, splat_rec, rec);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// The FULL sort + pack of lists of kSortSmallCap+ entries: of the tiles in redo_list (production: the few tiles whose
// sorted prefix did not suffice), or of every such list (redo_list == nullptr: the test export).
int launch_tile_sort_full(int P, uint32_t T, const uint2* ranges, const uint32_t* order, const GeomHeader* hdr,
                          uint32_t capacity, uint64_t* items, uint64_t* scratch, const float4* splat_rec, float4* rec,
                          bool do_mid, bool do_long, const uint32_t* redo_list, const uint32_t* redo_count,
                          cudaStream_t s) {
  if (T == 0) return 0;
  if (do_mid) {
    // persistent: 2 CTAs of 16 warps per SM (exits at once if there is nothing to do)
    tile_sort_pack_mid_kernel<<<148 * 2, kSortMidThreads, 0, s>>>(ranges, order, hdr, T, capacity, items, splat_rec,
                                                                 rec, redo_list, redo_count);
    GSR_LAUNCHED();
  }
  if (!do_long) return (int)cudaGetLastError();
  const int dyn = kLongWarps * 256 * (int)sizeof(uint32_t) + 2 * kSortLongSmemCap * (int)sizeof(uint64_t);
  {
    // the attribute belongs to the (function, device) pair: set it once per device
    constexpr int kMaxDev = 64;
    static std::atomic<bool> attr_set[kMaxDev];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDev || !attr_set[dev].load(std::memory_order_acquire)) {
      cudaError_t e = cudaFuncSetAttribute(tile_sort_pack_long_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
      if (e != cudaSuccess) return (int)e;
      if (dev >= 0 && dev < kMaxDev) attr_set[dev].store(true, std::memory_order_release);
    }
  }
  int idx_bits = 1;
  while (idx_bits < 32 && (1ull << idx_bits) < (unsigned long long)P) ++idx_bits;
  const int idx_passes = (idx_bits + 7) / 8;
  // two CTAs per SM (80 KB of shared memory each)
  tile_sort_pack_long_kernel<<<148 * 2, kSortLargeThreads, dyn, s>>>(ranges, order, hdr, T, capacity, idx_passes,
                                                                     items, scratch, splat_rec, rec, redo_list, redo_count);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// Debug/test export: rebuild the reference's 64-bit sorted keys (tile<<32 | depth bits)
// and the sorted Gaussian-id list from the per-tile lists.  One CTA per tile.
__global__ void export_binning_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ items,
                                      uint64_t* keys, uint32_t* point_list) {
  const uint32_t tile = blockIdx.x;
  const uint2 range = ranges[tile];
  for (uint32_t i = range.x + threadIdx.x; i < range.y; i += blockDim.x) {
    const uint64_t it = items[i];
    if (keys) keys[i] = ((uint64_t)tile << 32) | (it >> 32);
    if (point_list) point_list[i] = (uint32_t)(it & 0xffffffffu);
  }
}

int launch_export_binning(uint32_t T, const uint2* ranges, const uint64_t* items, uint64_t* keys,
                          uint32_t* point_list, cudaStream_t s) {
  if (T == 0) return 0;
  export_binning_kernel<<<T, 128, 0, s>>>(ranges, items, keys, point_list);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
