// Hand-written stable LSD "onesweep" radix sort of 64-bit items and a chained
// (decoupled look-back) exclusive scan.  Replaces the reference's use of
// cub::DeviceRadixSort::SortPairs (rasterizer_impl.cu:486-491, simple_knn.cu:210-213)
// and cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:456-459).
//
// Onesweep = one up-front histogram pass for ALL digit positions, then one
// read + one write per 8-bit digit: every thread block ranks its 4096-item tile
// in shared memory, publishes its per-digit counts, and resolves its global
// offsets by looking back at the counts of the tiles before it instead of
// waiting for a separate scan kernel.  Only the key bits [begin_bit, end_bit)
// the caller says are in use are sorted.
//
// HBM traffic per item: 8 B (histogram) + passes * 16 B.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr_common.cuh"

namespace gsr {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
// Items per thread: 16 (4096-item tiles) for large inputs; 4 (1024-item tiles) for small
// ones, where a pass is latency-bound and more, shorter CTAs fill the 148 SMs better.
constexpr int kIptLarge = 16;
constexpr int kIptSmall = 4;
constexpr int64_t kSmallSortLimit = 96 * 1024;
constexpr int kMaxPasses = 8;

constexpr uint32_t kFlagAgg = 1u << 30;   // tile published its own count
constexpr uint32_t kFlagInc = 2u << 30;   // tile published the inclusive prefix
constexpr uint32_t kValMask = (1u << 30) - 1;

// temp layout (uint32 words): ghist[kMaxPasses][256] | tile_counter[kMaxPasses] (padded to 64)
//                             | lookback[kMaxPasses][num_tiles][256]
static inline int sort_ipt(int64_t n) { return n < kSmallSortLimit ? kIptSmall : kIptLarge; }
static inline int64_t sort_num_tiles(int64_t n) {
  const int64_t tile = (int64_t)kSortThreads * sort_ipt(n);
  return (n + tile - 1) / tile;
}
static inline size_t sort_header_words() { return (size_t)kMaxPasses * kRadix + 64; }

size_t sort_temp_bytes(int64_t n) {
  if (n < 0) n = 0;
  size_t words = sort_header_words() + (size_t)kMaxPasses * (size_t)sort_num_tiles(n) * kRadix;
  return words * 4 + 256;
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  return *reinterpret_cast<const volatile uint32_t*>(p);
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
  *reinterpret_cast<volatile uint32_t*>(p) = v;
}

// Block-wide exclusive scan of one value per thread (256 threads).
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_warp /*[8]*/,
                                                        uint32_t* total_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t wbase = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) {
    uint32_t c = s_warp[w];
    if (w < warp) wbase += c;
    tot += c;
  }
  __syncthreads();  // s_warp may be reused
  if (total_out) *total_out = tot;
  return wbase + inc - v;
}

// ---------------------------------------------------------------------------
// Histograms of every digit position in one read of the items.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
sort_histogram_kernel(const uint64_t* __restrict__ items, uint32_t n, int begin_bit, int end_bit,
                      int npass, uint32_t* __restrict__ ghist) {
  __shared__ uint32_t s_hist[kMaxPasses][kRadix];
  for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) (&s_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t it = items[i];
    for (int p = 0; p < npass; ++p) {
      const int sh = begin_bit + p * kRadixBits;
      const int bits = min(kRadixBits, end_bit - sh);
      const uint32_t d = (uint32_t)(it >> sh) & ((1u << bits) - 1u);
      atomicAdd(&s_hist[p][d], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * kRadix; i += blockDim.x) {
    const uint32_t c = (&s_hist[0][0])[i];
    if (c) atomicAdd(&ghist[i], c);
  }
}

// ---------------------------------------------------------------------------
// One onesweep pass: stable scatter of the items by one 8-bit digit.
// ---------------------------------------------------------------------------
// kDropInvalid: items whose high word is 0xFFFFFFFF (preprocess: Gaussian not rendered) are
// neither ranked nor written, i.e. the pass doubles as an order-preserving compaction; `ghist`
// must then count the valid items only.  n_dev (optional) overrides n with a device-side
// count; CTAs whose tile starts beyond it leave at once (the grid is sized for the host bound).
template <int kItemsPerThread, bool kDropInvalid>
__global__ void __launch_bounds__(kSortThreads)
onesweep_pass_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t n,
                     const uint32_t* __restrict__ n_dev, int shift, uint32_t digit_mask,
                     const uint32_t* __restrict__ ghist, uint32_t* __restrict__ tile_counter,
                     uint32_t* __restrict__ lookback) {
  constexpr int kTileItems = kSortThreads * kItemsPerThread;
  __shared__ uint64_t s_items[kTileItems];              // 32 KB at 16 items per thread
  __shared__ uint32_t s_warp_hist[kSortWarps][kRadix];  // 8 KB
  __shared__ uint32_t s_bin_start[kRadix];
  __shared__ uint32_t s_digit_base[kRadix];
  __shared__ uint32_t s_scan[kSortWarps];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (n_dev != nullptr) n = *n_dev;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * (uint32_t)kTileItems;
  if (base >= n) return;  // whole CTA: nobody looks back at a tile past the end
  const uint32_t tile_count = min((uint32_t)kTileItems, n - base);
  const uint32_t lanemask_lt = (1u << lane) - 1u;

  // ---- load (warp-striped) and rank inside the warp -------------------------
  uint64_t items[kItemsPerThread];
  uint32_t ranks[kItemsPerThread];
  const uint32_t wbase = warp * (32 * kItemsPerThread);
#pragma unroll
  for (int k = 0; k < kItemsPerThread; ++k) {
    const uint32_t li = wbase + k * 32 + lane;
    items[k] = (li < tile_count) ? in[base + li] : ~0ull;
  }
#pragma unroll
  for (int k = 0; k < kItemsPerThread; ++k) {
    const uint32_t li = wbase + k * 32 + lane;
    bool valid = li < tile_count;
    if (kDropInvalid) valid = valid && (uint32_t)(items[k] >> 32) != 0xffffffffu;
    const uint32_t digit = (uint32_t)(items[k] >> shift) & digit_mask;
    const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : 0xffffu);
    const uint32_t before = __popc(peers & lanemask_lt);
    uint32_t prev = 0;
    if (valid) prev = s_warp_hist[warp][digit];
    __syncwarp();
    if (valid && before == 0) s_warp_hist[warp][digit] = prev + __popc(peers);
    __syncwarp();
    ranks[k] = prev + before;
  }
  __syncthreads();

  // ---- per-digit: scan across warps, publish, look back ---------------------
  // thread `tid` owns digit `tid`
  uint32_t total = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; ++w) {
    const uint32_t c = s_warp_hist[w][tid];
    s_warp_hist[w][tid] = total;
    total += c;
  }
  uint32_t* my_flag = lookback + (size_t)tile * kRadix + tid;
  st_volatile_u32(my_flag, (tile == 0 ? kFlagInc : kFlagAgg) | total);

  uint32_t kept = tile_count;  // items of this tile that are written (all of them unless kDropInvalid)
  const uint32_t bin_start = block_excl_scan_256(total, s_scan, kDropInvalid ? &kept : nullptr);
  const uint32_t gcount = ghist[tid];
  const uint32_t gbase = block_excl_scan_256(gcount, s_scan, nullptr);

  uint32_t excl = 0;
  if (tile > 0) {
    int t = (int)tile - 1;
    while (true) {
      const uint32_t v = ld_volatile_u32(lookback + (size_t)t * kRadix + tid);
      const uint32_t flag = v & ~kValMask;
      if (flag == 0) continue;  // predecessor has not published yet
      excl += v & kValMask;
      if (flag == kFlagInc) break;
      --t;
    }
    st_volatile_u32(my_flag, kFlagInc | (excl + total));
  }
  s_bin_start[tid] = bin_start;
  s_digit_base[tid] = gbase + excl - bin_start;
  __syncthreads();

  // ---- reorder inside the tile, then coalesced scatter ----------------------
#pragma unroll
  for (int k = 0; k < kItemsPerThread; ++k) {
    const uint32_t li = wbase + k * 32 + lane;
    bool valid = li < tile_count;
    if (kDropInvalid) valid = valid && (uint32_t)(items[k] >> 32) != 0xffffffffu;
    if (valid) {
      const uint32_t digit = (uint32_t)(items[k] >> shift) & digit_mask;
      s_items[s_bin_start[digit] + s_warp_hist[warp][digit] + ranks[k]] = items[k];
    }
  }
  __syncthreads();
#pragma unroll 4
  for (uint32_t i = tid; i < kept; i += kSortThreads) {
    const uint64_t it = s_items[i];
    const uint32_t digit = (uint32_t)(it >> shift) & digit_mask;
    out[s_digit_base[digit] + i] = it;
  }
}

template <bool kDrop>
static void launch_pass(int ipt, int64_t tiles, const uint64_t* src, uint64_t* dst, uint32_t n,
                        const uint32_t* n_dev, int sh, uint32_t mask, const uint32_t* ghist,
                        uint32_t* counter, uint32_t* lookback, cudaStream_t s) {
  if (ipt == kIptSmall)
    onesweep_pass_kernel<kIptSmall, kDrop><<<(unsigned)tiles, kSortThreads, 0, s>>>(
        src, dst, n, n_dev, sh, mask, ghist, counter, lookback);
  else
    onesweep_pass_kernel<kIptLarge, kDrop><<<(unsigned)tiles, kSortThreads, 0, s>>>(
        src, dst, n, n_dev, sh, mask, ghist, counter, lookback);
  GSR_LAUNCHED();
}

// Sorts on bits [begin_bit,end_bit).  `in` is not modified; the result lands in
// `out` (*result == out).  `in`, `out`, `alt` must be distinct when more than
// one pass is needed (alt may be NULL for a single pass).
int launch_sort_u64(const uint64_t* in, uint64_t* out, uint64_t* alt, int64_t n, int begin_bit,
                    int end_bit, char* tmp, cudaStream_t s, uint64_t** result) {
  if (result) *result = out;
  if (n <= 0) return 0;
  if (n >= (int64_t)kValMask) return -1;
  int npass = (end_bit - begin_bit + kRadixBits - 1) / kRadixBits;
  if (npass <= 0) {
    cudaError_t e = cudaMemcpyAsync(out, in, (size_t)n * 8, cudaMemcpyDeviceToDevice, s);
    return (int)e;
  }
  if (npass > kMaxPasses) return -1;
  const int64_t tiles = sort_num_tiles(n);
  uint32_t* words = reinterpret_cast<uint32_t*>(tmp);
  uint32_t* ghist = words;
  uint32_t* counters = words + (size_t)kMaxPasses * kRadix;
  uint32_t* lookback = words + sort_header_words();
  const size_t zero_words = sort_header_words() + (size_t)npass * tiles * kRadix;
  cudaError_t e = cudaMemsetAsync(tmp, 0, zero_words * 4, s);
  if (e != cudaSuccess) return (int)e;

  int64_t hb = (n + kSortThreads * 8 - 1) / (kSortThreads * 8);
  int hist_blocks = (int)(hb < 148 * 8 ? hb : 148 * 8);
  sort_histogram_kernel<<<hist_blocks, kSortThreads, 0, s>>>(in, (uint32_t)n, begin_bit, end_bit,
                                                             npass, ghist);
  GSR_LAUNCHED();
  const uint64_t* src = in;
  uint64_t* dst = (npass & 1) ? out : alt;
  for (int p = 0; p < npass; ++p) {
    const int sh = begin_bit + p * kRadixBits;
    const int bits = (end_bit - sh) < kRadixBits ? (end_bit - sh) : kRadixBits;
    launch_pass<false>(sort_ipt(n), tiles, src, dst, (uint32_t)n, nullptr, sh, (1u << bits) - 1u,
                       ghist + (size_t)p * kRadix, counters + p, lookback + (size_t)p * tiles * kRadix, s);
    src = dst;
    dst = (dst == out) ? alt : out;
  }
  return (int)cudaGetLastError();
}

// ---- the two sorts of the forward pass, with their histograms computed upstream ----------
// Depth sort: `tmp` (sort_depth_temp_bytes) starts with the 4 x 256 digit histograms and must
// be zeroed by the caller BEFORE the kernel that accumulates them (preprocess) runs.
uint32_t* sort_ghist(char* tmp) { return reinterpret_cast<uint32_t*>(tmp); }

// Depth order of the rendered Gaussians: items (depth_bits<<32 | idx) for all P Gaussians, the
// non-rendered ones marked with 0xFFFFFFFF in the high word.  Pass 1 reads all P items and
// keeps only the *n_valid rendered ones; passes 2-4 run on those (grid sized for P, surplus
// CTAs exit).  ghist (4 x 256, valid items only) was accumulated by preprocess.
// a -> b -> a -> b -> a: the sorted list ends in `a`.
int launch_sort_depth(uint64_t* a, uint64_t* b, int64_t P, const uint32_t* n_valid, char* tmp,
                      cudaStream_t s) {
  if (P <= 0) return 0;
  if (P >= (int64_t)kValMask) return -1;
  uint32_t* words = reinterpret_cast<uint32_t*>(tmp);
  uint32_t* ghist = words;
  uint32_t* counters = words + (size_t)kMaxPasses * kRadix;
  uint32_t* lookback = words + sort_header_words();
  const int64_t tiles_full = sort_num_tiles(P);
  // passes over the rendered subset: short tiles unless the scene is very large
  const int ipt_sub = P <= (int64_t)2 * 1024 * 1024 ? kIptSmall : kIptLarge;
  const int64_t tiles_sub = (P + (int64_t)kSortThreads * ipt_sub - 1) / ((int64_t)kSortThreads * ipt_sub);
  // the look-back area was sized (and zeroed) for 4 passes of sort_num_tiles(P) tiles of the
  // small kind when P < kSmallSortLimit, else of the large kind; tiles_sub may exceed that
  launch_pass<true>(sort_ipt(P), tiles_full, a, b, (uint32_t)P, nullptr, 32, 0xffu, ghist, counters, lookback, s);
  const uint64_t* src = b;
  uint64_t* dst = a;
  for (int p = 1; p < 4; ++p) {
    launch_pass<false>(ipt_sub, tiles_sub, src, dst, 0u, n_valid, 32 + 8 * p, 0xffu,
                       ghist + (size_t)p * kRadix, counters + p,
                       lookback + (size_t)tiles_full * kRadix + (size_t)(p - 1) * tiles_sub * kRadix, s);
    uint64_t* t = const_cast<uint64_t*>(src);
    src = dst;
    dst = t;
  }
  return (int)cudaGetLastError();
}
size_t sort_depth_temp_bytes(int64_t P) {
  if (P < 0) P = 0;
  const int ipt_sub = P <= (int64_t)2 * 1024 * 1024 ? kIptSmall : kIptLarge;
  const int64_t tiles_sub = (P + (int64_t)kSortThreads * ipt_sub - 1) / ((int64_t)kSortThreads * ipt_sub);
  return (sort_header_words() + (size_t)(sort_num_tiles(P) + 3 * tiles_sub) * kRadix) * 4 + 256;
}

// Tile order of the duplicates: only the tile digits (bits 32..32+tile_bits) are sorted; their
// histograms come from the per-tile counts (binning.cu: tile_meta_kernel).  Result in `out`.
int launch_sort_tile(const uint64_t* in, uint64_t* out, uint64_t* alt, int64_t R, int tile_bits,
                     const uint32_t* ghist, char* tmp, cudaStream_t s) {
  if (R <= 0) return 0;
  if (R >= (int64_t)kValMask) return -1;
  const int npass = (tile_bits + kRadixBits - 1) / kRadixBits;
  if (npass <= 0 || npass > kMaxPasses) return -1;
  uint32_t* words = reinterpret_cast<uint32_t*>(tmp);
  uint32_t* counters = words + (size_t)kMaxPasses * kRadix;
  uint32_t* lookback = words + sort_header_words();
  const int64_t tiles = sort_num_tiles(R);
  cudaError_t e = cudaMemsetAsync(counters, 0, (64 + (size_t)npass * tiles * kRadix) * 4, s);
  if (e != cudaSuccess) return (int)e;
  const uint64_t* src = in;
  uint64_t* dst = (npass & 1) ? out : alt;
  for (int p = 0; p < npass; ++p) {
    const int sh = 32 + p * kRadixBits;
    const int bits = (tile_bits - p * kRadixBits) < kRadixBits ? (tile_bits - p * kRadixBits) : kRadixBits;
    launch_pass<false>(sort_ipt(R), tiles, src, dst, (uint32_t)R, nullptr, sh, (1u << bits) - 1u,
                       ghist + (size_t)p * kRadix, counters + p, lookback + (size_t)p * tiles * kRadix, s);
    src = dst;
    dst = (dst == out) ? alt : out;
  }
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Chained exclusive scan of counts[order[j]] (j in depth order) -> offsets[j];
// the grand total is written to *total_out.  One pass over the data.
// ---------------------------------------------------------------------------
constexpr int kScanItems = 8;
constexpr int kScanTile = 256 * kScanItems;

__global__ void __launch_bounds__(256)
scan_gather_kernel(const uint32_t* __restrict__ counts, const uint64_t* __restrict__ order_items,
                   const uint32_t* __restrict__ n_dev, uint32_t* __restrict__ offsets,
                   uint32_t* __restrict__ total_out, uint32_t* __restrict__ tile_counter,
                   uint32_t* __restrict__ lookback) {
  __shared__ uint32_t s_scan[8];
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_prefix;
  const int tid = threadIdx.x;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  // only the first *n_dev sorted positions hold rendered Gaussians; the grid is sized for all P,
  // tiles past the end leave at once (nobody looks back at them)
  const uint32_t n = *n_dev;
  if (tile * (uint32_t)kScanTile >= n) return;
  const uint32_t base = tile * kScanTile + tid * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint32_t j = base + k;
    uint32_t c = 0;
    if (j < n) {
      const uint32_t g = (uint32_t)(order_items[j] & 0xffffffffu);
      c = counts[g];
    }
    v[k] = c;
    sum += c;
  }
  uint32_t block_total;
  const uint32_t excl = block_excl_scan_256(sum, s_scan, &block_total);
  if (tid == 0) {
    st_volatile_u32(lookback + tile, (tile == 0 ? kFlagInc : kFlagAgg) | block_total);
    uint32_t pre = 0;
    if (tile > 0) {
      int t = (int)tile - 1;
      while (true) {
        const uint32_t x = ld_volatile_u32(lookback + t);
        const uint32_t flag = x & ~kValMask;
        if (flag == 0) continue;
        pre += x & kValMask;
        if (flag == kFlagInc) break;
        --t;
      }
      st_volatile_u32(lookback + tile, kFlagInc | (pre + block_total));
    }
    s_prefix = pre;
    if (total_out != nullptr && block_total) atomicAdd(total_out, block_total);  // zeroed by the caller
  }
  __syncthreads();
  uint32_t run = s_prefix + excl;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint32_t j = base + k;
    if (j < n) offsets[j] = run;
    run += v[k];
  }
}

// tmp: (1 + tiles) words, zeroed here.
size_t scan_temp_bytes(int64_t n) { return (size_t)((n + kScanTile - 1) / kScanTile + 64) * 4; }

// n_max bounds *n_dev (sizes the grid); tmp ((tiles + 64) words) must be zeroed by the caller.
int launch_scan_gather(const uint32_t* counts, const uint64_t* order_items, int64_t n_max,
                       const uint32_t* n_dev, uint32_t* offsets, uint32_t* total_out, char* tmp,
                       cudaStream_t s) {
  if (n_max <= 0) return 0;
  const int64_t tiles = (n_max + kScanTile - 1) / kScanTile;
  uint32_t* words = reinterpret_cast<uint32_t*>(tmp);
  scan_gather_kernel<<<(unsigned)tiles, 256, 0, s>>>(counts, order_items, n_dev, offsets,
                                                      total_out, words, words + 64);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
