// Hand-written stable LSD "onesweep" radix sort of 64-bit items (decoupled look-back).
// Replaces the reference's use of cub::DeviceRadixSort::SortPairs in simple-knn
// (simple_knn.cu:210-213) and is exposed through gsr_sort_u64.  (The rasterizer's own
// (tile | depth) ordering no longer needs a global sort: binning.cu sorts every tile's list
// in shared memory.)
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
template <int kItemsPerThread>
__global__ void __launch_bounds__(kSortThreads)
onesweep_pass_kernel(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t n,
                     int shift, uint32_t digit_mask, const uint32_t* __restrict__ ghist,
                     uint32_t* __restrict__ tile_counter, uint32_t* __restrict__ lookback) {
  constexpr int kTileItems = kSortThreads * kItemsPerThread;
  __shared__ uint64_t s_items[kTileItems];              // 32 KB at 16 items per thread
  __shared__ uint32_t s_warp_hist[kSortWarps][kRadix];  // 8 KB
  __shared__ uint32_t s_bin_start[kRadix];
  __shared__ uint32_t s_digit_base[kRadix];
  __shared__ uint32_t s_scan[kSortWarps];
  __shared__ uint32_t s_tile;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
  for (int i = tid; i < kSortWarps * kRadix; i += kSortThreads) (&s_warp_hist[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * (uint32_t)kTileItems;
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
    const bool valid = li < tile_count;
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

  const uint32_t bin_start = block_excl_scan_256(total, s_scan, nullptr);
  const uint32_t gcount = ghist[tid];
  const uint32_t gbase = block_excl_scan_256(gcount, s_scan, nullptr);

  // Decoupled look-back, kLookWide predecessors per round trip: the flags of tiles t, t-1, ...
  // are fetched together (independent loads) and consumed nearest-first until one carries an
  // inclusive prefix or is not published yet (then the window restarts there).
  uint32_t excl = 0;
  if (tile > 0) {
    constexpr int kLookWide = 8;
    int t = (int)tile - 1;
    bool done = false;
    while (!done) {
      uint32_t v[kLookWide];
#pragma unroll
      for (int i = 0; i < kLookWide; ++i)
        v[i] = (t - i >= 0) ? ld_volatile_u32(lookback + (size_t)(t - i) * kRadix + tid) : kFlagInc;
      int used = kLookWide;
#pragma unroll
      for (int i = kLookWide - 1; i >= 0; --i) {  // nearest stop wins: scan far -> near
        const uint32_t flag = v[i] & ~kValMask;
        if (flag == 0 || flag == kFlagInc) used = i;
      }
      // tiles t .. t-used+1 hold plain aggregates; tile t-used stops the window
#pragma unroll
      for (int i = 0; i < kLookWide; ++i)
        if (i < used) excl += v[i] & kValMask;
      if (used < kLookWide) {
        uint32_t stop = 0;
#pragma unroll
        for (int i = 0; i < kLookWide; ++i)
          if (i == used) stop = v[i];
        if ((stop & ~kValMask) == kFlagInc) {
          excl += stop & kValMask;
          done = true;
        }
      }
      t -= used;  // not published yet: look at it again; all consumed: continue further back
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
    if (li < tile_count) {
      const uint32_t digit = (uint32_t)(items[k] >> shift) & digit_mask;
      s_items[s_bin_start[digit] + s_warp_hist[warp][digit] + ranks[k]] = items[k];
    }
  }
  __syncthreads();
#pragma unroll 4
  for (uint32_t i = tid; i < tile_count; i += kSortThreads) {
    const uint64_t it = s_items[i];
    const uint32_t digit = (uint32_t)(it >> shift) & digit_mask;
    out[s_digit_base[digit] + i] = it;
  }
}

static void launch_pass(int ipt, int64_t tiles, const uint64_t* src, uint64_t* dst, uint32_t n, int sh,
                        uint32_t mask, const uint32_t* ghist, uint32_t* counter, uint32_t* lookback,
                        cudaStream_t s) {
  if (ipt == kIptSmall)
    onesweep_pass_kernel<kIptSmall><<<(unsigned)tiles, kSortThreads, 0, s>>>(src, dst, n, sh, mask, ghist,
                                                                             counter, lookback);
  else
    onesweep_pass_kernel<kIptLarge><<<(unsigned)tiles, kSortThreads, 0, s>>>(src, dst, n, sh, mask, ghist,
                                                                             counter, lookback);
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
    launch_pass(sort_ipt(n), tiles, src, dst, (uint32_t)n, sh, (1u << bits) - 1u,
                ghist + (size_t)p * kRadix, counters + p, lookback + (size_t)p * tiles * kRadix, s);
    src = dst;
    dst = (dst == out) ? alt : out;
  }
  return (int)cudaGetLastError();
}

}  // namespace gsr
