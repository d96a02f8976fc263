// Gaussian-store surgery: prune / densify (clone + split) / append / the
// per-iteration densification statistics / opacity reset / the voxel-hash duplicate test of
// increasePcd.  Replaces the LibTorch op sequences of GaussianModel::prunePoints,
// densificationPostfix, densifyAndSplit, densifyAndClone, densifyAndPrune, addDensificationStats,
// resetOpacity and increasePcd (src/gaussian_model.cpp:201-317, 399-681).
//
// Design (B200-first, not the reference's): the reference rebuilds 22 tensors with boolean-mask
// index + cat + zeros_like up to four times per densification (clone postfix, split postfix, the
// prune inside split, the final prune) and trims the caching allocator afterwards.  Here the map
// lives in row arrays with spare capacity and a whole densifyAndPrune is three launches:
//   1. classify   one thread per source row: the clone / split / final-prune decisions of the row
//                 and of the rows it will produce (they depend only on the row's own values)
//                 -> one flag byte per row + four counts per 1024-row CTA
//   2. scan       one CTA: exclusive scan of the per-CTA counts, the four totals
//   3. scatter    one CTA per 1024 source rows: in-CTA ranks -> final destination of the kept
//                 original, its clone and its N split copies; then every array's slab is streamed
//                 once through a 3-stage shared-memory ring filled by TMA bulk copies
//                 (cp.async.bulk + mbarrier, 24 KB in flight per CTA without holding registers) and
//                 every surviving row written exactly once; the few rows that spawn clones / split
//                 copies are listed in shared memory and handled in a second, sparse pass; Adam
//                 moments of new rows and the statistics of all rows are written as zeros.
// HBM traffic = one read of the source store + one write of the destination store (the
// reference: >= 4 reads + 4 writes of everything, plus the masks' nonzero passes).
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr_capi.h"
#include "gsr_common.cuh"
#include "gsr_params.cuh"
#include "tma.cuh"

namespace gsr {

namespace {

constexpr int kStoreThreads = 256;
constexpr int kRowsPerThread = 4;
constexpr int kRowsPerCta = kStoreThreads * kRowsPerThread;  // 1024

// flag byte of a source row
constexpr uint32_t kKeep = 1;      // the row itself survives
constexpr uint32_t kClone = 2;     // a surviving clone is appended
constexpr uint32_t kSplit = 4;     // surviving split copies are appended
constexpr uint32_t kSplitSel = 8;  // selected for splitting (rank = index into the noise rows)

// how an array takes part in the scatter
enum ArrayKind : int {
  kCopyAll = 0,     // copied to the kept row, the clone and the split copies
  kMoment = 1,      // copied to the kept row, zero for new rows
  kZero = 2,        // zero for every emitted row (source not read)
  kXyz = 3,         // like kCopyAll, split copies computed
  kScaling = 4,     // like kCopyAll, split copies computed
  kKeepOnly = 5,    // copied to the kept row; clone / split never occur (prunePoints)
};

constexpr int kMaxArrays = 24;
struct StoreArrays {
  const uint32_t* src[kMaxArrays];
  uint32_t* dst[kMaxArrays];
  uint32_t wpr[kMaxArrays];    // 4-byte words per row
  uint32_t magic[kMaxArrays];  // ceil(2^32 / wpr): w / wpr == umulhi(w, magic) for w * wpr < 2^32
  int kind[kMaxArrays];
  int n;
  // split inputs (source store)
  const float* xyz;
  const float* scaling;
  const float* rotation;
  float* dst_xyz;
  float* dst_scaling;
};

struct Workspace {
  uint8_t* flags;      // [rows]
  uint4* block_count;  // [nblk]
  uint4* block_base;   // [nblk]
  uint4* totals;       // [1] + pinned copy target
  int nblk;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline int num_blocks(int64_t rows) { return (int)((rows + kRowsPerCta - 1) / kRowsPerCta); }

inline size_t workspace_bytes(int64_t rows) {
  const size_t nblk = (size_t)num_blocks(rows) + 1;
  return align_up((size_t)rows + 16, 256) + 2 * align_up(nblk * sizeof(uint4), 256) + 256;
}

inline Workspace carve(char* ws, int64_t rows) {
  Workspace w;
  w.nblk = num_blocks(rows);
  const size_t nblk = (size_t)w.nblk + 1;
  w.flags = reinterpret_cast<uint8_t*>(ws);
  ws += align_up((size_t)rows + 16, 256);
  w.block_count = reinterpret_cast<uint4*>(ws);
  ws += align_up(nblk * sizeof(uint4), 256);
  w.block_base = reinterpret_cast<uint4*>(ws);
  ws += align_up(nblk * sizeof(uint4), 256);
  w.totals = reinterpret_cast<uint4*>(ws);
  return w;
}

__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

__device__ __forceinline__ uint4 flag_counts(uint32_t f) {
  return {f & kKeep ? 1u : 0u, f & kClone ? 1u : 0u, f & kSplit ? 1u : 0u, f & kSplitSel ? 1u : 0u};
}

// CTA-wide sum of a uint4 (all threads get the result in s_out[0])
__device__ __forceinline__ uint4 cta_sum4(uint4 v, uint4* s_warp /* [8] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    v.z += __shfl_xor_sync(0xffffffffu, v.z, o);
    v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  uint4 t = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < kStoreThreads / 32; ++k) t = add4(t, s_warp[k]);
  return t;
}

// ---- 1a. classify for prunePoints / select: flag = keep where the mask byte says so ----
__global__ void __launch_bounds__(kStoreThreads)
classify_mask_kernel(const uint8_t* __restrict__ mask, int64_t rows, int keep_if_nonzero,
                     uint8_t* __restrict__ flags, uint4* __restrict__ block_count) {
  __shared__ uint4 s_warp[kStoreThreads / 32];
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerCta + (int64_t)threadIdx.x * kRowsPerThread;
  uint4 cnt = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int64_t r = row0 + k;
    if (r < rows) {
      const bool nz = mask[r] != 0;
      const uint32_t f = (nz == (keep_if_nonzero != 0)) ? kKeep : 0u;
      flags[r] = (uint8_t)f;
      cnt.x += f;
    }
  }
  const uint4 t = cta_sum4(cnt, s_warp);
  if (threadIdx.x == 0) block_count[blockIdx.x] = t;
}

// ---- 1b. classify for densifyAndPrune ----
// Per source row (all float arithmetic in the reference's op order; expf / logf / IEEE division
// are the same libdevice routines ATen's CUDA kernels call):
//   g        = xyz_gradient_accum / denom, NaN -> 0                        (gaussian_model.cpp:650-651)
//   smax     = max_k exp(scaling_k)                                        (:584, :623)
//   clone    = |g| >= max_grad  &&  smax <= percent_dense * extent         (:621-623)
//   split    =  g  >= max_grad  &&  smax >  percent_dense * extent         (:581-584; the clones
//              appended before it carry a padded gradient of 0 < max_grad and are never selected)
//   final prune of a row with opacity o and scales s: sigmoid(o) < min_opacity, or (when
//              max_screen_size != 0) max_k exp(s_k) > 0.1 * extent         (:655-661; the
//              max_radii2D test sees the zeros densificationPostfix just wrote)
//   the clone shares its original's values; a split copy has scaling log(exp(s) / (0.8 N)).
struct DensifyParams {
  float max_grad, min_opacity, pd_extent, ws_extent, inv_split_div;
  int size_tests;
};

__device__ __forceinline__ float sigmoid_aten(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kStoreThreads)
classify_densify_kernel(const float* __restrict__ accum, const float* __restrict__ denom,
                        const float* __restrict__ scaling, const float* __restrict__ opacity, int64_t rows,
                        DensifyParams p, uint8_t* __restrict__ flags, uint4* __restrict__ block_count) {
  __shared__ uint4 s_warp[kStoreThreads / 32];
  // thread t of the CTA owns rows 4t..4t+3 in the scatter kernel; here any mapping works, so
  // take the coalesced one
  uint4 cnt = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int64_t r = (int64_t)blockIdx.x * kRowsPerCta + k * kStoreThreads + threadIdx.x;
    if (r < rows) {
      float g = accum[r] / denom[r];
      if (g != g) g = 0.f;
      const float e0 = expf(scaling[3 * r]), e1 = expf(scaling[3 * r + 1]), e2 = expf(scaling[3 * r + 2]);
      const float smax = fmaxf(fmaxf(e0, e1), e2);
      const bool small = smax <= p.pd_extent;
      const bool clone = (fabsf(g) >= p.max_grad) && small;
      const bool split = (g >= p.max_grad) && !small && (smax > p.pd_extent);
      const bool low_op = sigmoid_aten(opacity[r]) < p.min_opacity;
      const bool prune_orig = low_op || (p.size_tests && smax > p.ws_extent);
      // a split copy: exp(log(exp(s) * inv)) per axis, then the max (gaussian_model.cpp:594, :659)
      const float n0 = expf(logf(e0 * p.inv_split_div)), n1 = expf(logf(e1 * p.inv_split_div)),
                  n2 = expf(logf(e2 * p.inv_split_div));
      const bool prune_split = low_op || (p.size_tests && fmaxf(fmaxf(n0, n1), n2) > p.ws_extent);
      uint32_t f = 0;
      if (!split && !prune_orig) f |= kKeep;
      if (clone && !prune_orig) f |= kClone;
      if (split) f |= kSplitSel;
      if (split && !prune_split) f |= kSplit;
      flags[r] = (uint8_t)f;
      cnt = add4(cnt, flag_counts(f));
    }
  }
  const uint4 t = cta_sum4(cnt, s_warp);
  if (threadIdx.x == 0) block_count[blockIdx.x] = t;
}

// ---- 2. exclusive scan of the per-CTA counts (one CTA) ----
__global__ void __launch_bounds__(1024)
scan_blocks_kernel(const uint4* __restrict__ block_count, int nblk, uint4* __restrict__ block_base,
                   uint4* __restrict__ totals) {
  __shared__ uint4 s_warp[32];
  __shared__ uint4 s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = {0, 0, 0, 0};
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const uint4 mine = i < nblk ? block_count[i] : uint4{0, 0, 0, 0};
    uint4 inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, inc.x, o), y = __shfl_up_sync(0xffffffffu, inc.y, o),
                     z = __shfl_up_sync(0xffffffffu, inc.z, o), w = __shfl_up_sync(0xffffffffu, inc.w, o);
      if (lane >= o) inc = add4(inc, {x, y, z, w});
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint4 v = s_warp[lane];
      uint4 winc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, winc.x, o), y = __shfl_up_sync(0xffffffffu, winc.y, o),
                       z = __shfl_up_sync(0xffffffffu, winc.z, o), w = __shfl_up_sync(0xffffffffu, winc.w, o);
        if (lane >= o) winc = add4(winc, {x, y, z, w});
      }
      s_warp[lane] = {winc.x - v.x, winc.y - v.y, winc.z - v.z, winc.w - v.w};  // exclusive over warps
    }
    __syncthreads();
    const uint4 carry = s_carry;
    const uint4 wbase = s_warp[warp];
    const uint4 incl = add4(add4(carry, wbase), inc);
    if (i < nblk) block_base[i] = {incl.x - mine.x, incl.y - mine.y, incl.z - mine.z, incl.w - mine.w};
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *totals = s_carry;
}

// ---- Philox-4x32-10 + Box-Muller (split noise when the caller supplies none) ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = {hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0};
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__device__ __forceinline__ void normal3(uint64_t seed, uint32_t idx, float (&z)[3]) {
  const uint4 r = philox4x32_10({idx, 0u, 0x5eedu, 0u}, {(uint32_t)seed, (uint32_t)(seed >> 32)});
  const float u0 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u1 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(r.z >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u3 = ((float)(r.w >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  float s, c;
  sincospif(2.0f * u1, &s, &c);
  z[0] = ra * c;
  z[1] = ra * s;
  z[2] = rb * cospif(2.0f * u3);
}

// ---- 3. scatter ----
struct ScatterParams {
  int64_t rows;
  int n_split;
  float inv_split_div;   // 1 / (0.8 N) as ATen forms it (float reciprocal of the scalar)
  const float* noise;    // [noise_rows, 3] or NULL
  uint64_t seed;
  long long noise_rows;
};

// Shared-memory plan of one CTA (dynamic): a 3-stage ring of 12 KB source chunks filled by TMA bulk
// copies (one elected thread issues them two chunks ahead, completion on an mbarrier), then the
// destination tables of its 1024 rows.
constexpr int kChunkWords = 3072;  // 12 KB
constexpr int kStages = 3;
constexpr size_t kScatterSmem = (size_t)kStages * kChunkWords * 4 + 4 * kRowsPerCta * 4 + kRowsPerCta * 2 + 64;

__global__ void __launch_bounds__(kStoreThreads)
scatter_rows_kernel(StoreArrays A, const uint8_t* __restrict__ flags, const uint4* __restrict__ block_base,
                    const uint4* __restrict__ totals, ScatterParams p) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  uint32_t* s_stage = reinterpret_cast<uint32_t*>(s_raw);                       // [kStages][kChunkWords]
  int* s_keep = reinterpret_cast<int*>(s_raw + (size_t)kStages * kChunkWords * 4);  // [kRowsPerCta] each
  int* s_clone = s_keep + kRowsPerCta;
  int* s_split = s_clone + kRowsPerCta;
  int* s_noise = s_split + kRowsPerCta;
  uint16_t* s_list = reinterpret_cast<uint16_t*>(s_noise + kRowsPerCta);        // rows that produce new rows
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_list + kRowsPerCta);         // [kStages]
  __shared__ uint4 s_warp[kStoreThreads / 32];
  __shared__ int s_nlist;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t cta_row0 = (int64_t)blockIdx.x * kRowsPerCta;
  const int nrows = (int)min((int64_t)kRowsPerCta, p.rows - cta_row0);
  if (tid == 0) {
    for (int k = 0; k < kStages; ++k) mbar_init(&s_full[k], 1);
    fence_mbar_init();
    s_nlist = 0;
  }

  // the chunk sequence: for every array that is read, its slab [cta_row0*wpr, +nrows*wpr) in pieces
  auto chunk_words = [&](int a) { return (uint32_t)nrows * A.wpr[a]; };
  auto reads = [&](int a) { return A.kind[a] != kZero && A.src[a] != nullptr; };
  struct Cursor { int a; uint32_t off; };
  auto first = [&](Cursor& c) {
    c.a = 0;
    c.off = 0;
    while (c.a < A.n && !reads(c.a)) ++c.a;
  };
  auto advance = [&](Cursor& c) {
    c.off += kChunkWords;
    if (c.off >= chunk_words(c.a)) {
      c.off = 0;
      ++c.a;
      while (c.a < A.n && !reads(c.a)) ++c.a;
    }
  };
  // words of the chunk that arrive through the bulk copy (16-byte granules, 16-byte aligned source)
  auto bulk_words = [&](const Cursor& c) {
    const uint32_t n = min((uint32_t)kChunkWords, chunk_words(c.a) - c.off);
    const uintptr_t addr = reinterpret_cast<uintptr_t>(A.src[c.a] + (size_t)cta_row0 * A.wpr[c.a] + c.off);
    return (addr & 15) ? 0u : (n & ~3u);
  };
  auto issue = [&](const Cursor& c, int stage) {
    const uint32_t bw = bulk_words(c);
    if (bw == 0) return;
    mbar_expect_tx(&s_full[stage], bw * 4);
    bulk_g2s(s_stage + (size_t)stage * kChunkWords, A.src[c.a] + (size_t)cta_row0 * A.wpr[c.a] + c.off, bw * 4,
             &s_full[stage]);
  };
  __syncthreads();
  Cursor prod;
  first(prod);
  if (tid == 0) {
    Cursor c = prod;
    for (int k = 0; k < kStages - 1 && c.a < A.n; ++k) {
      issue(c, k);
      advance(c);
    }
  }
  for (int k = 0; k < kStages - 1 && prod.a < A.n; ++k) advance(prod);  // every thread tracks the producer cursor

  const uint4 tot = *totals;
  const uint4 base = block_base[blockIdx.x];

  // ranks of this thread's four consecutive rows
  uint32_t f[kRowsPerThread];
  uint4 mine = {0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int r = tid * kRowsPerThread + k;
    f[k] = r < nrows ? flags[cta_row0 + r] : 0u;
    mine = add4(mine, flag_counts(f[k]));
  }
  uint4 inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, inc.x, o), y = __shfl_up_sync(0xffffffffu, inc.y, o),
                   z = __shfl_up_sync(0xffffffffu, inc.z, o), w = __shfl_up_sync(0xffffffffu, inc.w, o);
    if (lane >= o) inc = add4(inc, {x, y, z, w});
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint4 run = base;
  for (int k = 0; k < warp; ++k) run = add4(run, s_warp[k]);
  run = add4(run, {inc.x - mine.x, inc.y - mine.y, inc.z - mine.z, inc.w - mine.w});
  const uint32_t K = tot.x, Cn = tot.y, Ss = tot.z, Sall = tot.w;
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    const int r = tid * kRowsPerThread + k;
    s_keep[r] = (f[k] & kKeep) ? (int)run.x : -1;
    s_clone[r] = (f[k] & kClone) ? (int)(K + run.y) : -1;
    s_split[r] = (f[k] & kSplit) ? (int)(K + Cn + run.z) : -1;
    s_noise[r] = (int)run.w;
    if (f[k] & (kClone | kSplit)) s_list[atomicAdd(&s_nlist, 1)] = (uint16_t)r;
    run = add4(run, flag_counts(f[k]));
  }
  __syncthreads();

  // ---- pass 1: the kept rows, streamed through the TMA ring ----
  Cursor cur;
  first(cur);
  uint32_t phase = 0;  // one parity bit per stage, flipped whenever that stage's barrier has been waited on
  for (int it = 0; cur.a < A.n; ++it) {
    const int stage = it % kStages;
    // the stage consumed in the previous iteration is free (trailing barrier): refill it with the chunk
    // kStages-1 ahead, so that two chunks are in flight while this one is being scattered
    if (prod.a < A.n) {
      if (tid == 0) issue(prod, (it + kStages - 1) % kStages);
      advance(prod);
    }
    const int a = cur.a;
    const uint32_t wpr = A.wpr[a], magic = A.magic[a];
    const uint32_t n = min((uint32_t)kChunkWords, chunk_words(a) - cur.off);
    const uint32_t bw = bulk_words(cur);
    const uint32_t* __restrict__ src = A.src[a] + (size_t)cta_row0 * wpr + cur.off;
    uint32_t* __restrict__ dst = A.dst[a];
    if (bw) {
      mbar_wait(&s_full[stage], (phase >> stage) & 1u);
      phase ^= 1u << stage;
    }
    const uint32_t* __restrict__ sm = s_stage + (size_t)stage * kChunkWords;
#pragma unroll 4
    for (uint32_t w = tid; w < n; w += kStoreThreads) {
      const uint32_t v = w < bw ? sm[w] : src[w];
      const uint32_t g = cur.off + w;
      const uint32_t r = wpr == 1 ? g : __umulhi(g, magic);
      const int dk = s_keep[r];
      if (dk >= 0) dst[(size_t)dk * wpr + (g - r * wpr)] = v;
    }
    __syncthreads();  // the stage is free again
    advance(cur);
  }
  // arrays that are all zeros (statistics after a densification): kept rows
  for (int a = 0; a < A.n; ++a) {
    if (reads(a) || A.kind[a] == kKeepOnly) continue;
    const uint32_t wpr = A.wpr[a], magic = A.magic[a], nw = (uint32_t)nrows * wpr;
    uint32_t* __restrict__ dst = A.dst[a];
    for (uint32_t g = tid; g < nw; g += kStoreThreads) {
      const uint32_t r = wpr == 1 ? g : __umulhi(g, magic);
      const int dk = s_keep[r];
      if (dk >= 0) dst[(size_t)dk * wpr + (g - r * wpr)] = 0u;
    }
  }

  // ---- pass 2: the new rows (clones and split copies) of the listed source rows ----
  const int nlist = s_nlist;
  if (nlist == 0) return;
  for (int a = 0; a < A.n; ++a) {
    const int kind = A.kind[a];
    if (kind == kKeepOnly) continue;
    const uint32_t wpr = A.wpr[a], magic = A.magic[a];
    const bool new_zero = kind == kMoment || kind == kZero || A.src[a] == nullptr;
    const bool split_copy = kind == kCopyAll || kind == kMoment || kind == kZero;
    const uint32_t* __restrict__ src = new_zero ? nullptr : A.src[a] + (size_t)cta_row0 * wpr;
    uint32_t* __restrict__ dst = A.dst[a];
    const uint32_t nw = (uint32_t)nlist * wpr;
    for (uint32_t t = tid; t < nw; t += kStoreThreads) {
      const uint32_t li = wpr == 1 ? t : __umulhi(t, magic);
      const uint32_t c = t - li * wpr;
      const uint32_t r = s_list[li];
      const uint32_t v = new_zero ? 0u : src[r * wpr + c];
      const int dc = s_clone[r];
      if (dc >= 0) dst[(size_t)dc * wpr + c] = v;
      const int ds = s_split[r];
      if (ds >= 0 && split_copy) {
        for (int cp = 0; cp < p.n_split; ++cp) dst[((size_t)ds + (size_t)cp * Ss) * wpr + c] = v;
      }
    }
  }

  // split copies: position and scaling are computed (gaussian_model.cpp:586-594)
  if (A.xyz != nullptr) {
    for (int li = tid; li < nlist; li += kStoreThreads) {
      const int r = s_list[li];
      const int ds = s_split[r];
      if (ds < 0) continue;
      const size_t i = (size_t)(cta_row0 + r);
      const float s0 = A.scaling[3 * i], s1 = A.scaling[3 * i + 1], s2 = A.scaling[3 * i + 2];
      const float e0 = expf(s0), e1 = expf(s1), e2 = expf(s2);
      // build_rotation (include/general_utils.h:32-57), op by op, no contraction
      const float q0 = A.rotation[4 * i], q1 = A.rotation[4 * i + 1], q2 = A.rotation[4 * i + 2],
                  q3 = A.rotation[4 * i + 3];
      const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)), __fmul_rn(q2, q2)),
                                        __fmul_rn(q3, q3)));
      const float qr = q0 / nrm, x = q1 / nrm, y = q2 / nrm, z = q3 / nrm;
      auto mul = [](float a, float b) { return __fmul_rn(a, b); };
      auto add = [](float a, float b) { return __fadd_rn(a, b); };
      auto sub = [](float a, float b) { return __fsub_rn(a, b); };
      const float R00 = sub(1.f, mul(2.f, add(mul(y, y), mul(z, z))));
      const float R01 = mul(2.f, sub(mul(x, y), mul(qr, z)));
      const float R02 = mul(2.f, add(mul(x, z), mul(qr, y)));
      const float R10 = mul(2.f, add(mul(x, y), mul(qr, z)));
      const float R11 = sub(1.f, mul(2.f, add(mul(x, x), mul(z, z))));
      const float R12 = mul(2.f, sub(mul(y, z), mul(qr, x)));
      const float R20 = mul(2.f, sub(mul(x, z), mul(qr, y)));
      const float R21 = mul(2.f, add(mul(y, z), mul(qr, x)));
      const float R22 = sub(1.f, mul(2.f, add(mul(x, x), mul(y, y))));
      const float px = A.xyz[3 * i], py = A.xyz[3 * i + 1], pz = A.xyz[3 * i + 2];
      const float ns0 = logf(mul(e0, p.inv_split_div)), ns1 = logf(mul(e1, p.inv_split_div)),
                  ns2 = logf(mul(e2, p.inv_split_div));
      for (int cp = 0; cp < p.n_split; ++cp) {
        const uint32_t nidx = (uint32_t)cp * Sall + (uint32_t)s_noise[r];
        float zn[3];
        if (p.noise && (long long)nidx < p.noise_rows) {
          zn[0] = p.noise[3 * (size_t)nidx];
          zn[1] = p.noise[3 * (size_t)nidx + 1];
          zn[2] = p.noise[3 * (size_t)nidx + 2];
        } else {
          normal3(p.seed, nidx, zn);
        }
        // at::normal(mean = 0, std): z * std + 0
        const float a0 = mul(zn[0], e0), a1 = mul(zn[1], e1), a2 = mul(zn[2], e2);
        const size_t d = (size_t)ds + (size_t)cp * Ss;
        A.dst_xyz[3 * d] = add(fmaf(R02, a2, fmaf(R01, a1, mul(R00, a0))), px);
        A.dst_xyz[3 * d + 1] = add(fmaf(R12, a2, fmaf(R11, a1, mul(R10, a0))), py);
        A.dst_xyz[3 * d + 2] = add(fmaf(R22, a2, fmaf(R21, a1, mul(R20, a0))), pz);
        A.dst_scaling[3 * d] = ns0;
        A.dst_scaling[3 * d + 1] = ns1;
        A.dst_scaling[3 * d + 2] = ns2;
      }
    }
  }
}

int launch_scatter(const StoreArrays& A, const Workspace& w, const ScatterParams& p, cudaStream_t s) {
  // per call, not once per process: the attribute belongs to the current device's context
  const cudaError_t attr = cudaFuncSetAttribute(scatter_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)kScatterSmem);
  if (attr != cudaSuccess) return (int)attr;
  scatter_rows_kernel<<<w.nblk, kStoreThreads, kScatterSmem, s>>>(A, w.flags, w.block_base, w.totals, p);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// ---- select: compacted ascending index list ----
__global__ void __launch_bounds__(kStoreThreads)
write_index_kernel(const uint8_t* __restrict__ flags, int64_t rows, const uint4* __restrict__ block_base,
                   const uint4* __restrict__ totals, uint32_t* __restrict__ idx_out,
                   uint32_t* __restrict__ count_out) {
  __shared__ uint32_t s_warp[kStoreThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kRowsPerCta + (int64_t)tid * kRowsPerThread;
  uint32_t f[kRowsPerThread], mine = 0;
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    f[k] = (row0 + k < rows) ? (flags[row0 + k] & kKeep) : 0u;
    mine += f[k];
  }
  uint32_t inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint32_t run = block_base[blockIdx.x].x + inc - mine;
  for (int k = 0; k < warp; ++k) run += s_warp[k];
#pragma unroll
  for (int k = 0; k < kRowsPerThread; ++k) {
    if (f[k]) idx_out[run++] = (uint32_t)(row0 + k);
  }
  if (blockIdx.x == 0 && tid == 0 && count_out) *count_out = totals->x;
}

// ---- append / zero (densificationPostfix, increasePcd) ----
constexpr int kMaxJobs = 32;
struct CopyJobs {
  const uint32_t* src[kMaxJobs];  // NULL = zero fill
  uint32_t* dst[kMaxJobs];
  unsigned long long words[kMaxJobs];
  unsigned long long first_chunk[kMaxJobs + 1];
  int n;
};
constexpr int kCopyChunkWords = kStoreThreads * 16;

__global__ void __launch_bounds__(kStoreThreads)
copy_jobs_kernel(CopyJobs J) {
  int j = 0;
  while (j + 1 < J.n && blockIdx.x >= J.first_chunk[j + 1]) ++j;
  const unsigned long long w0 = (blockIdx.x - J.first_chunk[j]) * (unsigned long long)kCopyChunkWords;
  const unsigned long long nw = J.words[j];
  const uint32_t* __restrict__ src = J.src[j];
  uint32_t* __restrict__ dst = J.dst[j];
  uint32_t v[16];
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const unsigned long long w = w0 + (unsigned long long)u * kStoreThreads + threadIdx.x;
    v[u] = (src && w < nw) ? src[w] : 0u;
  }
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const unsigned long long w = w0 + (unsigned long long)u * kStoreThreads + threadIdx.x;
    if (w < nw) dst[w] = v[u];
  }
}

int launch_copy_jobs(CopyJobs& J, cudaStream_t s) {
  unsigned long long chunks = 0;
  int n2 = 0;
  for (int j = 0; j < J.n; ++j) {
    if (J.words[j] == 0 || J.dst[j] == nullptr) continue;
    J.src[n2] = J.src[j];
    J.dst[n2] = J.dst[j];
    J.words[n2] = J.words[j];
    J.first_chunk[n2] = chunks;
    chunks += (J.words[j] + kCopyChunkWords - 1) / kCopyChunkWords;
    ++n2;
  }
  J.n = n2;
  J.first_chunk[n2] = chunks;
  if (chunks == 0) return 0;
  if (chunks > 0x7fffffffull) return GSR_ERR_BAD_ARG;
  copy_jobs_kernel<<<(unsigned)chunks, kStoreThreads, 0, s>>>(J);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// ---- per-iteration statistics ----
__global__ void __launch_bounds__(256)
densification_stats_kernel(int64_t P, const int32_t* __restrict__ radii, const float* __restrict__ dmean2D,
                           float* __restrict__ max_radii2D, float* __restrict__ accum,
                           float* __restrict__ denom) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= P) return;
  const int32_t r = radii[i];
  if (r <= 0) return;
  max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
  const float gx = dmean2D[3 * i], gy = dmean2D[3 * i + 1];
  accum[i] = accum[i] + sqrtf(fmaf(gy, gy, gx * gx));
  denom[i] = denom[i] + 1.0f;
}

__global__ void __launch_bounds__(256)
reset_opacity_kernel(int64_t P, float* __restrict__ opacity, float* __restrict__ m, float* __restrict__ v,
                     float cap) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= P) return;
  const float x = fminf(sigmoid_aten(opacity[i]), cap);
  opacity[i] = logf(x / (1.0f - x));  // general_utils::inverse_sigmoid
  if (m) m[i] = 0.f;
  if (v) v[i] = 0.f;
}

// ---- voxel-hash duplicate test ----
constexpr long long kEmptyKey = (long long)0x8000000000000000ull;  // |hash| < 2^60: never a key

__device__ __forceinline__ long long voxel_hash(const float* __restrict__ p, float inv_voxel) {
  // torch::floor(xyz / voxel).to(kInt32), then int64 arithmetic (gaussian_model.cpp:217-221);
  // ATen divides by a scalar as a multiplication with its float reciprocal
  const int vx = (int)floorf(p[0] * inv_voxel), vy = (int)floorf(p[1] * inv_voxel),
            vz = (int)floorf(p[2] * inv_voxel);
  return (long long)vx * 73856093ll + (long long)vy * 19349663ll + (long long)vz * 83492791ll;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

__global__ void __launch_bounds__(256)
hash_clear_kernel(long long* __restrict__ table, uint64_t size) {
  const uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < size) table[i] = kEmptyKey;
}

__global__ void __launch_bounds__(256)
hash_insert_kernel(const float* __restrict__ xyz, int64_t n, float inv_voxel, long long* __restrict__ table,
                   uint64_t mask) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const long long key = voxel_hash(xyz + 3 * i, inv_voxel);
  uint64_t slot = mix64((uint64_t)key) & mask;
  while (true) {
    const long long old = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(table + slot),
                                               (unsigned long long)kEmptyKey, (unsigned long long)key);
    if (old == kEmptyKey || old == key) return;
    slot = (slot + 1) & mask;
  }
}

__global__ void __launch_bounds__(256)
hash_probe_kernel(const float* __restrict__ xyz, int64_t n, float inv_voxel, const long long* __restrict__ table,
                  uint64_t mask, uint8_t* __restrict__ keep) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const long long key = voxel_hash(xyz + 3 * i, inv_voxel);
  uint64_t slot = mix64((uint64_t)key) & mask;
  uint8_t k = 1;
  while (true) {
    const long long cur = table[slot];
    if (cur == key) { k = 0; break; }
    if (cur == kEmptyKey) break;
    slot = (slot + 1) & mask;
  }
  keep[i] = k;
}

inline uint64_t table_size(int64_t n) {
  uint64_t s = 1024;
  while (s < 2ull * (uint64_t)(n > 0 ? n : 0)) s <<= 1;
  return s;
}

// ---- increasePcd rows ----
__global__ void __launch_bounds__(256)
gather_xyz_kernel(const float* __restrict__ new_xyz, const uint32_t* __restrict__ idx, int64_t n,
                  float* __restrict__ dst_xyz) {
  const int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (w >= 3 * n) return;
  const int64_t j = w / 3;
  const int c = (int)(w - 3 * j);
  dst_xyz[w] = new_xyz[3 * (size_t)idx[j] + c];
}

__global__ void __launch_bounds__(256)
init_new_rows_kernel(const float* __restrict__ new_rgb, const uint32_t* __restrict__ idx, int64_t n,
                     const float* __restrict__ dist2, int32_t iteration, float* __restrict__ dc,
                     float* __restrict__ opacity, float* __restrict__ scaling, float* __restrict__ rotation,
                     int32_t* __restrict__ exist) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const float inv_c0 = 1.0f / 0.28209479177387814f;  // (rgb - 0.5) / C0, include/sh_utils.h:138-141
  const size_t s = idx[j];
#pragma unroll
  for (int c = 0; c < 3; ++c) dc[3 * j + c] = __fmul_rn(new_rgb[3 * s + c] - 0.5f, inv_c0);
  const float sc = logf(sqrtf(fmaxf(dist2[j], 0.0000001f)));  // gaussian_model.cpp:263-265
  scaling[3 * j] = sc;
  scaling[3 * j + 1] = sc;
  scaling[3 * j + 2] = sc;
  rotation[4 * j] = 1.f;
  rotation[4 * j + 1] = 0.f;
  rotation[4 * j + 2] = 0.f;
  rotation[4 * j + 3] = 0.f;
  const float x = 0.1f * 1.0f;
  opacity[j] = logf(x / (1.0f - x));  // inverse_sigmoid(0.1 * ones)
  exist[j] = iteration;
}

const int kGroupWidth[GSR_STORE_GROUPS] = {3, 3, 0 /* 3 * n_rest */, 1, 3, 4};

inline int group_width(const gsr_gaussian_store* st, int g) { return g == 2 ? 3 * st->n_rest : kGroupWidth[g]; }

bool store_ok(const gsr_gaussian_store* st) {
  if (!st || st->rows < 0 || st->capacity < st->rows || st->n_rest < 0 || st->n_rest > 1024) return false;
  if (st->capacity >= (1ll << 31)) return false;
  for (int g = 0; g < GSR_STORE_GROUPS; ++g) {
    if (group_width(st, g) > 0 && !st->param[g]) return false;
  }
  return st->exist_since_iter && st->xyz_gradient_accum && st->denom && st->max_radii2D;
}

void add_array(StoreArrays& A, const void* src, void* dst, int wpr, int kind) {
  if (wpr <= 0 || dst == nullptr) return;
  const int k = A.n++;
  A.src[k] = static_cast<const uint32_t*>(src);
  A.dst[k] = static_cast<uint32_t*>(dst);
  A.wpr[k] = (uint32_t)wpr;
  A.magic[k] = (uint32_t)(((1ull << 32) + (uint64_t)wpr - 1) / (uint64_t)wpr);
  A.kind[k] = kind;
}

// pinned 4-word landing pad for the counts (per thread, like gsr_forward's R word)
uint4* pinned_counts() {
  static thread_local uint4* p = nullptr;
  if (!p) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&p), sizeof(uint4), cudaHostAllocDefault) != cudaSuccess) p = nullptr;
  }
  return p;
}

int read_totals(const uint4* totals_dev, uint4& out, cudaStream_t s) {
  uint4* h = pinned_counts();
  if (!h) return GSR_ERR_ALLOC;
  cudaError_t e = cudaMemcpyAsync(h, totals_dev, sizeof(uint4), cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return (int)e;
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return (int)e;
  out = *h;
  return 0;
}

}  // namespace

}  // namespace gsr

using namespace gsr;

extern "C" {

size_t gsr_store_workspace_bytes(int64_t rows) { return workspace_bytes(rows > 0 ? rows : 0); }

int gsr_prune_points(const gsr_gaussian_store* src, const gsr_gaussian_store* dst, const uint8_t* prune_mask,
                     char* workspace, int64_t* counts_out_host, void* stream) {
  if (!store_ok(src) || !store_ok(dst) || src->n_rest != dst->n_rest || dst->capacity < src->rows)
    return GSR_ERR_BAD_ARG;
  if (counts_out_host) counts_out_host[0] = 0;
  const int64_t rows = src->rows;
  if (rows == 0) return GSR_OK;
  if (!prune_mask || !workspace) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Workspace w = carve(workspace, rows);
  classify_mask_kernel<<<w.nblk, kStoreThreads, 0, s>>>(prune_mask, rows, /*keep_if_nonzero=*/0, w.flags,
                                                        w.block_count);
  GSR_LAUNCHED();
  scan_blocks_kernel<<<1, 1024, 0, s>>>(w.block_count, w.nblk, w.block_base, w.totals);
  GSR_LAUNCHED();
  StoreArrays A{};
  for (int g = 0; g < GSR_STORE_GROUPS; ++g) {
    const int wd = group_width(src, g);
    add_array(A, src->param[g], dst->param[g], wd, kKeepOnly);
    if (src->exp_avg[g]) add_array(A, src->exp_avg[g], dst->exp_avg[g], wd, kKeepOnly);
    if (src->exp_avg_sq[g]) add_array(A, src->exp_avg_sq[g], dst->exp_avg_sq[g], wd, kKeepOnly);
  }
  add_array(A, src->exist_since_iter, dst->exist_since_iter, 1, kKeepOnly);
  add_array(A, src->xyz_gradient_accum, dst->xyz_gradient_accum, 1, kKeepOnly);
  add_array(A, src->denom, dst->denom, 1, kKeepOnly);
  add_array(A, src->max_radii2D, dst->max_radii2D, 1, kKeepOnly);
  ScatterParams p{};
  p.rows = rows;
  p.n_split = 0;
  {
    const int rc = launch_scatter(A, w, p, s);
    if (rc) return rc;
  }
  if (counts_out_host) {
    uint4 t;
    const int rc = read_totals(w.totals, t, s);
    if (rc) return rc;
    counts_out_host[0] = t.x;
  }
  return GSR_OK;
}

int gsr_densify_and_prune(const gsr_gaussian_store* src, const gsr_gaussian_store* dst,
                          const gsr_densify_args* a, char* workspace, int64_t* counts_out_host, void* stream) {
  if (!store_ok(src) || !store_ok(dst) || !a || src->n_rest != dst->n_rest) return GSR_ERR_BAD_ARG;
  if (!(a->max_grad > 0.f) || a->n_split < 1 || a->n_split > 4) return GSR_ERR_BAD_ARG;
  const int64_t rows = src->rows;
  // worst case: every row cloned, or every row replaced by N copies
  const int64_t worst = rows * (a->n_split > 2 ? a->n_split : 2);
  if (dst->capacity < worst && !counts_out_host) return GSR_ERR_BAD_ARG;
  if (counts_out_host) counts_out_host[0] = counts_out_host[1] = counts_out_host[2] = counts_out_host[3] = 0;
  if (rows == 0) return GSR_OK;
  if (!workspace) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Workspace w = carve(workspace, rows);
  DensifyParams dp;
  dp.max_grad = a->max_grad;
  dp.min_opacity = a->min_opacity;
  dp.pd_extent = a->percent_dense * a->extent;
  dp.ws_extent = 0.1f * a->extent;
  dp.inv_split_div = 1.0f / (float)(0.8 * a->n_split);
  dp.size_tests = a->max_screen_size != 0;
  classify_densify_kernel<<<w.nblk, kStoreThreads, 0, s>>>(src->xyz_gradient_accum, src->denom, src->param[4],
                                                           src->param[3], rows, dp, w.flags, w.block_count);
  GSR_LAUNCHED();
  scan_blocks_kernel<<<1, 1024, 0, s>>>(w.block_count, w.nblk, w.block_base, w.totals);
  GSR_LAUNCHED();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  uint4 t{};
  if (dst->capacity < worst) {
    // the destination is smaller than the worst case: check the real size before writing
    const int rc = read_totals(w.totals, t, s);
    if (rc) return rc;
    const int64_t need = (int64_t)t.x + t.y + (int64_t)a->n_split * t.z;
    if (need > dst->capacity) {
      counts_out_host[0] = need;
      counts_out_host[1] = t.x;
      counts_out_host[2] = t.y;
      counts_out_host[3] = t.w;
      return GSR_ERR_ALLOC;  // caller grows dst to counts_out_host[0] rows and calls again
    }
  }
  StoreArrays A{};
  static const int param_kind[GSR_STORE_GROUPS] = {kXyz, kCopyAll, kCopyAll, kCopyAll, kScaling, kCopyAll};
  for (int g = 0; g < GSR_STORE_GROUPS; ++g) {
    const int wd = group_width(src, g);
    add_array(A, src->param[g], dst->param[g], wd, param_kind[g]);
    // new rows get zero moments; a store without optimiser state on the source side gets zeros
    if (dst->exp_avg[g]) add_array(A, src->exp_avg[g], dst->exp_avg[g], wd, src->exp_avg[g] ? kMoment : kZero);
    if (dst->exp_avg_sq[g])
      add_array(A, src->exp_avg_sq[g], dst->exp_avg_sq[g], wd, src->exp_avg_sq[g] ? kMoment : kZero);
  }
  add_array(A, src->exist_since_iter, dst->exist_since_iter, 1, kCopyAll);
  add_array(A, nullptr, dst->xyz_gradient_accum, 1, kZero);
  add_array(A, nullptr, dst->denom, 1, kZero);
  add_array(A, nullptr, dst->max_radii2D, 1, kZero);
  A.xyz = src->param[0];
  A.scaling = src->param[4];
  A.rotation = src->param[5];
  A.dst_xyz = dst->param[0];
  A.dst_scaling = dst->param[4];
  ScatterParams p{};
  p.rows = rows;
  p.n_split = a->n_split;
  p.inv_split_div = dp.inv_split_div;
  p.noise = a->noise;
  p.seed = a->seed;
  p.noise_rows = a->noise ? a->noise_rows : 0;
  {
    const int rc = launch_scatter(A, w, p, s);
    if (rc) return rc;
  }
  if (counts_out_host) {
    const int rc = read_totals(w.totals, t, s);
    if (rc) return rc;
    counts_out_host[0] = (int64_t)t.x + t.y + (int64_t)a->n_split * t.z;
    counts_out_host[1] = t.x;
    counts_out_host[2] = t.y;
    counts_out_host[3] = t.w;
  }
  return GSR_OK;
}

int gsr_densification_postfix(const gsr_gaussian_store* st, const float* const* new_param,
                              const int32_t* new_exist_since_iter, int64_t n_new, void* stream) {
  if (!store_ok(st) || n_new < 0 || st->rows + n_new > st->capacity) return GSR_ERR_BAD_ARG;
  if (n_new > 0 && (!new_param || !new_exist_since_iter)) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint64_t P = (uint64_t)st->rows, n = (uint64_t)n_new;
  CopyJobs J{};
  auto job = [&](const void* src, void* dst, uint64_t words) {
    J.src[J.n] = static_cast<const uint32_t*>(src);
    J.dst[J.n] = static_cast<uint32_t*>(dst);
    J.words[J.n] = words;
    ++J.n;
  };
  for (int g = 0; g < GSR_STORE_GROUPS; ++g) {
    const uint64_t wd = (uint64_t)group_width(st, g);
    if (wd == 0) continue;
    if (n > 0 && !new_param[g]) return GSR_ERR_BAD_ARG;
    if (n > 0) job(new_param[g], st->param[g] + P * wd, n * wd);
    if (st->exp_avg[g]) job(nullptr, st->exp_avg[g] + P * wd, n * wd);
    if (st->exp_avg_sq[g]) job(nullptr, st->exp_avg_sq[g] + P * wd, n * wd);
  }
  if (n > 0) job(new_exist_since_iter, st->exist_since_iter + P, n);
  job(nullptr, st->xyz_gradient_accum, P + n);
  job(nullptr, st->denom, P + n);
  job(nullptr, st->max_radii2D, P + n);
  return launch_copy_jobs(J, s);
}

int gsr_add_densification_stats(int64_t P, const int32_t* radii, const float* dL_dmeans2D, float* max_radii2D,
                                float* xyz_gradient_accum, float* denom, void* stream) {
  if (P < 0) return GSR_ERR_BAD_ARG;
  if (P == 0) return GSR_OK;
  if (!radii || !dL_dmeans2D || !max_radii2D || !xyz_gradient_accum || !denom) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  densification_stats_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(P, radii, dL_dmeans2D, max_radii2D,
                                                                        xyz_gradient_accum, denom);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

int gsr_reset_opacity(int64_t P, float* opacity, float* exp_avg, float* exp_avg_sq, float cap, void* stream) {
  if (P < 0) return GSR_ERR_BAD_ARG;
  if (P == 0) return GSR_OK;
  if (!opacity) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  reset_opacity_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(P, opacity, exp_avg, exp_avg_sq, cap);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

size_t gsr_voxel_dedup_workspace_bytes(int64_t n_existing) { return table_size(n_existing) * sizeof(long long); }

int gsr_voxel_dedup(const float* existing_xyz, int64_t n_existing, const float* new_xyz, int64_t n_new,
                    float voxel_size, uint8_t* keep_out, char* workspace, void* stream) {
  if (n_existing < 0 || n_new < 0 || !(voxel_size > 0.f)) return GSR_ERR_BAD_ARG;
  if (n_new == 0) return GSR_OK;
  if (!new_xyz || !keep_out || !workspace || (n_existing > 0 && !existing_xyz)) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint64_t size = table_size(n_existing);
  long long* table = reinterpret_cast<long long*>(workspace);
  const float inv = 1.0f / voxel_size;
  hash_clear_kernel<<<(unsigned)((size + 255) / 256), 256, 0, s>>>(table, size);
  GSR_LAUNCHED();
  if (n_existing > 0) {
    hash_insert_kernel<<<(unsigned)((n_existing + 255) / 256), 256, 0, s>>>(existing_xyz, n_existing, inv, table,
                                                                          size - 1);
    GSR_LAUNCHED();
  }
  hash_probe_kernel<<<(unsigned)((n_new + 255) / 256), 256, 0, s>>>(new_xyz, n_new, inv, table, size - 1, keep_out);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

int gsr_select_rows(const uint8_t* mask, int64_t rows, int invert, uint32_t* idx_out, uint32_t* count_out_dev,
                    char* workspace, int64_t* counts_out_host, void* stream) {
  if (rows < 0 || rows >= (1ll << 31)) return GSR_ERR_BAD_ARG;
  if (counts_out_host) counts_out_host[0] = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (rows == 0) {
    if (count_out_dev) return (int)cudaMemsetAsync(count_out_dev, 0, sizeof(uint32_t), s);
    return GSR_OK;
  }
  if (!mask || !idx_out || !workspace) return GSR_ERR_BAD_ARG;
  Workspace w = carve(workspace, rows);
  classify_mask_kernel<<<w.nblk, kStoreThreads, 0, s>>>(mask, rows, invert ? 0 : 1, w.flags, w.block_count);
  GSR_LAUNCHED();
  scan_blocks_kernel<<<1, 1024, 0, s>>>(w.block_count, w.nblk, w.block_base, w.totals);
  GSR_LAUNCHED();
  write_index_kernel<<<w.nblk, kStoreThreads, 0, s>>>(w.flags, rows, w.block_base, w.totals, idx_out,
                                                      count_out_dev);
  GSR_LAUNCHED();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (counts_out_host) {
    uint4 t;
    const int rc = read_totals(w.totals, t, s);
    if (rc) return rc;
    counts_out_host[0] = t.x;
  }
  return GSR_OK;
}

int gsr_increase_pcd_rows(const gsr_gaussian_store* st, const float* new_xyz, const float* new_rgb,
                          const uint32_t* idx, int64_t n, int stage, const float* dist2, int32_t iteration,
                          void* stream) {
  if (!store_ok(st) || n < 0 || st->rows + n > st->capacity || (stage != 0 && stage != 1)) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const uint64_t P = (uint64_t)st->rows, nn = (uint64_t)n;
  if (stage == 0) {
    if (n == 0) return GSR_OK;
    if (!new_xyz || !idx) return GSR_ERR_BAD_ARG;
    gather_xyz_kernel<<<(unsigned)((3 * n + 255) / 256), 256, 0, s>>>(new_xyz, idx, n, st->param[0] + 3 * P);
    GSR_LAUNCHED();
    return (int)cudaGetLastError();
  }
  if (n > 0) {
    if (!new_rgb || !idx || !dist2) return GSR_ERR_BAD_ARG;
    init_new_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(new_rgb, idx, n, dist2, iteration,
                                                                     st->param[1] + 3 * P, st->param[3] + P,
                                                                     st->param[4] + 3 * P, st->param[5] + 4 * P,
                                                                     st->exist_since_iter + P);
    GSR_LAUNCHED();
  }
  CopyJobs J{};
  auto job = [&](void* dst, uint64_t words) {
    J.src[J.n] = nullptr;
    J.dst[J.n] = static_cast<uint32_t*>(dst);
    J.words[J.n] = words;
    ++J.n;
  };
  const uint64_t wrest = 3ull * (uint64_t)st->n_rest;
  if (wrest) job(st->param[2] + P * wrest, nn * wrest);
  for (int g = 0; g < GSR_STORE_GROUPS; ++g) {
    const uint64_t wd = (uint64_t)group_width(st, g);
    if (wd == 0) continue;
    if (st->exp_avg[g]) job(st->exp_avg[g] + P * wd, nn * wd);
    if (st->exp_avg_sq[g]) job(st->exp_avg_sq[g] + P * wd, nn * wd);
  }
  job(st->xyz_gradient_accum, P + nn);
  job(st->denom, P + nn);
  job(st->max_radii2D, P + nn);
  return launch_copy_jobs(J, s);
}

}  // extern "C"

// ---- PLY vertex rows (GaussianModel::savePly / loadPly, src/gaussian_model.cpp:688-953) ----
// The on-disk vertex is x y z nx ny nz f_dc_0..2 f_rest_0..3M-1 opacity scale_0..2 rot_0..3 (17 + 3M
// floats), f_rest CHANNEL-major (features_rest_.transpose(1,2).flatten(1)).  The reference moves seven
// tensors to the host (after two transposing copies) and lets tinyply interleave them property by
// property; here one kernel builds the interleaved rows in device memory (coalesced writes, the row is
// produced word by word) and the host side is a header plus one copy.
namespace gsr {
namespace {

__global__ void __launch_bounds__(256)
ply_pack_kernel(int64_t P, int M, const float* __restrict__ xyz, const float* __restrict__ dc,
                const float* __restrict__ rest, const float* __restrict__ opacity,
                const float* __restrict__ scaling, const float* __restrict__ rotation, float* __restrict__ rows) {
  const int wpr = 17 + 3 * M;
  const int64_t total = P * wpr;
  for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < total; w += (int64_t)gridDim.x * 256) {
    const int64_t i = w / wpr;
    const int c = (int)(w - i * wpr);
    float v;
    if (c < 3) v = xyz[3 * i + c];
    else if (c < 6) v = 0.f;                                       // normals
    else if (c < 9) v = dc[3 * i + (c - 6)];                       // [P,1,3] transposed = same order
    else if (c < 9 + 3 * M) {
      const int k = c - 9, ch = k / M, m = k - ch * M;             // file index ch*M + m <- rest[i][m][ch]
      v = rest[(size_t)i * 3 * M + 3 * m + ch];
    } else if (c < 10 + 3 * M) v = opacity[i];
    else if (c < 13 + 3 * M) v = scaling[3 * i + (c - 10 - 3 * M)];
    else v = rotation[4 * i + (c - 13 - 3 * M)];
    rows[w] = v;
  }
}

__global__ void __launch_bounds__(256)
ply_unpack_kernel(int64_t P, int M, const float* __restrict__ rows, float* __restrict__ xyz, float* __restrict__ dc,
                  float* __restrict__ rest, float* __restrict__ opacity, float* __restrict__ scaling,
                  float* __restrict__ rotation) {
  const int wpr = 17 + 3 * M;
  const int64_t total = P * wpr;
  for (int64_t w = (int64_t)blockIdx.x * 256 + threadIdx.x; w < total; w += (int64_t)gridDim.x * 256) {
    const int64_t i = w / wpr;
    const int c = (int)(w - i * wpr);
    const float v = rows[w];
    if (c < 3) xyz[3 * i + c] = v;
    else if (c < 6) {}
    else if (c < 9) dc[3 * i + (c - 6)] = v;
    else if (c < 9 + 3 * M) {
      const int k = c - 9, ch = k / M, m = k - ch * M;
      rest[(size_t)i * 3 * M + 3 * m + ch] = v;
    } else if (c < 10 + 3 * M) opacity[i] = v;
    else if (c < 13 + 3 * M) scaling[3 * i + (c - 10 - 3 * M)] = v;
    else rotation[4 * i + (c - 13 - 3 * M)] = v;
  }
}

}  // namespace
}  // namespace gsr

extern "C" {

int gsr_ply_pack_rows(int64_t P, int n_rest, const float* xyz, const float* features_dc, const float* features_rest,
                      const float* opacity, const float* scaling, const float* rotation, float* rows, void* stream) {
  if (P < 0 || n_rest < 0 || n_rest > 1024) return GSR_ERR_BAD_ARG;
  if (P == 0) return GSR_OK;
  if (!xyz || !features_dc || (n_rest > 0 && !features_rest) || !opacity || !scaling || !rotation || !rows)
    return GSR_ERR_BAD_ARG;
  const int64_t total = P * (17 + 3 * (int64_t)n_rest);
  const int64_t want = (total + 255) / 256;
  const unsigned grid = (unsigned)(want < 148 * 32 ? want : 148 * 32);
  gsr::ply_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(P, n_rest, xyz, features_dc, features_rest,
                                                                           opacity, scaling, rotation, rows);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

int gsr_ply_unpack_rows(int64_t P, int n_rest, const float* rows, float* xyz, float* features_dc, float* features_rest,
                        float* opacity, float* scaling, float* rotation, void* stream) {
  if (P < 0 || n_rest < 0 || n_rest > 1024) return GSR_ERR_BAD_ARG;
  if (P == 0) return GSR_OK;
  if (!xyz || !features_dc || (n_rest > 0 && !features_rest) || !opacity || !scaling || !rotation || !rows)
    return GSR_ERR_BAD_ARG;
  const int64_t total = P * (17 + 3 * (int64_t)n_rest);
  const int64_t want = (total + 255) / 256;
  const unsigned grid = (unsigned)(want < 148 * 32 ? want : 148 * 32);
  gsr::ply_unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(P, n_rest, rows, xyz, features_dc,
                                                                             features_rest, opacity, scaling, rotation);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // extern "C"
