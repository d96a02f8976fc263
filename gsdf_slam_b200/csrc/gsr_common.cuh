// Shared definitions of the B200-native rasterizer: scratch-buffer layouts,
// launch bookkeeping and the small fp32 maths helpers.
//
// Rounding contract: the per-Gaussian / per-pixel arithmetic must reproduce the
// reference's results bit-for-bit where they drive discrete decisions (radii,
// tile membership, n_contrib), so the expressions below keep the
// reference's operation order (cited per function) and are compiled with the
// same nvcc defaults (-fmad=true, no fast-math).  Kernel structure, memory
// layout and scheduling are new.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cfloat>

#define GSR_TILE_X 16
#define GSR_TILE_Y 16
#define GSR_TILE_PIX 256

namespace gsr {

extern std::atomic<long long> g_launch_count;  // bumped by GSR_LAUNCHED(); read by gsr_launch_count()
#define GSR_LAUNCHED() (gsr::g_launch_count.fetch_add(1, std::memory_order_relaxed))

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ----------------------------------------------------------------------------
// Scratch buffer layouts (opaque to callers; the reference's counterparts are
// GeometryState / ImageState / BinningState, rasterizer_impl.h:36-84).
// ----------------------------------------------------------------------------
struct GeomHeader {          // lives at the start of the geometry buffer (256 B)
  uint32_t num_dup;          // R: total (Gaussian,tile) duplicates
  uint32_t max_tile_len;     // longest per-tile list
  uint32_t num_visible;      // V: rendered Gaussians (length of vis_items)
  uint32_t num_large;        // length of the large-rectangle work queue (binning.cu)
  uint32_t num_long_tiles;   // tiles with >= kSortSmallCap entries (mid + huge)
  uint32_t num_huge_tiles;   // tiles with >= kSortMidCap entries
  uint32_t num_cand;         // Gaussians that passed the conservative cull (preprocess.cu, stage 1)
  uint32_t pad[57];
};

struct GeomView {
  GeomHeader* hdr;
  float* depths;             // [P]
  float2* means2D;           // [P]
  float4* conic_opacity;     // [P]
  float* rgb;                // [3P]
  uint32_t* tiles_touched;   // [P] exact number of tiles passing the opacity-aware test
  uint8_t* clamped;          // [P] bit c set <=> channel c was clamped
  float* acc;                // [12P] backward moment sums {Mx,My,Mxx,Mxy,Myy,M0,r,g,b,-,-,-}
  uint64_t* vis_items;       // [P] compacted rendered Gaussians (depth_bits<<32 | idx), unordered
  float4* splat_rec;         // [P][3] packed splat record of every rendered Gaussian (copied per duplicate)
  uint32_t* large_list;      // [2P] work queue of the large rectangles (emit stage)
};

constexpr int kTileHistPasses = 4;   // (kept for the stand-alone sort API)
struct ImageView {
  // tile_count | tile_fill are contiguous and zeroed by one memset before preprocess
  uint32_t* tile_count;      // [T] exact number of splats per tile (accumulated by preprocess)
  uint32_t* tile_fill;       // [T] slots handed out so far by the emit stage
  uint32_t* redo_flag;       // [T] != 0: the tile's sorted prefix was too short (or could not be selected): full sort + re-blend
  uint32_t* redo_count;      // [4] length of redo_list (word 0)
  size_t zero_bytes;         // bytes to clear starting at tile_count (the four arrays above)
  float* final_T;            // [N]
  uint32_t* n_contrib;       // [N]
  float* accum_c;            // [3N] blended colour WITHOUT background (planar)
  uint2* ranges;             // [T]
  uint32_t* max_contrib;     // [T]
  uint32_t* tile_order;      // [T] tiles sorted by list length, longest first (CTA launch order)
  uint32_t* sorted_len;      // [T] lists of kSortSmallCap+ entries: length of the depth-sorted, packed PREFIX (binning.cu)
  uint32_t* redo_list;       // [T] tiles flagged for the full sort
};

// per-tile sort (binning.cu): lists shorter than kSortSmallCap are sorted in registers by one
// warp; longer ones by a CTA-wide radix sort, in shared memory up to kSortLongSmemCap entries and
// through global scratch beyond.  kSortSmallCap must be a multiple of 4 and < 4096 (LPT bucket
// layout of tile_meta_kernel).
constexpr int kSortSmallCap = 512;
constexpr int kSortMidCap = 4092;   // the lower edge of the last bucket of the LPT order
constexpr int kSortLongSmemCap = 4096;

// Per-duplicate buffer.  Its layout must not depend on the exact R (the forward pass sizes it
// optimistically before R has reached the host): a 128-byte header holding the capacity, then
// the packed splat records (the only part the backward pass needs), then the sort arrays.
struct BinHeader {
  uint64_t capacity;         // records / items the buffer was carved for
  uint64_t pad[15];
};
struct BinView {
  BinHeader* hdr;
  float4* rec;               // [cap][3] {x, y, conic.x, conic.y | conic.z, opacity, reject_power, idx | r, g, b, 0}
  uint64_t* items;           // [cap] (depth_bits<<32 | idx), tile-major; depth-sorted per tile after tile_sort_pack
  uint64_t* scratch;         // [cap] ping-pong twin of `items` for the lists too long for shared memory
};

size_t sort_temp_bytes(int64_t n);

template <typename T>
__host__ __device__ inline T* carve(char*& p, size_t count) {
  uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 127) & ~uintptr_t(127);
  T* r = reinterpret_cast<T*>(a);
  p = reinterpret_cast<char*>(r + count);
  return r;
}

inline GeomView geom_view(char* base, size_t P) {
  GeomView g;
  char* p = base;
  g.hdr = carve<GeomHeader>(p, 1);
  g.depths = carve<float>(p, P);
  g.means2D = carve<float2>(p, P);
  g.conic_opacity = carve<float4>(p, P);
  g.rgb = carve<float>(p, 3 * P);
  g.tiles_touched = carve<uint32_t>(p, P);
  g.clamped = carve<uint8_t>(p, P);
  g.acc = carve<float>(p, 12 * P);
  g.vis_items = carve<uint64_t>(p, P);
  g.splat_rec = carve<float4>(p, 3 * P);
  g.large_list = carve<uint32_t>(p, 2 * P);
  return g;
}
inline size_t geom_bytes(size_t P) {
  GeomView g = geom_view(nullptr, P);
  return reinterpret_cast<size_t>(g.large_list) + 2 * P * 4 + 128;
}

inline ImageView image_view(char* base, size_t N, size_t T) {
  ImageView v;
  char* p = base;
  v.tile_count = carve<uint32_t>(p, T);
  v.tile_fill = carve<uint32_t>(p, T);
  v.redo_flag = carve<uint32_t>(p, T);
  v.redo_count = carve<uint32_t>(p, 4);
  v.zero_bytes = (size_t)(p - reinterpret_cast<char*>(v.tile_count));
  v.final_T = carve<float>(p, N);
  v.n_contrib = carve<uint32_t>(p, N);
  v.accum_c = carve<float>(p, 3 * N);
  v.ranges = carve<uint2>(p, T);
  v.max_contrib = carve<uint32_t>(p, T);
  v.tile_order = carve<uint32_t>(p, T);
  v.sorted_len = carve<uint32_t>(p, T);
  v.redo_list = carve<uint32_t>(p, T);
  return v;
}
inline size_t image_bytes(size_t N, size_t T) {
  ImageView v = image_view(nullptr, N, T);
  return reinterpret_cast<size_t>(v.redo_list) + T * 4 + 128;
}

inline BinView bin_view(char* base, size_t cap) {
  BinView b;
  char* p = base;
  const size_t c = cap + 8;  // slack so bulk copies may round up to 16 B multiples
  b.hdr = carve<BinHeader>(p, 1);
  b.rec = carve<float4>(p, 3 * c);
  b.items = carve<uint64_t>(p, c);
  b.scratch = carve<uint64_t>(p, c);
  return b;
}
inline size_t bin_bytes(size_t cap) {
  BinView b = bin_view(nullptr, cap);
  return reinterpret_cast<size_t>(b.scratch) + (cap + 8) * 8 + 128;
}

// ----------------------------------------------------------------------------
// fp32 helpers with the reference's operation order
// ----------------------------------------------------------------------------

// column-major 3x3 (m[c][r]), same storage convention the reference gets from glm
struct M3 {
  float m[3][3];
  __device__ __forceinline__ float* operator[](int c) { return m[c]; }
  __device__ __forceinline__ const float* operator[](int c) const { return m[c]; }
};

__device__ __forceinline__ M3 m3_make(float x1, float y1, float z1, float x2, float y2,
                                      float z2, float x3, float y3, float z3) {
  M3 r;
  r[0][0] = x1; r[0][1] = y1; r[0][2] = z1;
  r[1][0] = x2; r[1][1] = y2; r[1][2] = z2;
  r[2][0] = x3; r[2][1] = y3; r[2][2] = z3;
  return r;
}
__device__ __forceinline__ M3 m3_transpose(const M3& a) {
  return m3_make(a[0][0], a[1][0], a[2][0], a[0][1], a[1][1], a[2][1], a[0][2], a[1][2],
                 a[2][2]);
}
// (a*b)[c][r] = a[0][r]*b[c][0] + a[1][r]*b[c][1] + a[2][r]*b[c][2]   (left to right)
__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int q = 0; q < 3; ++q)
      r[c][q] = a[0][q] * b[c][0] + a[1][q] * b[c][1] + a[2][q] * b[c][2];
  return r;
}
__device__ __forceinline__ float dot3(const float* a, const float* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// view-space point: reference transformPoint4x3 (auxiliary.h:68-77)
__device__ __forceinline__ float3 xform_4x3(const float3& p, const float* m) {
  float3 t = {
      m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
      m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
      m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
  };
  return t;
}
// homogeneous point: reference transformPoint4x4 (auxiliary.h:79-88)
__device__ __forceinline__ float4 xform_4x4(const float3& p, const float* m) {
  float4 t = {m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
              m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
              m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
              m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15]};
  return t;
}

// NDC -> pixel centre; evaluated in double like the reference (auxiliary.h:33-36)
__device__ __forceinline__ float ndc_to_pix(float v, int S) {
  return ((v + 1.0) * S - 1.0) * 0.5;
}

// Parameter activations of the Gaussian model (reference GaussianModel::get*Activation,
// src/gaussian_model.cpp:48-71: torch::sigmoid, torch::exp, torch::nn::functional::normalize with
// eps = 1e-12), for the entry points that take RAW parameters.
__device__ __forceinline__ float act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float3 act_exp3(float3 s) { return {expf(s.x), expf(s.y), expf(s.z)}; }
__device__ __forceinline__ float quat_norm(float4 q) {
  return fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
}
__device__ __forceinline__ float4 act_normalize4(float4 q) {
  const float n = quat_norm(q);
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}

// tile rectangle of a disc (reference getRect, auxiliary.h:38-51)
__device__ __forceinline__ void tile_rect(const float2 p, int max_radius, int gx, int gy,
                                          uint2& rect_min, uint2& rect_max) {
  rect_min = {(unsigned)min(gx, max((int)0, (int)((p.x - max_radius) / GSR_TILE_X))),
              (unsigned)min(gy, max((int)0, (int)((p.y - max_radius) / GSR_TILE_Y)))};
  rect_max = {(unsigned)min(gx, max((int)0, (int)((p.x + max_radius + GSR_TILE_X - 1) / GSR_TILE_X))),
              (unsigned)min(gy, max((int)0, (int)((p.y + max_radius + GSR_TILE_Y - 1) / GSR_TILE_Y)))};
}

// Opacity-aware tile test of the Taming-3DGS rasterizer: smallest value of the
// Gaussian exponent 0.5*d^T*Conic*d over the pixel-centre rectangle of tile
// (tx,ty) (reference max_contrib_power_rect_gaussian_float<15,15> and
// evaluate_opacity_factor, rasterizer_impl.cu:59-118).  The tile is kept iff the
// returned value <= logf(opacity * 255)  (rasterizer_impl.cu:169-170,193).
// rcp_x = __frcp_rn(15 * 15 * co.x), rcp_z = __frcp_rn(15 * 15 * co.z): per-Gaussian constants,
// hoisted out of the per-tile evaluation by the warp-cooperative walks (same operations on the
// same inputs, so the result is bit-identical to evaluating them here)
__device__ __forceinline__ float tile_min_power_pre(const float4 co, const float2 mean, int tx, int ty,
                                                    float rcp_dxdxcox, float rcp_dydycoz) {
  const float rminx = (float)(tx * GSR_TILE_X), rminy = (float)(ty * GSR_TILE_Y);
  const float rmaxx = (float)((tx + 1) * GSR_TILE_X - 1), rmaxy = (float)((ty + 1) * GSR_TILE_Y - 1);
  const float PW = 15.0f, PH = 15.0f;

  const float x_min_diff = rminx - mean.x;
  const float x_left = x_min_diff > 0.0f;
  const float not_in_x_range = x_left + (mean.x > rmaxx);
  const float y_min_diff = rminy - mean.y;
  const float y_above = y_min_diff > 0.0f;
  const float not_in_y_range = y_above + (mean.y > rmaxy);

  float result = 0.0f;
  if ((not_in_y_range + not_in_x_range) > 0.0f) {
    const float px = x_left * rminx + (1.0f - x_left) * rmaxx;
    const float py = y_above * rminy + (1.0f - y_above) * rmaxy;
    const float dx = copysign(PW, x_min_diff);
    const float dy = copysign(PH, y_min_diff);
    const float diffx = mean.x - px;
    const float diffy = mean.y - py;
    const float tx_ = not_in_y_range * __saturatef((dx * co.x * diffx + dx * co.y * diffy) * rcp_dxdxcox);
    const float ty_ = not_in_x_range * __saturatef((dy * co.y * diffx + dy * co.z * diffy) * rcp_dydycoz);
    const float mx = px + tx_ * dx, my = py + ty_ * dy;
    const float ddx = mean.x - mx, ddy = mean.y - my;
    result = 0.5f * (co.x * ddx * ddx + co.z * ddy * ddy) + co.y * ddx * ddy;
  }
  return result;
}
__device__ __forceinline__ float tile_min_power(const float4 co, const float2 mean, int tx,
                                                int ty) {
  return tile_min_power_pre(co, mean, tx, ty, __frcp_rn(15 * 15 * co.x), __frcp_rn(15 * 15 * co.z));
}

__device__ __forceinline__ float opacity_power_threshold(float opacity) {
  const float opacity_threshold = 1.0f / 255.0f;
  return logf(opacity / opacity_threshold);
}

// Rectangles up to this many tiles are walked by their own thread; larger ones are
// walked by the whole warp (32 tiles per step), which removes the heavy tail the
// reference's one-thread-per-Gaussian loop has on near-camera splats.
#ifndef GSR_SERIAL_RECT_TILES
#define GSR_SERIAL_RECT_TILES 12
#endif
constexpr int kSerialRectTiles = GSR_SERIAL_RECT_TILES;

// Warp-cooperative walk over the tile rectangles of the lanes whose rect is large.
// Must be called by all 32 lanes.  For every large rect (lane `src`), tiles are
// visited in row-major order, 32 per step; `on_hits(src, pass_mask, pass, x, y,
// hits_before)` is called by every lane with the ballot of lanes whose tile passed
// the test (this lane's tile is (x,y)); the per-rect hit count is returned to its
// owner lane.
template <typename OnHits>
__device__ __forceinline__ uint32_t warp_walk_large_rects(bool has_large, const float4 co,
                                                          const float2 xy, const uint2 rect_min,
                                                          const uint2 rect_max, float thr,
                                                          OnHits on_hits) {
  const int lane = threadIdx.x & 31;
  uint32_t my_count = 0;
  unsigned m = __ballot_sync(0xffffffffu, has_large);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    float4 c;
    c.x = __shfl_sync(0xffffffffu, co.x, src);
    c.y = __shfl_sync(0xffffffffu, co.y, src);
    c.z = __shfl_sync(0xffffffffu, co.z, src);
    c.w = __shfl_sync(0xffffffffu, co.w, src);
    float2 p;
    p.x = __shfl_sync(0xffffffffu, xy.x, src);
    p.y = __shfl_sync(0xffffffffu, xy.y, src);
    const int x0 = __shfl_sync(0xffffffffu, (int)rect_min.x, src);
    const int y0 = __shfl_sync(0xffffffffu, (int)rect_min.y, src);
    const int w = __shfl_sync(0xffffffffu, (int)(rect_max.x - rect_min.x), src);
    const int h = __shfl_sync(0xffffffffu, (int)(rect_max.y - rect_min.y), src);
    const float th = __shfl_sync(0xffffffffu, thr, src);
    const int area = w * h;
    uint32_t count = 0;
    for (int t0 = 0; t0 < area; t0 += 32) {
      const int t = t0 + lane;
      bool pass = false;
      int x = 0, y = 0;
      if (t < area) {
        y = y0 + t / w;
        x = x0 + t - (t / w) * w;
        pass = tile_min_power(c, p, x, y) <= th;
      }
      const unsigned pm = __ballot_sync(0xffffffffu, pass);
      on_hits(src, pm, pass, x, y, count);
      count += __popc(pm);
    }
    if (lane == src) my_count = count;
  }
  return my_count;
}

// Flattened warp-cooperative walk over the tile rectangles of ALL lanes of a warp: the
// (lane, tile) pairs are enumerated 32 per step, so every step tests 32 tiles no matter how
// the rectangle sizes are distributed over the lanes (a one-thread-per-Gaussian loop runs at
// ~2 live lanes per warp when few lanes hold visible Gaussians).  Must be called by all 32
// lanes; `on_hit(owner_lane, x, y, pay0, pay1)` is invoked by each lane whose tile passed the
// test (pay0/pay1 = the owner lane's payload words), in no particular order.
#ifndef GSR_QUEUE_RECT_TILES
#define GSR_QUEUE_RECT_TILES 128
#endif
constexpr int kQueueRectTiles = GSR_QUEUE_RECT_TILES;  // emit: larger rectangles go to the work queue
// a queued rectangle is cut into at most 256 chunks (8-bit chunk id) of >= kQueueRectTiles tiles
__host__ __device__ inline uint32_t queue_chunks(uint32_t area) {
  const uint32_t n = (area + kQueueRectTiles - 1) / kQueueRectTiles;
  return n < 256u ? n : 256u;
}

template <typename OnHit>
__device__ __forceinline__ void warp_walk_rects_flat(bool has, const float4 co, const float2 xy,
                                                     const uint2 rect_min, const uint2 rect_max,
                                                     float thr, uint32_t pay0, uint32_t pay1,
                                                     OnHit on_hit) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int w_own = has ? (int)(rect_max.x - rect_min.x) : 0;
  const int area = has ? w_own * (int)(rect_max.y - rect_min.y) : 0;
  int incl = area;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(full, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(full, incl, 31);
  const int excl = incl - area;
  const float rcp_x = __frcp_rn(15 * 15 * co.x), rcp_z = __frcp_rn(15 * 15 * co.z);  // per Gaussian
  // t / w by multiplication: ceil(2^32 / w) is exact for t, w < 2^16 (w == 1: the quotient is t)
  const uint32_t w_magic = w_own > 1 ? 0xffffffffu / (uint32_t)w_own + 1u : 0u;
  for (int f0 = 0; f0 < total; f0 += 32) {
    const int f = f0 + lane;
    // owner = number of lanes whose inclusive prefix is <= f (prefix is non-decreasing)
    int owner = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const int v = __shfl_sync(full, incl, owner + step - 1);
      if (v <= f) owner += step;
    }
    float4 c;
    c.x = __shfl_sync(full, co.x, owner);
    c.y = __shfl_sync(full, co.y, owner);
    c.z = __shfl_sync(full, co.z, owner);
    c.w = __shfl_sync(full, co.w, owner);
    float2 p;
    p.x = __shfl_sync(full, xy.x, owner);
    p.y = __shfl_sync(full, xy.y, owner);
    const int x0 = __shfl_sync(full, (int)rect_min.x, owner);
    const int y0 = __shfl_sync(full, (int)rect_min.y, owner);
    const int w = __shfl_sync(full, w_own, owner);
    const float th = __shfl_sync(full, thr, owner);
    const int t = f - __shfl_sync(full, excl, owner);
    const uint32_t q0 = __shfl_sync(full, pay0, owner);
    const uint32_t q1 = __shfl_sync(full, pay1, owner);
    const float rx = __shfl_sync(full, rcp_x, owner);
    const float rz = __shfl_sync(full, rcp_z, owner);
    const uint32_t wm = __shfl_sync(full, w_magic, owner);
    if (f < total) {
      const int q = w > 1 ? (int)__umulhi((uint32_t)t, wm) : t;
      const int y = y0 + q;
      const int x = x0 + t - q * w;
      if (tile_min_power_pre(c, p, x, y, rx, rz) <= th) on_hit(owner, x, y, q0, q1);
    }
  }
}

// The counting twin of warp_walk_rects_flat (preprocess.cu): +1 on tile_count[tile] for every (Gaussian, tile) pair
// that passes the test, and the number of passing tiles of each lane's own rectangle as the return value — from the
// hit ballots of the steps its rectangle spans (no shared-memory atomics).
__device__ __forceinline__ uint32_t warp_count_rects_flat(bool has, const float4 co, const float2 xy,
                                                          const uint2 rect_min, const uint2 rect_max, float thr,
                                                          uint32_t* __restrict__ tile_count, int grid_x) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int w_own = has ? (int)(rect_max.x - rect_min.x) : 0;
  const int area = has ? w_own * (int)(rect_max.y - rect_min.y) : 0;
  int incl = area;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(full, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(full, incl, 31);
  const int excl = incl - area;
  const float rcp_x = __frcp_rn(15 * 15 * co.x), rcp_z = __frcp_rn(15 * 15 * co.z);  // per Gaussian
  // t / w by multiplication: ceil(2^32 / w) is exact for t, w < 2^16 (w == 1: the quotient is t)
  const uint32_t w_magic = w_own > 1 ? 0xffffffffu / (uint32_t)w_own + 1u : 0u;
  uint32_t mine = 0;
  for (int f0 = 0; f0 < total; f0 += 32) {
    const int f = f0 + lane;
    int owner = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
      const int v = __shfl_sync(full, incl, owner + step - 1);
      if (v <= f) owner += step;
    }
    float4 c;
    c.x = __shfl_sync(full, co.x, owner);
    c.y = __shfl_sync(full, co.y, owner);
    c.z = __shfl_sync(full, co.z, owner);
    c.w = __shfl_sync(full, co.w, owner);
    float2 p;
    p.x = __shfl_sync(full, xy.x, owner);
    p.y = __shfl_sync(full, xy.y, owner);
    const int x0 = __shfl_sync(full, (int)rect_min.x, owner);
    const int y0 = __shfl_sync(full, (int)rect_min.y, owner);
    const int w = __shfl_sync(full, w_own, owner);
    const float th = __shfl_sync(full, thr, owner);
    const int t = f - __shfl_sync(full, excl, owner);
    const float rx = __shfl_sync(full, rcp_x, owner);
    const float rz = __shfl_sync(full, rcp_z, owner);
    const uint32_t wm = __shfl_sync(full, w_magic, owner);
    bool hit = false;
    if (f < total) {
      const int q = w > 1 ? (int)__umulhi((uint32_t)t, wm) : t;
      const int y = y0 + q;
      const int x = x0 + t - q * w;
      hit = tile_min_power_pre(c, p, x, y, rx, rz) <= th;
      if (hit) atomicAdd(&tile_count[y * grid_x + x], 1u);
    }
    const unsigned hits = __ballot_sync(full, hit);
    // the part of this step that belongs to my own rectangle: flat positions [excl, incl) ∩ [f0, f0 + 32)
    const int lo = max(excl - f0, 0), hi = min(incl - f0, 32);
    if (hi > lo) {
      const unsigned below_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
      mine += (uint32_t)__popc(hits & below_hi & ~((1u << lo) - 1u));
    }
  }
  return mine;
}

// ---- launch helpers implemented across the .cu files ------------------------
int launch_sort_u64(const uint64_t* in, uint64_t* out, uint64_t* alt, int64_t n, int begin_bit,
                    int end_bit, char* tmp, cudaStream_t s, uint64_t** result);

}  // namespace gsr
