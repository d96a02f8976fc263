// distCUDA2: mean squared distance to the 3 nearest neighbours of every point.
// Replaces SimpleKNN::knn (submodules/simple-knn/simple_knn.cu:185-220).
//
// The reference orders the points along a Morton curve, cuts the order into boxes of 1024
// points and lets every query thread test all P/1024 boxes and scan the surviving ones point by
// point (O(P^2/1024) box tests).  The result — the three smallest squared distances, computed
// with the reference's expression — does not depend on how candidates are found, so this
// implementation keeps the Morton order but searches differently:
//   * no host round trips: the bounding box is reduced on the device and consumed
//     on the device (the reference does 2 blocking D2H copies, :193-199);
//   * Morton ordering uses this library's onesweep sort on the 30 bits in use;
//   * points are physically re-ordered once into float4 rows;
//   * an implicit 32-ary AABB tree (leaves of 32 points, two inner levels) is walked by one
//     WARP per leaf of 32 queries, with warp-uniform control flow: 32 child boxes tested per
//     ballot, candidate leaves read as one coalesced 512-byte row and handed round by
//     shuffles (see knn_search_kernel).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cfloat>

#include "gsr_common.cuh"

namespace gsr {

constexpr int kLeaf = 32;            // points per leaf = lanes per warp
constexpr int kKnnThreads = 256;
constexpr int kKnnSearchThreads = 128;

struct Aabb {
  float3 minn;
  float3 maxx;
};

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

// bbox[0..2] = min, bbox[3..5] = max (ordered-int encoding); initialised to the
// encoding of 0.0f because the reference reduces with init {0,0,0} (:191-199).
__global__ void knn_bbox_kernel(int P, const float* __restrict__ pts, int* __restrict__ bbox) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = pts[3 * (size_t)i + c];
      mn[c] = fminf(mn[c], v);
      mx[c] = fmaxf(mx[c], v);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
  // one atomic per CTA and component (six addresses are shared by the whole grid)
  __shared__ float s_mn[3][8], s_mx[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s_mn[c][warp] = mn[c];
      s_mx[c][warp] = mx[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    float a = s_mn[c][0], b = s_mx[c][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      a = fminf(a, s_mn[c][w]);
      b = fmaxf(b, s_mx[c][w]);
    }
    atomicMin(&bbox[c], float_to_ordered(a));
    atomicMax(&bbox[3 + c], float_to_ordered(b));
  }
}

__device__ __forceinline__ uint32_t prep_morton(uint32_t x) {
  x = (x | (x << 16)) & 0x030000FF;
  x = (x | (x << 8)) & 0x0300F00F;
  x = (x | (x << 4)) & 0x030C30C3;
  x = (x | (x << 2)) & 0x09249249;
  return x;
}

// 30-bit Morton code from bbox-normalised coordinates (reference coord2Morton, :53-61)
__global__ void knn_morton_kernel(int P, const float* __restrict__ pts, const int* __restrict__ bbox,
                                  uint64_t* __restrict__ items) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float3 minn = {ordered_to_float(bbox[0]), ordered_to_float(bbox[1]), ordered_to_float(bbox[2])};
  const float3 maxx = {ordered_to_float(bbox[3]), ordered_to_float(bbox[4]), ordered_to_float(bbox[5])};
  const float3 c = {pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2]};
  const uint32_t x = prep_morton(((c.x - minn.x) / (maxx.x - minn.x)) * ((1 << 10) - 1));
  const uint32_t y = prep_morton(((c.y - minn.y) / (maxx.y - minn.y)) * ((1 << 10) - 1));
  const uint32_t z = prep_morton(((c.z - minn.z) / (maxx.z - minn.z)) * ((1 << 10) - 1));
  const uint32_t code = x | (y << 1) | (z << 2);
  items[i] = ((uint64_t)code << 32) | (uint32_t)i;
}

// physically re-order the points: sorted_pts[j] = {xyz of point order[j], idx bits}
__global__ void knn_reorder_kernel(int P, const float* __restrict__ pts,
                                   const uint64_t* __restrict__ sorted, float4* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P) return;
  const uint32_t g = (uint32_t)(sorted[j] & 0xffffffffu);
  out[j] = make_float4(pts[3 * (size_t)g], pts[3 * (size_t)g + 1], pts[3 * (size_t)g + 2],
                       __uint_as_float(g));
}

// ---- search structure: an implicit 32-ary tree of AABBs over the Morton-sorted points ----
// Leaves hold 32 consecutive points (one per lane of the querying warp), level 1 boxes 32 leaves,
// level 2 boxes 32 level-1 boxes.  The reference (boxMinMax / boxMeanDist, :78-183) has a single
// level of 1024-point boxes that every query thread tests one by one and scans point by point.
__device__ __forceinline__ void warp_minmax(float (&mn)[3], float (&mx)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  }
}

// one warp per leaf
__global__ void __launch_bounds__(kKnnThreads)
knn_leaf_box_kernel(int P, const float4* __restrict__ sp, Aabb* __restrict__ boxes, int n_leaves) {
  const int leaf = (int)((blockIdx.x * (unsigned)kKnnThreads + threadIdx.x) >> 5);
  if (leaf >= n_leaves) return;
  const int j = leaf * kLeaf + (threadIdx.x & 31);
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (j < P) {
    const float4 p = sp[j];
    mn[0] = mx[0] = p.x;
    mn[1] = mx[1] = p.y;
    mn[2] = mx[2] = p.z;
  }
  warp_minmax(mn, mx);
  if ((threadIdx.x & 31) == 0) boxes[leaf] = {{mn[0], mn[1], mn[2]}, {mx[0], mx[1], mx[2]}};
}

// one warp per parent box: union of up to 32 children
__global__ void __launch_bounds__(kKnnThreads)
knn_parent_box_kernel(const Aabb* __restrict__ child, int n_child, Aabb* __restrict__ parent, int n_parent) {
  const int b = (int)((blockIdx.x * (unsigned)kKnnThreads + threadIdx.x) >> 5);
  if (b >= n_parent) return;
  const int c = b * 32 + (threadIdx.x & 31);
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (c < n_child) {
    const Aabb a = child[c];
    mn[0] = a.minn.x; mn[1] = a.minn.y; mn[2] = a.minn.z;
    mx[0] = a.maxx.x; mx[1] = a.maxx.y; mx[2] = a.maxx.z;
  }
  warp_minmax(mn, mx);
  if ((threadIdx.x & 31) == 0) parent[b] = {{mn[0], mn[1], mn[2]}, {mx[0], mx[1], mx[2]}};
}

// squared distance between two boxes (0 when they overlap): a lower bound of the distance between any
// query of the one and any point of the other
__device__ __forceinline__ float dist_box_box(const Aabb& a, const Aabb& q) {
  const float gx = fmaxf(0.f, fmaxf(a.minn.x - q.maxx.x, q.minn.x - a.maxx.x));
  const float gy = fmaxf(0.f, fmaxf(a.minn.y - q.maxx.y, q.minn.y - a.maxx.y));
  const float gz = fmaxf(0.f, fmaxf(a.minn.z - q.maxx.z, q.minn.z - a.maxx.z));
  return gx * gx + gy * gy + gz * gz;
}

// reference updateKBest<3> (:131-145); the distance expression is the reference's, so that the three
// smallest values — whatever the order candidates arrive in — are the reference's bit for bit
__device__ __forceinline__ void update_3best(const float3& ref, const float3& point, float* knn) {
  float3 d = {point.x - ref.x, point.y - ref.y, point.z - ref.z};
  float dist = d.x * d.x + d.y * d.y + d.z * d.z;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    if (knn[j] > dist) {
      float t = knn[j];
      knn[j] = dist;
      dist = t;
    }
  }
}

// One warp per leaf of 32 Morton-consecutive queries.  The exact 3-NN distances do not depend on the
// order or on the number of candidates examined, only on no true neighbour being skipped, so pruning
// is done once per WARP: a box is entered when its distance to the queries' own bounding box does not
// exceed the largest third-best distance any lane currently holds (with a 1e-4 relative margin for
// rounding).  Child boxes are tested 32 at a time (one per lane, one ballot), candidate leaves are
// loaded one point per lane (a single coalesced 512-byte read) and handed round with shuffles: no shared
// memory, no block-wide barrier, no per-thread divergence in the tree walk.
__device__ __forceinline__ void scan_leaf(const float4* __restrict__ sp, int P, int leaf, int skip_k,
                                          const float3& q, float* best, int lane) {
  const int j = leaf * kLeaf + lane;
  float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < P) c = sp[j];
  const int cnt = min(kLeaf, P - leaf * kLeaf);
  for (int k = 0; k < cnt; ++k) {
    const float3 o = {__shfl_sync(0xffffffffu, c.x, k), __shfl_sync(0xffffffffu, c.y, k),
                      __shfl_sync(0xffffffffu, c.z, k)};
    if (k != skip_k) update_3best(q, o, best);
  }
}

__global__ void __launch_bounds__(kKnnSearchThreads)
knn_search_kernel(int P, const float4* __restrict__ sp, const Aabb* __restrict__ box0, int n0,
                  const Aabb* __restrict__ box1, int n1, const Aabb* __restrict__ box2, int n2,
                  float* __restrict__ dists) {
  const int lane = threadIdx.x & 31;
  const int my_leaf = (int)((blockIdx.x * (unsigned)kKnnSearchThreads + threadIdx.x) >> 5);
  if (my_leaf >= n0) return;
  const int idx = my_leaf * kLeaf + lane;
  const bool valid = idx < P;
  float3 q = {0.f, 0.f, 0.f};
  uint32_t gid = 0;
  if (valid) {
    const float4 me = sp[idx];
    q = {me.x, me.y, me.z};
    gid = __float_as_uint(me.w);
  }
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  // own leaf first (every lane skips itself): a first bound from 31 Morton neighbours
  scan_leaf(sp, P, my_leaf, lane, q, best, lane);
  const Aabb qbox = box0[my_leaf];
  const float kMargin = 1.0001f;
  auto warp_bound = [&]() {
    // non-negative floats order like their bit patterns; lanes without a query do not count
    const uint32_t b = __reduce_max_sync(0xffffffffu, valid ? __float_as_uint(best[2]) : 0u);
    return __uint_as_float(b) * kMargin;
  };
  float bound = warp_bound();

  // Two tests per child box.  Coarse, 32 children per ballot: box against the queries' own box and the
  // warp-wide bound.  Exact, one child at a time (its box handed round by shuffles from the lane that
  // holds it): does ANY lane's point lie within its own third-best distance of the box?  The second
  // test matters for leaves that straddle a jump of the Morton curve: their own box is huge and the
  // coarse test passes almost everything.
  auto load_box = [&](const Aabb* __restrict__ boxes, int i, int n) {
    Aabb b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
    if (i < n) b = boxes[i];
    return b;
  };
  auto coarse = [&](const Aabb& b, int i, int n) { return i < n && dist_box_box(b, qbox) <= bound; };
  auto any_lane_needs = [&](const Aabb& mine, int src_lane) {
    Aabb b;
    b.minn.x = __shfl_sync(0xffffffffu, mine.minn.x, src_lane);
    b.minn.y = __shfl_sync(0xffffffffu, mine.minn.y, src_lane);
    b.minn.z = __shfl_sync(0xffffffffu, mine.minn.z, src_lane);
    b.maxx.x = __shfl_sync(0xffffffffu, mine.maxx.x, src_lane);
    b.maxx.y = __shfl_sync(0xffffffffu, mine.maxx.y, src_lane);
    b.maxx.z = __shfl_sync(0xffffffffu, mine.maxx.z, src_lane);
    const float gx = fmaxf(0.f, fmaxf(b.minn.x - q.x, q.x - b.maxx.x));
    const float gy = fmaxf(0.f, fmaxf(b.minn.y - q.y, q.y - b.maxx.y));
    const float gz = fmaxf(0.f, fmaxf(b.minn.z - q.z, q.z - b.maxx.z));
    const float d = gx * gx + gy * gy + gz * gz;
    return __any_sync(0xffffffffu, valid && d <= best[2] * kMargin);
  };

  for (int base2 = 0; base2 < n2; base2 += 32) {
    const Aabb b2 = load_box(box2, base2 + lane, n2);
    unsigned m2 = __ballot_sync(0xffffffffu, coarse(b2, base2 + lane, n2));
    while (m2) {
      const int c2 = __ffs(m2) - 1;
      m2 &= m2 - 1;
      if (!any_lane_needs(b2, c2)) continue;
      const int base1 = (base2 + c2) * 32;
      const Aabb b1 = load_box(box1, base1 + lane, n1);
      unsigned m1 = __ballot_sync(0xffffffffu, coarse(b1, base1 + lane, n1));
      while (m1) {
        const int c1 = __ffs(m1) - 1;
        m1 &= m1 - 1;
        if (!any_lane_needs(b1, c1)) continue;
        const int base0 = (base1 + c1) * 32;
        const Aabb b0 = load_box(box0, base0 + lane, n0);
        const int i0 = base0 + lane;
        unsigned m0 = __ballot_sync(0xffffffffu, i0 != my_leaf && coarse(b0, i0, n0));
        while (m0) {
          const int c0 = __ffs(m0) - 1;
          m0 &= m0 - 1;
          if (!any_lane_needs(b0, c0)) continue;
          scan_leaf(sp, P, base0 + c0, -1, q, best, lane);
          bound = warp_bound();
        }
      }
    }
  }
  if (valid) dists[gid] = (best[0] + best[1] + best[2]) / 3.0f;
}

size_t knn_temp_bytes(int64_t P) {
  const size_t n0 = (size_t)((P + kLeaf - 1) / kLeaf), n1 = (n0 + 31) / 32, n2 = (n1 + 31) / 32;
  return 256 + 2 * align_up((size_t)P * 8, 128) + align_up((size_t)P * 16, 128) +
         align_up(n0 * sizeof(Aabb), 128) + align_up(n1 * sizeof(Aabb), 128) + align_up(n2 * sizeof(Aabb), 128) +
         sort_temp_bytes(P) + 512;
}

int launch_knn(int P, const float* points, float* mean_dist2, char* tmp, cudaStream_t s) {
  if (P <= 0) return 0;
  char* p = tmp;
  int* bbox = carve<int>(p, 8);
  uint64_t* items_a = carve<uint64_t>(p, (size_t)P);
  uint64_t* items_b = carve<uint64_t>(p, (size_t)P);
  float4* sorted_pts = carve<float4>(p, (size_t)P);
  const int n0 = (P + kLeaf - 1) / kLeaf, n1 = (n0 + 31) / 32, n2 = (n1 + 31) / 32;
  Aabb* box0 = carve<Aabb>(p, (size_t)n0);
  Aabb* box1 = carve<Aabb>(p, (size_t)n1);
  Aabb* box2 = carve<Aabb>(p, (size_t)n2);
  char* sort_tmp = carve<char>(p, sort_temp_bytes(P));

  cudaError_t e = cudaMemsetAsync(bbox, 0, 8 * sizeof(int), s);  // ordered(0.0f) == 0
  if (e != cudaSuccess) return (int)e;
  int blocks = (P + 255) / 256;
  knn_bbox_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, s>>>(P, points, bbox);
  GSR_LAUNCHED();
  knn_morton_kernel<<<blocks, 256, 0, s>>>(P, points, bbox, items_a);
  GSR_LAUNCHED();
  // 30 Morton bits -> 4 passes (even): a -> b -> a -> b -> a
  uint64_t* sorted = nullptr;
  int rc = launch_sort_u64(items_a, items_a, items_b, P, 32, 62, sort_tmp, s, &sorted);
  if (rc) return rc;
  knn_reorder_kernel<<<blocks, 256, 0, s>>>(P, points, sorted, sorted_pts);
  GSR_LAUNCHED();
  const int wpc = kKnnThreads / 32;
  knn_leaf_box_kernel<<<(n0 + wpc - 1) / wpc, kKnnThreads, 0, s>>>(P, sorted_pts, box0, n0);
  GSR_LAUNCHED();
  knn_parent_box_kernel<<<(n1 + wpc - 1) / wpc, kKnnThreads, 0, s>>>(box0, n0, box1, n1);
  GSR_LAUNCHED();
  knn_parent_box_kernel<<<(n2 + wpc - 1) / wpc, kKnnThreads, 0, s>>>(box1, n1, box2, n2);
  GSR_LAUNCHED();
  const int swpc = kKnnSearchThreads / 32;
  knn_search_kernel<<<(n0 + swpc - 1) / swpc, kKnnSearchThreads, 0, s>>>(P, sorted_pts, box0, n0, box1, n1, box2,
                                                                        n2, mean_dist2);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
