// distCUDA2: mean squared distance to the 3 nearest neighbours of every point.
// Replaces SimpleKNN::knn (submodules/simple-knn/simple_knn.cu:185-220).
//
// Same exact-3-NN search structure as the reference (Morton order, boxes of 1024
// consecutive points, AABB pruning seeded from the +-3 Morton neighbours) so the
// candidate sets — and therefore the results — are identical, but:
//   * no host round trips: the bounding box is reduced on the device and consumed
//     on the device (the reference does 2 blocking D2H copies, :193-199);
//   * Morton ordering uses this library's onesweep sort on the 30 bits in use;
//   * points are physically re-ordered once, so the brute-force phase streams
//     contiguous float4 rows (the reference gathers points[indices[i]] in the
//     innermost loop, :175-179), staged per box through shared memory and shared
//     by all queries of the CTA that still need that box.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cfloat>

#include "gsr_common.cuh"

namespace gsr {

constexpr int kBox = 1024;
constexpr int kKnnThreads = 256;

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
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicMin(&bbox[c], float_to_ordered(mn[c]));
      atomicMax(&bbox[3 + c], float_to_ordered(mx[c]));
    }
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

// AABB of each box of 1024 Morton-consecutive points (reference boxMinMax, :78-117)
__global__ void __launch_bounds__(kKnnThreads)
knn_box_kernel(int P, const float4* __restrict__ sp, Aabb* __restrict__ boxes) {
  __shared__ float s_red[6][kKnnThreads / 32];
  const int b = blockIdx.x;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int k = threadIdx.x; k < kBox; k += kKnnThreads) {
    const int j = b * kBox + k;
    if (j < P) {
      const float4 p = sp[j];
      mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
      mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
      mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    if ((threadIdx.x & 31) == 0) {
      s_red[c][threadIdx.x >> 5] = mn[c];
      s_red[3 + c][threadIdx.x >> 5] = mx[c];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Aabb r;
    float m[6];
    for (int c = 0; c < 6; ++c) {
      m[c] = s_red[c][0];
      for (int w = 1; w < kKnnThreads / 32; ++w)
        m[c] = c < 3 ? fminf(m[c], s_red[c][w]) : fmaxf(m[c], s_red[c][w]);
    }
    r.minn = {m[0], m[1], m[2]};
    r.maxx = {m[3], m[4], m[5]};
    boxes[b] = r;
  }
}

// reference distBoxPoint (:119-129)
__device__ __forceinline__ float dist_box_point(const Aabb& box, const float3& p) {
  float3 diff = {0, 0, 0};
  if (p.x < box.minn.x || p.x > box.maxx.x) diff.x = min(abs(p.x - box.minn.x), abs(p.x - box.maxx.x));
  if (p.y < box.minn.y || p.y > box.maxx.y) diff.y = min(abs(p.y - box.minn.y), abs(p.y - box.maxx.y));
  if (p.z < box.minn.z || p.z > box.maxx.z) diff.z = min(abs(p.z - box.minn.z), abs(p.z - box.maxx.z));
  return diff.x * diff.x + diff.y * diff.y + diff.z * diff.z;
}

// reference updateKBest<3> (:131-145)
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

// One thread per Morton-sorted query (reference boxMeanDist, :147-183).  Each CTA
// walks the boxes in order; a box is staged into shared memory only if some query
// of the CTA cannot prune it, and only those queries scan it.
__global__ void __launch_bounds__(kKnnThreads)
knn_search_kernel(int P, const float4* __restrict__ sp, const Aabb* __restrict__ boxes,
                  int num_boxes, float* __restrict__ dists) {
  __shared__ float4 s_pts[kBox];
  const int idx = blockIdx.x * kKnnThreads + threadIdx.x;
  const bool valid = idx < P;
  float3 point = {0.f, 0.f, 0.f};
  uint32_t gid = 0;
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  float reject = 0.f;
  if (valid) {
    const float4 me = sp[idx];
    point = {me.x, me.y, me.z};
    gid = __float_as_uint(me.w);
    for (int i = max(0, idx - 3); i <= min(P - 1, idx + 3); i++) {
      if (i == idx) continue;
      const float4 o = sp[i];
      update_3best(point, {o.x, o.y, o.z}, best);
    }
    reject = best[2];
    best[0] = FLT_MAX;
    best[1] = FLT_MAX;
    best[2] = FLT_MAX;
  }

  for (int b = 0; b < num_boxes; b++) {
    bool need = false;
    if (valid) {
      const Aabb box = boxes[b];
      const float dist = dist_box_point(box, point);
      need = !(dist > reject || dist > best[2]);
    }
    if (!__syncthreads_or(need)) continue;
    const int start = b * kBox;
    const int cnt = min(kBox, P - start);
    for (int k = threadIdx.x; k < cnt; k += kKnnThreads) s_pts[k] = sp[start + k];
    __syncthreads();
    if (need) {
      for (int k = 0; k < cnt; k++) {
        if (start + k == idx) continue;
        const float4 o = s_pts[k];
        update_3best(point, {o.x, o.y, o.z}, best);
      }
    }
    __syncthreads();
  }
  if (valid) dists[gid] = (best[0] + best[1] + best[2]) / 3.0f;
}

size_t knn_temp_bytes(int64_t P) {
  size_t nb = (size_t)((P + kBox - 1) / kBox);
  return 256 + 2 * align_up((size_t)P * 8, 128) + align_up((size_t)P * 16, 128) +
         align_up(nb * sizeof(Aabb), 128) + sort_temp_bytes(P) + 512;
}

int launch_knn(int P, const float* points, float* mean_dist2, char* tmp, cudaStream_t s) {
  if (P <= 0) return 0;
  char* p = tmp;
  int* bbox = carve<int>(p, 8);
  uint64_t* items_a = carve<uint64_t>(p, (size_t)P);
  uint64_t* items_b = carve<uint64_t>(p, (size_t)P);
  float4* sorted_pts = carve<float4>(p, (size_t)P);
  const int num_boxes = (P + kBox - 1) / kBox;
  Aabb* boxes = carve<Aabb>(p, (size_t)num_boxes);
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
  knn_box_kernel<<<num_boxes, kKnnThreads, 0, s>>>(P, sorted_pts, boxes);
  GSR_LAUNCHED();
  knn_search_kernel<<<(P + kKnnThreads - 1) / kKnnThreads, kKnnThreads, 0, s>>>(
      P, sorted_pts, boxes, num_boxes, mean_dist2);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
