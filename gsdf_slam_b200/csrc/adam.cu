// Fused multi-tensor Adam: one launch updates all parameter groups.
// Replaces torch::optim::Adam::step on the reference's six groups
// (src/gaussian_model.cpp:319-352, stepped at src/gaussian_mapper.cpp:465) —
// there ~8 eager element-wise kernels per tensor with temporaries.
// HBM traffic: read p,g,m,v + write p,m,v = 28 B per element (the algorithmic
// minimum).
//
// Arithmetic follows LibTorch's single-tensor Adam step kernel by kernel, so the
// result is the one the reference's optimizer would produce:
//   m  = m*b1 ; m = m + (1-b1)*g            (mul_, add_ with alpha -> one FMA)
//   v  = v*b2 ; v = v + (1-b2)*(g*g)        (mul_, addcmul_       -> one FMA)
//   d  = sqrt(v) * (1/sqrt(bc2)) ; d = d + eps   (div by CPU scalar = mul by reciprocal; add_)
//   p  = p + (-lr/bc1) * (m / d)            (addcdiv_             -> div + FMA)
// Scalars are rounded to fp32 exactly where ATen rounds them.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdlib>

#include "gsr_capi.h"
#include "gsr_common.cuh"

namespace gsr {

constexpr int kAdamMaxGroups = 16;
#ifndef GSR_ADAM_THREADS
#define GSR_ADAM_THREADS 256
#endif
#ifndef GSR_ADAM_VEC
#define GSR_ADAM_VEC 4
#endif
#ifndef GSR_ADAM_STREAM
#define GSR_ADAM_STREAM 0
#endif
constexpr int kAdamThreads = GSR_ADAM_THREADS;
constexpr int kAdamVecPerThread = GSR_ADAM_VEC;                        // float4 per thread
constexpr int kAdamChunk = kAdamThreads * kAdamVecPerThread * 4;      // 4096 elements per CTA

struct AdamGroupDev {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long numel;
  float neg_step_size;   // -(lr / bias_correction1)
  float inv_sqrt_bc2;    // 1 / sqrt(bias_correction2)
  int first_block;       // first CTA of this group
  int vec_ok;            // pointers 16 B aligned
};
struct AdamParams {
  AdamGroupDev grp[kAdamMaxGroups];
  int n_groups;
  float beta1, one_minus_beta1, beta2, one_minus_beta2, eps, grad_scale;
  int use_scale;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamParams& a,
                                          const AdamGroupDev& G) {
  if (a.use_scale) g = __fmul_rn(g, a.grad_scale);
  const float m1 = __fmul_rn(m, a.beta1);
  const float m2 = __fmaf_rn(a.one_minus_beta1, g, m1);
  const float v1 = __fmul_rn(v, a.beta2);
  const float v2 = __fmaf_rn(a.one_minus_beta2, __fmul_rn(g, g), v1);  // a + alpha * (b * c)
  // Rows that never received a gradient have m == v == 0: sqrt(0) and 0 / d take the slow paths of the
  // IEEE routines (zero is a "special" operand for their range checks), which made this kernel issue-bound
  // on maps where most rows are untouched.  The results of the guarded forms are the same bits.
  float sq = v2;  // +0
  if (v2 != 0.f) sq = __fsqrt_rn(v2);
  const float d0 = __fmul_rn(sq, G.inv_sqrt_bc2);
  const float d = __fadd_rn(d0, a.eps);
  float q = m2;   // +-0 / d == +-0 for d > 0
  if (m2 != 0.f || !(d > 0.f)) q = __fdiv_rn(m2, d);
  p = __fmaf_rn(G.neg_step_size, q, p);
  m = m2;
  v = v2;
}

__global__ void __launch_bounds__(kAdamThreads) adam_kernel(const AdamParams a) {
  // locate the group of this CTA (<= 16 groups: linear scan)
  int gi = 0;
#pragma unroll 1
  for (int k = 1; k < a.n_groups; ++k)
    if ((int)blockIdx.x >= a.grp[k].first_block) gi = k;
  const AdamGroupDev G = a.grp[gi];
  const long long base = (long long)(blockIdx.x - G.first_block) * kAdamChunk;
  const long long end = min(G.numel, base + kAdamChunk);

  if (G.vec_ok && end - base == kAdamChunk) {
    float4 p4[kAdamVecPerThread], g4[kAdamVecPerThread], m4[kAdamVecPerThread], v4[kAdamVecPerThread];
#pragma unroll
    for (int k = 0; k < kAdamVecPerThread; ++k) {
      const long long e = base / 4 + k * kAdamThreads + threadIdx.x;
#if GSR_ADAM_STREAM
      p4[k] = __ldcs(reinterpret_cast<const float4*>(G.p) + e);
      g4[k] = __ldcs(reinterpret_cast<const float4*>(G.g) + e);
      m4[k] = __ldcs(reinterpret_cast<const float4*>(G.m) + e);
      v4[k] = __ldcs(reinterpret_cast<const float4*>(G.v) + e);
#else
      p4[k] = reinterpret_cast<const float4*>(G.p)[e];
      g4[k] = __ldcs(reinterpret_cast<const float4*>(G.g) + e);
      m4[k] = reinterpret_cast<const float4*>(G.m)[e];
      v4[k] = reinterpret_cast<const float4*>(G.v)[e];
#endif
    }
#pragma unroll
    for (int k = 0; k < kAdamVecPerThread; ++k) {
      adam_elem(p4[k].x, g4[k].x, m4[k].x, v4[k].x, a, G);
      adam_elem(p4[k].y, g4[k].y, m4[k].y, v4[k].y, a, G);
      adam_elem(p4[k].z, g4[k].z, m4[k].z, v4[k].z, a, G);
      adam_elem(p4[k].w, g4[k].w, m4[k].w, v4[k].w, a, G);
    }
#pragma unroll
    for (int k = 0; k < kAdamVecPerThread; ++k) {
      const long long e = base / 4 + k * kAdamThreads + threadIdx.x;
#if GSR_ADAM_STREAM
      __stcs(reinterpret_cast<float4*>(G.p) + e, p4[k]);
      __stcs(reinterpret_cast<float4*>(G.m) + e, m4[k]);
      __stcs(reinterpret_cast<float4*>(G.v) + e, v4[k]);
#else
      reinterpret_cast<float4*>(G.p)[e] = p4[k];
      reinterpret_cast<float4*>(G.m)[e] = m4[k];
      reinterpret_cast<float4*>(G.v)[e] = v4[k];
#endif
    }
  } else {
    for (long long e = base + threadIdx.x; e < end; e += kAdamThreads) {
      float p = G.p[e], m = G.m[e], v = G.v[e];
      adam_elem(p, G.g[e], m, v, a, G);
      G.p[e] = p;
      G.m[e] = m;
      G.v[e] = v;
    }
  }
}

// ---------------------------------------------------------------------------
// Visibility-split Adam (gsr_adam_step_rows).  Every parameter tensor of the
// Gaussian model is [rows, width]; in one mapping iteration only the rows of
// the Gaussians that were rendered (radii > 0) have a non-zero gradient.  The
// update of all OTHER rows (typically 85 %) does not depend on the backward
// pass at all — their gradient is exactly +0 — so the caller may run it on a
// side stream while the rasterizer kernels (which barely touch HBM) execute,
// and only the rendered rows are left for after the backward pass.  The two
// phases together perform bit-for-bit the update of one gsr_adam_step.
//   phase 0 (kWithGrad = false): rows with row_mask <= 0, gradient taken as +0.0f, `grad` not read
//   phase 1 (kWithGrad = true):  rows with row_mask  > 0, gradient read from `grad`
// ---------------------------------------------------------------------------
struct AdamRowsParams {
  AdamParams base;
  const int* row_mask;
  unsigned width[kAdamMaxGroups];
  unsigned magic[kAdamMaxGroups];  // ceil(2^32 / width): o / width == __umulhi(o, magic) for o * width < 2^32
};

// Phase 0 (rows WITHOUT gradient): a persistent grid-stride sweep over the 4096-element chunks of
// every group.  The grid is deliberately small (a CTA or two per SM) so that the kernel can sit
// on a side stream next to the rasterizer kernels without taking their registers away.
#ifndef GSR_ADAM_ROWS_VEC
#define GSR_ADAM_ROWS_VEC 1
#endif
constexpr int kAdamRowsVec = GSR_ADAM_ROWS_VEC;                    // float4 per thread and array
constexpr int kAdamRowsChunk = kAdamThreads * kAdamRowsVec * 4;   // 1024 elements per chunk

// kAll = false: phase 0, rows without gradient only (grad never read).
// kAll = true : every row; grad is read for the rendered rows and taken as +0 for the others
//               (whose grad memory may be uninitialised: select, never multiply).
template <bool kAll>
__device__ __forceinline__ void adam_zero_grad_chunk(const AdamRowsParams& ar, int chunk) {
  const AdamParams& a = ar.base;
  int gi = 0;
#pragma unroll 1
  for (int k = 1; k < a.n_groups; ++k)
    if (chunk >= a.grp[k].first_block) gi = k;
  const AdamGroupDev G = a.grp[gi];
  const unsigned w = ar.width[gi];
  const unsigned magic = ar.magic[gi];
  const long long base = (long long)(chunk - G.first_block) * kAdamRowsChunk;
  const long long end = min(G.numel, base + kAdamRowsChunk);
  const int* __restrict__ mask = ar.row_mask;
  // one wide division per chunk; everything below is 32-bit relative to the chunk's first row
  // (32-bit whenever the offset fits: the 64-bit division is ~100 instructions, a large part of a thread
  // that handles a single float4)
  unsigned row_base, rem_base;
  if (base <= 0xffffffffLL) {
    row_base = (unsigned)base / w;
    rem_base = (unsigned)base - row_base * w;
  } else {
    row_base = (unsigned)((unsigned long long)base / w);
    rem_base = (unsigned)((unsigned long long)base - (unsigned long long)row_base * w);
  }

  if (G.vec_ok && end - base == kAdamRowsChunk) {
    unsigned mine[kAdamRowsVec];  // bit j set <=> element j of the float4 belongs to a row WITHOUT gradient
#pragma unroll
    for (int k = 0; k < kAdamRowsVec; ++k) {
      const unsigned o = rem_base + 4u * (unsigned)(k * kAdamThreads + threadIdx.x);  // < chunk + width
      const unsigned q = (w == 1u) ? o : __umulhi(o, magic);
      const unsigned rem = o - q * w;
      const unsigned r0 = row_base + q;
      unsigned bits = 0;
      if (w >= 3u) {
        // a float4 spans at most two rows: elements 0 .. w-rem-1 lie in r0, the rest in r0 + 1
        const unsigned in_r0 = min(4u, w - rem);
        const unsigned b0 = (__ldg(mask + r0) > 0) ? 0u : 15u;
        const unsigned b1 = (in_r0 < 4u && __ldg(mask + r0 + 1) > 0) ? 0u : 15u;
        const unsigned first = (1u << in_r0) - 1u;
        bits = (b0 & first) | (b1 & ~first & 15u);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned rr = rem + j;
          const unsigned row = r0 + (rr >= w) + (rr >= 2 * w) + (rr >= 3 * w);
          bits |= (__ldg(mask + row) > 0) ? 0u : (1u << j);
        }
      }
      mine[k] = bits;
    }
    float4 p4[kAdamRowsVec], m4[kAdamRowsVec], v4[kAdamRowsVec], g4[kAdamRowsVec];
#pragma unroll
    for (int k = 0; k < kAdamRowsVec; ++k) {
      const long long e = base / 4 + k * kAdamThreads + threadIdx.x;
      g4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kAll && mine[k] != 15u) {
        const float4 g = __ldcs(reinterpret_cast<const float4*>(G.g) + e);
        g4[k].x = (mine[k] & 1u) ? 0.f : g.x;
        g4[k].y = (mine[k] & 2u) ? 0.f : g.y;
        g4[k].z = (mine[k] & 4u) ? 0.f : g.z;
        g4[k].w = (mine[k] & 8u) ? 0.f : g.w;
      }
      if (kAll || mine[k]) {
        // streaming (evict-first) accesses: this sweep must not push the rasterizer's working
        // set (splat records, moment sums) out of L2 while it runs beside the blend kernels
        p4[k] = __ldcs(reinterpret_cast<const float4*>(G.p) + e);
        m4[k] = __ldcs(reinterpret_cast<const float4*>(G.m) + e);
        v4[k] = __ldcs(reinterpret_cast<const float4*>(G.v) + e);
      }
    }
#pragma unroll
    for (int k = 0; k < kAdamRowsVec; ++k) {
      if (!kAll && !mine[k]) continue;
      adam_elem(p4[k].x, g4[k].x, m4[k].x, v4[k].x, a, G);
      adam_elem(p4[k].y, g4[k].y, m4[k].y, v4[k].y, a, G);
      adam_elem(p4[k].z, g4[k].z, m4[k].z, v4[k].z, a, G);
      adam_elem(p4[k].w, g4[k].w, m4[k].w, v4[k].w, a, G);
      const long long e = base / 4 + k * kAdamThreads + threadIdx.x;
      if (kAll || mine[k] == 15u) {
        __stcs(reinterpret_cast<float4*>(G.p) + e, p4[k]);
        __stcs(reinterpret_cast<float4*>(G.m) + e, m4[k]);
        __stcs(reinterpret_cast<float4*>(G.v) + e, v4[k]);
      } else {
        // a float4 that straddles rendered and non-rendered rows: store only the latter's elements
        const float pe[4] = {p4[k].x, p4[k].y, p4[k].z, p4[k].w};
        const float me[4] = {m4[k].x, m4[k].y, m4[k].z, m4[k].w};
        const float ve[4] = {v4[k].x, v4[k].y, v4[k].z, v4[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (mine[k] & (1u << j)) {
            G.p[e * 4 + j] = pe[j];
            G.m[e * 4 + j] = me[j];
            G.v[e * 4 + j] = ve[j];
          }
        }
      }
    }
  } else {
    for (long long e = base + threadIdx.x; e < end; e += kAdamThreads) {
      const unsigned o = rem_base + (unsigned)(e - base);
      const unsigned row = row_base + ((w == 1u) ? o : __umulhi(o, magic));
      const bool rendered = __ldg(mask + row) > 0;
      if (!kAll && rendered) continue;
      float p = G.p[e], m = G.m[e], v = G.v[e];
      adam_elem(p, (kAll && rendered) ? G.g[e] : 0.0f, m, v, a, G);
      G.p[e] = p;
      G.m[e] = m;
      G.v[e] = v;
    }
  }
}

__global__ void __launch_bounds__(kAdamThreads) adam_rows_zero_grad_kernel(const AdamRowsParams ar, int chunks) {
  for (int chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) adam_zero_grad_chunk<false>(ar, chunk);
}
__global__ void __launch_bounds__(kAdamThreads) adam_rows_all_kernel(const AdamRowsParams ar) {
  adam_zero_grad_chunk<true>(ar, (int)blockIdx.x);
}

// Phase 1 (rows WITH gradient, typically ~15 % and scattered): row-oriented.  A warp takes 32
// consecutive rows, ballots their mask and walks the marked rows; the lanes cover the
// `row_elems` (= 59 for the Gaussian model) elements of a row across all groups, so every lane
// does useful work and the group / offset of a lane's elements is resolved once, outside the loop.
constexpr int kAdamRowSlots = 4;  // rows of up to 128 elements
struct AdamVisParams {
  AdamRowsParams rows;
  unsigned prefix[kAdamMaxGroups + 1];  // element offset of each group inside a row
  unsigned row_elems;
  unsigned n_rows;
};

__global__ void __launch_bounds__(256) adam_rows_with_grad_kernel(const AdamVisParams vp) {
  const AdamParams& a = vp.rows.base;
  const int lane = threadIdx.x & 31;
  const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned num_warps = (gridDim.x * blockDim.x) >> 5;
  const int* __restrict__ mask = vp.rows.row_mask;

  // the (at most kAdamRowSlots) row elements this lane owns: e = lane + 32 * slot
  float* P[kAdamRowSlots]; const float* Gd[kAdamRowSlots]; float* Mo[kAdamRowSlots]; float* Vo[kAdamRowSlots];
  unsigned wdt[kAdamRowSlots];
  float neg_step[kAdamRowSlots], inv_bc2[kAdamRowSlots];
  bool have[kAdamRowSlots];
#pragma unroll
  for (int sl = 0; sl < kAdamRowSlots; ++sl) {
    const unsigned e = (unsigned)lane + 32u * sl;
    have[sl] = e < vp.row_elems;
    int gi = 0;
    for (int k = 1; k < a.n_groups; ++k)
      if (e >= vp.prefix[k]) gi = k;
    const unsigned off = have[sl] ? e - vp.prefix[gi] : 0u;
    P[sl] = a.grp[gi].p + off; Gd[sl] = a.grp[gi].g + off; Mo[sl] = a.grp[gi].m + off; Vo[sl] = a.grp[gi].v + off;
    wdt[sl] = vp.rows.width[gi];
    neg_step[sl] = a.grp[gi].neg_step_size;
    inv_bc2[sl] = a.grp[gi].inv_sqrt_bc2;
  }
  const unsigned n_chunks = (vp.n_rows + 31u) / 32u;
  for (unsigned c = warp_global; c < n_chunks; c += num_warps) {
    const unsigned row = c * 32u + lane;
    unsigned bal = __ballot_sync(0xffffffffu, row < vp.n_rows && __ldg(mask + row) > 0);
    while (bal) {
      const unsigned r = c * 32u + (unsigned)(__ffs(bal) - 1);
      bal &= bal - 1;
      float p[kAdamRowSlots], g[kAdamRowSlots], m[kAdamRowSlots], v[kAdamRowSlots];
#pragma unroll
      for (int sl = 0; sl < kAdamRowSlots; ++sl) {
        if (have[sl]) {
          const size_t i = (size_t)r * wdt[sl];
          p[sl] = P[sl][i]; g[sl] = __ldcs(Gd[sl] + i); m[sl] = Mo[sl][i]; v[sl] = Vo[sl][i];
        }
      }
#pragma unroll
      for (int sl = 0; sl < kAdamRowSlots; ++sl) {
        if (have[sl]) {
          AdamGroupDev Gs;
          Gs.neg_step_size = neg_step[sl];
          Gs.inv_sqrt_bc2 = inv_bc2[sl];
          adam_elem(p[sl], g[sl], m[sl], v[sl], a, Gs);
          const size_t i = (size_t)r * wdt[sl];
          P[sl][i] = p[sl]; Mo[sl][i] = m[sl]; Vo[sl][i] = v[sl];
        }
      }
    }
  }
}

static int build_adam_params(const gsr_adam_group* groups, int n_groups, double beta1, double beta2,
                             double eps, float grad_scale, AdamParams& a, int& blocks, int chunk = kAdamChunk) {
  if (n_groups > kAdamMaxGroups) return GSR_ERR_BAD_ARG;
  a.n_groups = 0;
  // ATen casts the double scalars to the fp32 op-math type inside each kernel
  a.beta1 = (float)beta1;
  a.beta2 = (float)beta2;
  a.one_minus_beta1 = (float)(1.0 - beta1);
  a.one_minus_beta2 = (float)(1.0 - beta2);
  a.eps = (float)eps;
  a.grad_scale = grad_scale;
  a.use_scale = grad_scale != 1.0f;
  blocks = 0;
  for (int k = 0; k < n_groups; ++k) {
    const gsr_adam_group& g = groups[k];
    if (g.numel <= 0) continue;
    if (g.step < 1) return GSR_ERR_BAD_ARG;
    AdamGroupDev& d = a.grp[a.n_groups++];
    d.p = g.param; d.g = g.grad; d.m = g.exp_avg; d.v = g.exp_avg_sq;
    d.numel = g.numel;
    const double bc1 = 1.0 - std::pow(beta1, (double)g.step);
    const double bc2 = 1.0 - std::pow(beta2, (double)g.step);
    const double step_size = g.lr / bc1;
    d.neg_step_size = (float)(-step_size);
    // ATen's div by a CPU scalar multiplies by the reciprocal formed in double
    // (measured on this image's torch 2.11: x / s == x * float(1.0 / s))
    d.inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
    d.first_block = blocks;
    const uintptr_t al = (uintptr_t)d.p | (uintptr_t)d.g | (uintptr_t)d.m | (uintptr_t)d.v;
    d.vec_ok = (al & 15) == 0;
    blocks += (int)((g.numel + chunk - 1) / chunk);
  }
  return 0;
}

int launch_adam(const gsr_adam_group* groups, int n_groups, double beta1, double beta2, double eps,
                float grad_scale, cudaStream_t s) {
  if (n_groups <= 0) return 0;
  AdamParams a;
  int blocks = 0;
  const int rc = build_adam_params(groups, n_groups, beta1, beta2, eps, grad_scale, a, blocks);
  if (rc) return rc;
  if (blocks == 0) return 0;
  adam_kernel<<<blocks, kAdamThreads, 0, s>>>(a);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

int launch_adam_rows(const gsr_adam_group* groups, int n_groups, double beta1, double beta2, double eps,
                     float grad_scale, int64_t rows, const int* row_mask, int phase, cudaStream_t s) {
  if (n_groups <= 0 || rows <= 0) return 0;
  if (!row_mask || phase < 0 || phase > 2) return GSR_ERR_BAD_ARG;
  AdamRowsParams ar;
  int blocks = 0;
  const int rc = build_adam_params(groups, n_groups, beta1, beta2, eps, grad_scale, ar.base, blocks,
                                   kAdamRowsChunk);
  if (rc) return rc;
  if (blocks == 0) return 0;
  ar.row_mask = row_mask;
  int k2 = 0;
  for (int k = 0; k < n_groups; ++k) {
    if (groups[k].numel <= 0) continue;
    if (groups[k].numel % rows != 0) return GSR_ERR_BAD_ARG;  // every tensor is [rows, width]
    const int64_t width = groups[k].numel / rows;
    if (width > 4096) return GSR_ERR_BAD_ARG;  // the in-CTA division is exact for o * width < 2^32
    ar.width[k2] = (unsigned)width;
    ar.magic[k2] = (unsigned)(((1ull << 32) + (uint64_t)width - 1) / (uint64_t)width);
    ++k2;
  }
  if (phase == 0) {
    // persistent: GSR_ADAM_SIDE_CTAS CTAs per SM (default 2) grid-stride over the chunks
    static const int per_sm = [] {
      const char* e = getenv("GSR_ADAM_SIDE_CTAS");
      const int v = e ? atoi(e) : 2;
      return v < 1 ? 1 : v;
    }();
    const int grid = blocks < 148 * per_sm ? blocks : 148 * per_sm;
    adam_rows_zero_grad_kernel<<<grid, kAdamThreads, 0, s>>>(ar, blocks);
  } else if (phase == 2) {
    adam_rows_all_kernel<<<blocks, kAdamThreads, 0, s>>>(ar);
  } else {
    AdamVisParams vp;
    vp.rows = ar;
    unsigned off = 0;
    for (int k = 0; k < ar.base.n_groups; ++k) {
      vp.prefix[k] = off;
      off += ar.width[k];
    }
    vp.prefix[ar.base.n_groups] = off;
    if (off > 32u * kAdamRowSlots) return GSR_ERR_BAD_ARG;  // rows wider than 128 elements are not supported
    vp.row_elems = off;
    vp.n_rows = (unsigned)rows;
    const int64_t warps = (rows + 31) / 32;
    const int64_t want = (warps + 7) / 8;
    const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    adam_rows_with_grad_kernel<<<grid, 256, 0, s>>>(vp);
  }
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
