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
  const float d0 = __fmul_rn(__fsqrt_rn(v2), G.inv_sqrt_bc2);
  const float d = __fadd_rn(d0, a.eps);
  p = __fmaf_rn(G.neg_step_size, __fdiv_rn(m2, d), p);
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

int launch_adam(const gsr_adam_group* groups, int n_groups, double beta1, double beta2, double eps,
                float grad_scale, cudaStream_t s) {
  if (n_groups <= 0) return 0;
  if (n_groups > kAdamMaxGroups) return GSR_ERR_BAD_ARG;
  AdamParams a;
  a.n_groups = 0;
  // ATen casts the double scalars to the fp32 op-math type inside each kernel
  a.beta1 = (float)beta1;
  a.beta2 = (float)beta2;
  a.one_minus_beta1 = (float)(1.0 - beta1);
  a.one_minus_beta2 = (float)(1.0 - beta2);
  a.eps = (float)eps;
  a.grad_scale = grad_scale;
  a.use_scale = grad_scale != 1.0f;
  int blocks = 0;
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
    blocks += (int)((g.numel + kAdamChunk - 1) / kAdamChunk);
  }
  if (blocks == 0) return 0;
  adam_kernel<<<blocks, kAdamThreads, 0, s>>>(a);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
