// Fused L1 + SSIM photometric loss, forward and backward.
//
// Replaces, between the rasterizer's forward and backward passes, what the reference does with
// ~25 LibTorch kernels plus their autograd graph (src/gaussian_mapper.cpp:414-431,
// include/loss_utils.h:28-124):
//   masked_image = where(mask > 0, image, 1);  masked_gt = where(mask > 0, gt, 1)
//   L1   = mean |masked_image - masked_gt|
//   SSIM = mean ssim_map(masked_image, gt)     <- the UNMASKED gt, as the reference does (:429)
//          (11x11 gaussian window, sigma 1.5, zero padding, per channel; C1 = 0.01^2, C2 = 0.03^2)
//   loss = (1 - lambda) * L1 + lambda * (1 - SSIM);   dL/d(image) = d loss / d image
//
// Two tile kernels, both separable 11-tap convolutions staged in shared memory:
//   ssim_fwd_kernel : mu1, mu2, E[x^2], E[y^2], E[xy] per pixel -> ssim_map; block-reduced sums of
//                     |x - y| and ssim_map; the three per-pixel partials dS/dmu1, dS/dE[x^2],
//                     dS/dE[xy] (all other dependencies of ssim_map on the image go through them)
//   ssim_bwd_kernel : dL/dx(p) = -lambda/N * ( conv(dS/dmu1) + 2 x(p) conv(dS/dE[x^2])
//                     + y(p) conv(dS/dE[xy]) )(p) + (1 - lambda)/N * sign(x - y_masked)(p), masked
// HBM traffic: 2 + 3 image-sized reads/writes forward, 5 + 1 backward (the reference moves ~60).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>

#include "gsr_capi.h"
#include "gsr_common.cuh"

namespace gsr {

constexpr int kLossTileW = 32, kLossTileH = 16, kLossHalo = 5, kLossTaps = 11;
constexpr int kLossInW = kLossTileW + 2 * kLossHalo;  // 42
constexpr int kLossInH = kLossTileH + 2 * kLossHalo;  // 26
constexpr int kLossThreads = 256;                     // 32 x 8, two output rows per thread

struct LossWeights { float w[kLossTaps]; };

struct LossParams {
  const float* image;   // [3,H,W] rendered
  const float* gt;      // [3,H,W]
  const float* mask;    // [H,W] or NULL
  int H, W;
  float lambda;
  float* maps;          // [3 maps][3,H,W] workspace
  float* sums;          // [2] {sum |x - y|, sum ssim_map} (zeroed by the launcher)
  float* loss_out;      // [3] {loss, L1, SSIM}
  float* dL_dimage;     // [3,H,W]
  LossWeights g;
};

__global__ void __launch_bounds__(kLossThreads) ssim_fwd_kernel(const LossParams a) {
  __shared__ float s_x[kLossInH][kLossInW];
  __shared__ float s_y[kLossInH][kLossInW];
  __shared__ float s_h[5][kLossInH][kLossTileW];
  __shared__ float s_red[2][kLossThreads / 32];
  const int tid = threadIdx.x;
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * kLossTileW, y0 = blockIdx.y * kLossTileH;
  const size_t plane = (size_t)a.H * a.W;
  const float* img = a.image + c * plane;
  const float* gt = a.gt + c * plane;

  for (int i = tid; i < kLossInH * kLossInW; i += kLossThreads) {
    const int r = i / kLossInW, col = i - r * kLossInW;
    const int gy = y0 - kLossHalo + r, gx = x0 - kLossHalo + col;
    float xv = 0.f, yv = 0.f;  // zero padding outside the image
    if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
      const size_t p = (size_t)gy * a.W + gx;
      xv = img[p];
      if (a.mask != nullptr && !(a.mask[p] > 0.f)) xv = 1.0f;
      yv = gt[p];
    }
    s_x[r][col] = xv;
    s_y[r][col] = yv;
  }
  __syncthreads();
  // horizontal pass
  for (int i = tid; i < kLossInH * kLossTileW; i += kLossThreads) {
    const int r = i / kLossTileW, col = i - r * kLossTileW;
    float m1 = 0.f, m2 = 0.f, m11 = 0.f, m22 = 0.f, m12 = 0.f;
#pragma unroll
    for (int k = 0; k < kLossTaps; ++k) {
      const float w = a.g.w[k], xv = s_x[r][col + k], yv = s_y[r][col + k];
      m1 += w * xv;
      m2 += w * yv;
      m11 += w * xv * xv;
      m22 += w * yv * yv;
      m12 += w * xv * yv;
    }
    s_h[0][r][col] = m1; s_h[1][r][col] = m2; s_h[2][r][col] = m11; s_h[3][r][col] = m22; s_h[4][r][col] = m12;
  }
  __syncthreads();
  // vertical pass + ssim map
  const int tx = tid & 31, ty = tid >> 5;
  float l1 = 0.f, ss = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = ty + 8 * half;
    const int gy = y0 + r, gx = x0 + tx;
    if (gy < a.H && gx < a.W) {
      float mu1 = 0.f, mu2 = 0.f, m11 = 0.f, m22 = 0.f, m12 = 0.f;
#pragma unroll
      for (int k = 0; k < kLossTaps; ++k) {
        const float w = a.g.w[k];
        mu1 += w * s_h[0][r + k][tx];
        mu2 += w * s_h[1][r + k][tx];
        m11 += w * s_h[2][r + k][tx];
        m22 += w * s_h[3][r + k][tx];
        m12 += w * s_h[4][r + k][tx];
      }
      const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
      const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
      const float s11 = m11 - mu1_sq, s22 = m22 - mu2_sq, s12 = m12 - mu12;
      const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2;
      const float B1 = mu1_sq + mu2_sq + C1, B2 = s11 + s22 + C2;
      const float inv = 1.0f / (B1 * B2);
      const float S = A1 * A2 * inv;
      ss += S;
      // partials of S w.r.t. mu1 (total: also through s11 and s12), E[x^2], E[xy]
      const float dS_dA1 = A2 * inv, dS_dA2 = A1 * inv, dS_dB1 = -S / B1, dS_dB2 = -S / B2;
      const float g_mu1 = dS_dA1 * (2.f * mu2) + dS_dA2 * (-2.f * mu2) + dS_dB1 * (2.f * mu1) + dS_dB2 * (-2.f * mu1);
      const float g_m11 = dS_dB2;
      const float g_m12 = 2.f * dS_dA2;
      const size_t p = (size_t)gy * a.W + gx;
      const size_t n3 = 3 * plane;
      a.maps[0 * n3 + c * plane + p] = g_mu1;
      a.maps[1 * n3 + c * plane + p] = g_m11;
      a.maps[2 * n3 + c * plane + p] = g_m12;
      const float xm = s_x[r + kLossHalo][tx + kLossHalo];
      float ym = s_y[r + kLossHalo][tx + kLossHalo];
      if (a.mask != nullptr && !(a.mask[p] > 0.f)) ym = 1.0f;
      l1 += fabsf(xm - ym);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((tid & 31) == 0) { s_red[0][tid >> 5] = l1; s_red[1][tid >> 5] = ss; }
  __syncthreads();
  if (tid == 0) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) { t1 += s_red[0][w]; t2 += s_red[1][w]; }
    atomicAdd(&a.sums[0], t1);
    atomicAdd(&a.sums[1], t2);
  }
}

__global__ void __launch_bounds__(kLossThreads) ssim_bwd_kernel(const LossParams a) {
  __shared__ float s_g[3][kLossInH][kLossInW];
  __shared__ float s_h[3][kLossInH][kLossTileW];
  const int tid = threadIdx.x;
  const int c = blockIdx.z;
  const int x0 = blockIdx.x * kLossTileW, y0 = blockIdx.y * kLossTileH;
  const size_t plane = (size_t)a.H * a.W, n3 = 3 * plane;
  const float N = (float)n3;
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && tid == 0 && a.loss_out != nullptr) {
    const float l1 = a.sums[0] / N, ssim = a.sums[1] / N;
    a.loss_out[0] = (1.0f - a.lambda) * l1 + a.lambda * (1.0f - ssim);
    a.loss_out[1] = l1;
    a.loss_out[2] = ssim;
  }
  for (int i = tid; i < kLossInH * kLossInW; i += kLossThreads) {
    const int r = i / kLossInW, col = i - r * kLossInW;
    const int gy = y0 - kLossHalo + r, gx = x0 - kLossHalo + col;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
      const size_t p = c * plane + (size_t)gy * a.W + gx;
      v0 = a.maps[p]; v1 = a.maps[n3 + p]; v2 = a.maps[2 * n3 + p];
    }
    s_g[0][r][col] = v0; s_g[1][r][col] = v1; s_g[2][r][col] = v2;
  }
  __syncthreads();
  for (int i = tid; i < kLossInH * kLossTileW; i += kLossThreads) {
    const int r = i / kLossTileW, col = i - r * kLossTileW;
    float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
    for (int k = 0; k < kLossTaps; ++k) {
      const float w = a.g.w[k];
      h0 += w * s_g[0][r][col + k];
      h1 += w * s_g[1][r][col + k];
      h2 += w * s_g[2][r][col + k];
    }
    s_h[0][r][col] = h0; s_h[1][r][col] = h1; s_h[2][r][col] = h2;
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = ty + 8 * half;
    const int gy = y0 + r, gx = x0 + tx;
    if (gy < a.H && gx < a.W) {
      float c0 = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
      for (int k = 0; k < kLossTaps; ++k) {
        const float w = a.g.w[k];
        c0 += w * s_h[0][r + k][tx];
        c1 += w * s_h[1][r + k][tx];
        c2 += w * s_h[2][r + k][tx];
      }
      const size_t p = (size_t)gy * a.W + gx;
      const bool on = a.mask == nullptr || a.mask[p] > 0.f;
      float grad = 0.f;
      if (on) {
        const float x = a.image[c * plane + p], y = a.gt[c * plane + p];
        const float dS = c0 + 2.f * x * c1 + y * c2;
        const float d = x - y;  // masked_gt == gt where the mask is on
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        grad = (1.0f - a.lambda) / N * sgn - a.lambda / N * dS;
      }
      a.dL_dimage[c * plane + p] = grad;
    }
  }
}

}  // namespace gsr

using namespace gsr;

extern "C" {

size_t gsr_l1_ssim_workspace_bytes(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  return (size_t)9 * H * W * sizeof(float) + 256;
}

int gsr_l1_ssim_loss(const float* image, const float* gt, const float* mask, int H, int W, float lambda_dssim,
                     float* loss_out, float* dL_dimage, char* workspace, void* stream) {
  if (H <= 0 || W <= 0 || !image || !gt || !dL_dimage || !workspace) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  LossParams a;
  a.image = image; a.gt = gt; a.mask = mask; a.H = H; a.W = W; a.lambda = lambda_dssim;
  a.sums = reinterpret_cast<float*>(workspace);
  a.maps = reinterpret_cast<float*>(workspace + 256);
  a.loss_out = loss_out;
  a.dL_dimage = dL_dimage;
  // the reference's window (loss_utils.h:48-63): exp(-(x-5)^2 / (2 sigma^2)) in fp32, normalised
  float gsum = 0.f;
  for (int x = 0; x < kLossTaps; ++x) {
    const int t = x - kLossTaps / 2;
    a.g.w[x] = std::exp(-t * t / (2.0f * 1.5f * 1.5f));
    gsum += a.g.w[x];
  }
  for (int x = 0; x < kLossTaps; ++x) a.g.w[x] /= gsum;
  cudaError_t e = cudaMemsetAsync(a.sums, 0, 2 * sizeof(float), s);
  if (e != cudaSuccess) return (int)e;
  const dim3 grid((W + kLossTileW - 1) / kLossTileW, (H + kLossTileH - 1) / kLossTileH, 3);
  ssim_fwd_kernel<<<grid, kLossThreads, 0, s>>>(a);
  GSR_LAUNCHED();
  ssim_bwd_kernel<<<grid, kLossThreads, 0, s>>>(a);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // extern "C"
