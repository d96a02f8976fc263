// Backward alpha-blend: gradients of the loss w.r.t. each splat's 2D mean, conic,
// opacity and colour.  Replaces the reference's per-splat-bucket kernel
// BACKWARD::PerGaussianRenderCUDA (backward.cu:463-653).
//
// Design (B200-first, not the reference's): one CTA per tile, one pixel per
// thread, the tile's list is REPLAYED FRONT-TO-BACK exactly like the forward, so
// T and the accumulated colour prefix are recomputed in registers and no
// checkpoint stream is needed.  With the suffix colour written as
// (prefix - final), the per-pair gradient terms are the same expressions the
// reference evaluates (backward.cu:597-637).  A pair only costs the gradient
// maths if some pixel of the warp's 8x4 block actually blends the splat; the nine
// per-splat partial sums are then pre-reduced with warp shuffles, accumulated per
// CTA in shared memory, and flushed with ONE global reduction per (splat, tile)
// and component — the same atomic count as the reference, issued coalesced.
// Splat records arrive through double-buffered TMA bulk copies.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr_common.cuh"
#include "tma.cuh"

namespace gsr {

constexpr int kBwdBatch = 256;
constexpr int kAccStride = 9;

__global__ void __launch_bounds__(GSR_TILE_PIX)
blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ max_contrib,
                 const float4* __restrict__ rec_a, const float4* __restrict__ rec_b,
                 const float4* __restrict__ rec_c, int W, int H, const float* __restrict__ bg,
                 const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                 const float* __restrict__ accum_c, const float* __restrict__ dL_dpix,
                 float* __restrict__ acc /* [P][12] */) {
  __shared__ __align__(128) float4 sA[2][kBwdBatch];
  __shared__ __align__(128) float4 sB[2][kBwdBatch];
  __shared__ __align__(128) float4 sC[2][kBwdBatch];
  __shared__ float s_acc[kBwdBatch * kAccStride];
  __shared__ __align__(8) uint64_t s_full[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile_id = blockIdx.y * gridDim.x + blockIdx.x;
  const uint32_t px = blockIdx.x * GSR_TILE_X + (warp & 1) * 8 + (lane & 7);
  const uint32_t py = blockIdx.y * GSR_TILE_Y + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < (uint32_t)W && py < (uint32_t)H;
  const uint32_t pix_id = (uint32_t)W * py + px;
  const float2 pixf = {(float)px, (float)py};
  const size_t N = (size_t)W * H;

  const uint2 range = ranges[tile_id];
  // nothing behind the deepest contributor of the tile receives gradient
  const int total = min((int)(range.y - range.x), (int)max_contrib[tile_id]);
  const int rounds = (total + kBwdBatch - 1) / kBwdBatch;

  if (tid == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < kBwdBatch * kAccStride; i += GSR_TILE_PIX) s_acc[i] = 0.f;
  __syncthreads();

  auto issue = [&](int round) {
    const int buf = round & 1;
    const int start = range.x + round * kBwdBatch;
    const int cnt = min(kBwdBatch, total - round * kBwdBatch);
    const uint32_t bytes = (uint32_t)cnt * 16u;
    mbar_expect_tx(&s_full[buf], 3u * bytes);
    bulk_g2s(&sA[buf][0], rec_a + start, bytes, &s_full[buf]);
    bulk_g2s(&sB[buf][0], rec_b + start, bytes, &s_full[buf]);
    bulk_g2s(&sC[buf][0], rec_c + start, bytes, &s_full[buf]);
  };
  if (tid == 0 && rounds > 0) issue(0);

  // per-pixel state
  float T = 1.0f;
  float T_final = 0.f;
  uint32_t last = 0;
  float ar[3] = {0.f, 0.f, 0.f};
  float dL[3] = {0.f, 0.f, 0.f};
  const float bgc[3] = {bg[0], bg[1], bg[2]};
  if (inside) {
    T_final = final_T[pix_id];
    last = n_contrib[pix_id];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      ar[ch] = -accum_c[ch * N + pix_id];  // minus the colour still to be blended
      dL[ch] = dL_dpix[ch * N + pix_id];
    }
  }
  float bg_dot_dpixel = 0.0f;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) bg_dot_dpixel += bgc[ch] * dL[ch];
  const float ddelx_dx = 0.5f * W;
  const float ddely_dy = 0.5f * H;

  for (int i = 0; i < rounds; ++i) {
    const int buf = i & 1;
    mbar_wait(&s_full[buf], (uint32_t)((i >> 1) & 1));
    __syncthreads();  // previous batch fully consumed and flushed by every thread
    if (tid == 0 && i + 1 < rounds) issue(i + 1);

    const int cnt = min(kBwdBatch, total - i * kBwdBatch);
    const uint32_t pos0 = (uint32_t)i * kBwdBatch;
    // warp-uniform bound: the deepest contributor among this warp's pixels
    const uint32_t warp_last = __reduce_max_sync(0xffffffffu, last);
    const int jmax = (int)min((uint32_t)cnt, warp_last > pos0 ? warp_last - pos0 : 0u);

    for (int j = 0; j < jmax; ++j) {
      const float4 A = sA[buf][j];
      const float4 B = sB[buf][j];
      const float2 d = {A.x - pixf.x, A.y - pixf.y};
      const float4 con_o = {A.z, A.w, B.x, B.y};
      const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
      bool contributes = (pos0 + j < last) && !(power > 0.0f) && !(power < B.z);
      float G = 0.f, alpha = 0.f;
      if (contributes) {
        G = exp(power);
        alpha = min(0.99f, con_o.w * G);
        contributes = !(alpha < 1.0f / 255.0f);
      }
      if (!__any_sync(0xffffffffu, contributes)) continue;

      float v[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k] = 0.f;
      if (contributes) {
        const float4 col = sC[buf][j];
        const float c[3] = {col.x, col.y, col.z};
        const float dchannel_dcolor = alpha * T;
        float dL_dalpha = 0.0f;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          ar[ch] += T * alpha * c[ch];
          const float dL_dchannel = dL[ch];
          v[6 + ch] = dchannel_dcolor * dL_dchannel;
          dL_dalpha += ((c[ch] * T) - (1.0f / (1.0f - alpha)) * (-ar[ch])) * dL_dchannel;
        }
        dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot_dpixel;
        T *= (1.0f - alpha);

        const float dL_dG = con_o.w * dL_dalpha;
        const float gdx = G * d.x;
        const float gdy = G * d.y;
        const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
        const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;
        v[0] = dL_dG * dG_ddelx * ddelx_dx;
        v[1] = dL_dG * dG_ddely * ddely_dy;
        v[2] = -0.5f * gdx * d.x * dL_dG;
        v[3] = -0.5f * gdx * d.y * dL_dG;
        v[4] = -0.5f * gdy * d.y * dL_dG;
        v[5] = G * dL_dalpha;
      }
      // warp pre-reduction of the nine partial sums
#pragma unroll
      for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
      }
      if (lane < 9) {
        float mine = v[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) mine = (lane == k) ? v[k] : mine;
        atomicAdd(&s_acc[j * kAccStride + lane], mine);
      }
    }
    __syncthreads();
    // flush this batch: one global reduction per (splat, tile) and component
    for (int e = tid; e < cnt * kAccStride; e += GSR_TILE_PIX) {
      const float val = s_acc[e];
      if (val != 0.f) {
        const int j = e / kAccStride, k = e - j * kAccStride;
        const uint32_t g = __float_as_uint(sB[buf][j].w);
        atomicAdd(&acc[(size_t)g * 12 + k], val);
        s_acc[e] = 0.f;
      }
    }
  }
}

int launch_blend_bwd(int W, int H, const uint2* ranges, const uint32_t* max_contrib,
                     const float4* rec_a, const float4* rec_b, const float4* rec_c, const float* bg,
                     const float* final_T, const uint32_t* n_contrib, const float* accum_c,
                     const float* dL_dpix, float* acc, cudaStream_t s) {
  dim3 grid((W + GSR_TILE_X - 1) / GSR_TILE_X, (H + GSR_TILE_Y - 1) / GSR_TILE_Y, 1);
  blend_bwd_kernel<<<grid, GSR_TILE_PIX, 0, s>>>(ranges, max_contrib, rec_a, rec_b, rec_c, W, H, bg,
                                                 final_T, n_contrib, accum_c, dL_dpix, acc);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
