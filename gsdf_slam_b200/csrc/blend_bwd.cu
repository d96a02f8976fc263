// Backward alpha-blend: per-splat sums of the pixel gradients.  Replaces the reference's
// per-splat-bucket kernel BACKWARD::PerGaussianRenderCUDA (backward.cu:463-653).
//
// Design (B200-first, not the reference's): one 64-thread CTA per tile, four pixels per
// thread (blend_common.cuh).  The tile's list is REPLAYED FRONT-TO-BACK exactly like the
// forward, so T and the accumulated-colour prefix are recomputed in registers (bit-identical
// to the forward) and no checkpoint stream exists.  For each (splat, pixel) pair that blends,
// the same quantities as the reference (backward.cu:597-637) are formed — alpha, T,
// dL/dalpha — but instead of nine per-pair gradient terms the kernel accumulates
//     w = G * dL/dalpha   and its moments   w*dx, w*dy, w*dx^2, w*dx*dy, w*dy^2
// plus the three colour sums: the gradients of the 2D mean, conic and opacity are linear
// in those moments with per-splat coefficients, so that conversion happens ONCE per
// Gaussian in preprocess_bwd.cu.  A thread first adds up its four pixels, the warp then
// combines its 32 lanes with a 14-shuffle transposed reduction (only when some pixel of
// the warp blends the splat), and nine lanes issue one global reduction each.
// Splat records arrive through double-buffered TMA bulk copies.
#include <cuda_runtime.h>
#include <stdint.h>

#include "blend_common.cuh"
#include "gsr_common.cuh"
#include "tma.cuh"

namespace gsr {

// Sum nine per-lane values over the warp with 14 shuffles.  On return lane L with
// (L & 3) == 0 holds the total of v[L >> 2] in `out8`; every lane holds the total of v[8]
// in `out9`.
__device__ __forceinline__ void warp_reduce9(const float (&v)[9], int lane, float& out8,
                                             float& out9) {
  const unsigned full = 0xffffffffu;
  float w[4];
  const bool u16 = (lane & 16) != 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = u16 ? v[k] : v[k + 4];
    const float keep = u16 ? v[k + 4] : v[k];
    w[k] = keep + __shfl_xor_sync(full, send, 16);
  }
  float x[2];
  const bool u8 = (lane & 8) != 0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = u8 ? w[k] : w[k + 2];
    const float keep = u8 ? w[k + 2] : w[k];
    x[k] = keep + __shfl_xor_sync(full, send, 8);
  }
  const bool u4 = (lane & 4) != 0;
  float y = (u4 ? x[1] : x[0]) + __shfl_xor_sync(full, u4 ? x[0] : x[1], 4);
  y += __shfl_xor_sync(full, y, 2);
  y += __shfl_xor_sync(full, y, 1);
  out8 = y;  // component index = 4*bit4 + 2*bit3 + bit2 = lane >> 2
  float z = v[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(full, z, o);
  out9 = z;
}

// Same result through shared memory: every lane stores its nine values (conflict-free rows of
// 36 floats), then four lanes per component add eight of them each (two LDS.128) and finish
// with two shuffles; the ninth component takes a plain shuffle tree.  ~35 instructions instead
// of ~63, which matters because this runs once per (splat, half tile).
constexpr int kRedStride = 36;
__device__ __forceinline__ void warp_reduce9_smem(const float (&v)[9], int lane, float* s_red /* [8][36] */,
                                                  float& out8, float& out9) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int k = 0; k < 8; ++k) s_red[k * kRedStride + lane] = v[k];
  __syncwarp();
  const float4* row = reinterpret_cast<const float4*>(s_red + (lane >> 2) * kRedStride + (lane & 3) * 8);
  const float4 a = row[0];
  const float4 b = row[1];
  float y = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
  y += __shfl_xor_sync(full, y, 1);
  y += __shfl_xor_sync(full, y, 2);
  out8 = y;  // total of v[lane >> 2]
  float z = v[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(full, z, o);
  out9 = z;
  __syncwarp();  // rows are rewritten by the next reduction
}

__global__ void __launch_bounds__(kBlendThreads, GSR_BLEND_BWD_MINBLOCKS)
blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ max_contrib,
                 const float4* __restrict__ rec, int W, int H, const float* __restrict__ bg,
                 const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib,
                 const float* __restrict__ accum_c, const float* __restrict__ dL_dpix,
                 float* __restrict__ acc /* [P][12] */) {
  __shared__ __align__(128) float4 sR[2][3 * kBlendBatch];  // 48-byte records, one bulk copy per batch
  __shared__ __align__(8) uint64_t s_full[2];
  __shared__ __align__(16) float s_red[kBlendThreads / 32][8 * kRedStride];

  const int lane = threadIdx.x;
  // 1-D grid in longest-list-first order (binning.cu: tile_order_kernel); two CTAs per tile
  const uint32_t tile_id = tile_order[blockIdx.x >> 1];
  const int warp = blockIdx.x & 1;  // upper / lower half of the tile
  const uint32_t grid_x = (uint32_t)(W + GSR_TILE_X - 1) / GSR_TILE_X;
  const uint32_t tile_x = tile_id % grid_x, tile_y = tile_id / grid_x;
  const size_t N = (size_t)W * H;

  const uint2 range = ranges[tile_id];
  // nothing behind the deepest contributor of the tile receives gradient
  int total = min((int)(range.y - range.x), (int)max_contrib[tile_id]);

  if (lane == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    fence_mbar_init();
  }
  __syncwarp();

  auto issue = [&](int round) {
    const int buf = round & 1;
    const int start = range.x + round * kBlendBatch;
    const int cnt = min(kBlendBatch, total - round * kBlendBatch);
    const uint32_t bytes = (uint32_t)cnt * 48u;
    mbar_expect_tx(&s_full[buf], bytes);
    bulk_g2s(&sR[buf][0], rec + 3 * (size_t)start, bytes, &s_full[buf]);
  };


  // per-pixel state of the four slots.  The reference tracks, per channel, ar = -(colour still
  // to be blended behind the current splat) and forms
  //   dL/dalpha = sum_ch (c_ch*T + ar_ch/(1-alpha)) * dL_ch - T_final/(1-alpha) * (bg . dL)
  // (backward.cu:609-618).  ar only ever meets dL in a dot product, so one scalar per pixel
  // suffices:  S = sum_ch ar_ch*dL_ch - T_final*(bg . dL);  each blended splat adds
  // T*alpha*(c . dL) to it, and  dL/dalpha = T*(c . dL) + S/(1-alpha).
  float pxf[kSlots], pyf[kSlots], T[kSlots], S[kSlots];
  uint32_t last[kSlots];
  float dL[kSlots][3];
  const float bgc[3] = {bg[0], bg[1], bg[2]};
  uint32_t my_last = 0;
#pragma unroll
  for (int q = 0; q < kSlots; ++q) {
    const uint32_t px = slot_px(tile_x, lane, q);
    const uint32_t py = slot_py(tile_y, warp, lane, q);
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint32_t pix_id = (uint32_t)W * py + px;
    pxf[q] = (float)px;
    pyf[q] = (float)py;
    T[q] = 1.0f;
    S[q] = 0.f;
    last[q] = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) dL[q][ch] = 0.f;
    if (inside) {
      const float T_final = final_T[pix_id];
      last[q] = n_contrib[pix_id];
      float bg_dot = 0.f, ar_dot = 0.f;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        dL[q][ch] = dL_dpix[ch * N + pix_id];
        ar_dot += accum_c[ch * N + pix_id] * dL[q][ch];
        bg_dot += bgc[ch] * dL[q][ch];
      }
      S[q] = -ar_dot - T_final * bg_dot;
    }
    my_last = max(my_last, last[q]);
  }
  // warp-uniform bound: the deepest contributor among this half tile's 128 pixels; nothing
  // behind it is even loaded
  const uint32_t warp_last = __reduce_max_sync(0xffffffffu, my_last);
  total = min(total, (int)warp_last);
  const int rounds = (total + kBlendBatch - 1) / kBlendBatch;
  if (lane == 0 && rounds > 0) issue(0);

  for (int i = 0; i < rounds; ++i) {
    const int buf = i & 1;
    mbar_wait(&s_full[buf], (uint32_t)((i >> 1) & 1));
    __syncwarp();  // previous batch fully consumed by every lane
    if (lane == 0 && i + 1 < rounds) issue(i + 1);

    const int cnt = min(kBlendBatch, total - i * kBlendBatch);
    const uint32_t pos0 = (uint32_t)i * kBlendBatch;
    const int jmax = (int)min((uint32_t)cnt, warp_last > pos0 ? warp_last - pos0 : 0u);

    for (int j = 0; j < jmax; ++j) {
      const float4 A = sR[buf][3 * j];
      const float4 B = sR[buf][3 * j + 1];
      const float4 con_o = {A.z, A.w, B.x, B.y};
      const uint32_t pos = pos0 + j;

      // v = {Mx, My, Mxx, Mxy, Myy, M0, cr, cg, cb}
      float v[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k] = 0.f;
      // gate of all four pixels first: independent chains, one vote for the common miss
      float pw[kSlots];
      bool cand[kSlots];
      bool any_cand = false;
#pragma unroll
      for (int q = 0; q < kSlots; ++q) {
        const float2 d = {A.x - pxf[q & 1], A.y - pyf[q & 2]};
        const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
        pw[q] = power;
        cand[q] = (pos < last[q]) && !(power > 0.0f) && !(power < B.z);
        any_cand = any_cand || cand[q];
      }
      if (!__any_sync(0xffffffffu, any_cand)) continue;  // splat misses this warp's half tile

      // the four exp / alpha chains are evaluated together (4-way ILP, shared constants, no branches) like
      // in the forward kernel; the chain of a pixel that is not a candidate computes a value nobody uses
      float G[kSlots], alpha[kSlots];
      bool live[kSlots];
#pragma unroll
      for (int q = 0; q < kSlots; ++q) {
        G[q] = exp(pw[q]);
        alpha[q] = min(0.99f, con_o.w * G[q]);
        live[q] = cand[q] && !(alpha[q] < 1.0f / 255.0f);
      }
      const bool contributed = live[0] || live[1] || live[2] || live[3];
      if (!__any_sync(0xffffffffu, contributed)) continue;
      {
        const float4 col = sR[buf][3 * j + 2];
#pragma unroll
        for (int q = 0; q < kSlots; ++q) {
          if (!live[q]) continue;
          const float2 d = {A.x - pxf[q & 1], A.y - pyf[q & 2]};
          const float one_m = 1.0f - alpha[q];
          // 1 / (1 - alpha), 1 - alpha in [0.01, 1]: the hardware reciprocal (1 ulp) is enough for a
          // gradient term (parity bar 1e-4 rel-L2) and, unlike __frcp_rn / IEEE division, compiles to a
          // single MUFU.RCP without the range check + fix-up + slow-path call (10 instructions per pair)
          float inv;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(one_m));
          const float Ta = T[q] * alpha[q];
          const float cdl = col.x * dL[q][0] + col.y * dL[q][1] + col.z * dL[q][2];
          v[6] += Ta * dL[q][0];
          v[7] += Ta * dL[q][1];
          v[8] += Ta * dL[q][2];
          S[q] += Ta * cdl;
          const float dL_dalpha = T[q] * cdl + inv * S[q];
          T[q] *= one_m;

          const float w = G[q] * dL_dalpha;
          const float wdx = w * d.x;
          const float wdy = w * d.y;
          v[0] += wdx;
          v[1] += wdy;
          v[2] += wdx * d.x;
          v[3] += wdx * d.y;
          v[4] += wdy * d.y;
          v[5] += w;
        }
      }

      float s8, s9;
      warp_reduce9_smem(v, lane, s_red[0], s8, s9);
      const uint32_t g = __float_as_uint(B.w);
      if ((lane & 3) == 0) {
        if (s8 != 0.f) atomicAdd(&acc[(size_t)g * 12 + (lane >> 2)], s8);
      } else if (lane == 1) {
        if (s9 != 0.f) atomicAdd(&acc[(size_t)g * 12 + 8], s9);
      }
    }
  }
}

int launch_blend_bwd(int W, int H, const uint2* ranges, const uint32_t* tile_order,
                     const uint32_t* max_contrib,
                     const float4* rec, const float* bg,
                     const float* final_T, const uint32_t* n_contrib, const float* accum_c,
                     const float* dL_dpix, float* acc, cudaStream_t s) {
  const unsigned grid = (unsigned)(((W + GSR_TILE_X - 1) / GSR_TILE_X) * ((H + GSR_TILE_Y - 1) / GSR_TILE_Y));
  blend_bwd_kernel<<<2 * grid, kBlendThreads, 0, s>>>(ranges, tile_order, max_contrib, rec, W, H, bg,
                                                  final_T, n_contrib, accum_c, dL_dpix, acc);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
