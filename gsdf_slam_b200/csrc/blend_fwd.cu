// Forward alpha-blend: one CTA per 16x16 tile, one pixel per thread.
// Replaces the reference's FORWARD::renderCUDA (forward.cu:317-477).
//
// What is new relative to the reference:
//  * the tile's depth-ordered splats are read as CONTIGUOUS packed records
//    (binning.cu) with cp.async.bulk (TMA 1-D bulk copies) completing on an
//    mbarrier, double-buffered so batch k+1 lands while batch k is blended —
//    no id -> xy/conic dependent gather and no per-pixel global colour read;
//  * no per-32-splat checkpoint stream (4 KB per bucket in the reference): the
//    backward kernel of this implementation replays the tile front-to-back;
//  * a conservative per-splat power threshold skips the exp for pixels far from
//    the splat (cannot change a result, see finalize_kernel);
//  * warps cover 8x4 pixel blocks instead of 16x2 strips (better coherence).
// The per-pixel arithmetic keeps the reference's operation order so n_contrib,
// final_T and colour replay bit-for-bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr_common.cuh"
#include "tma.cuh"

namespace gsr {

constexpr int kFwdBatch = 256;

__global__ void __launch_bounds__(GSR_TILE_PIX)
blend_fwd_kernel(const uint2* __restrict__ ranges, const float4* __restrict__ rec_a,
                 const float4* __restrict__ rec_b, const float4* __restrict__ rec_c, int W, int H,
                 const float* __restrict__ bg, float* __restrict__ out_color,
                 float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                 float* __restrict__ accum_c, uint32_t* __restrict__ max_contrib) {
  __shared__ __align__(128) float4 sA[2][kFwdBatch];
  __shared__ __align__(128) float4 sB[2][kFwdBatch];
  __shared__ __align__(128) float4 sC[2][kFwdBatch];
  __shared__ __align__(8) uint64_t s_full[2];
  __shared__ uint32_t s_max[8];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile_id = blockIdx.y * gridDim.x + blockIdx.x;
  // warp w covers the 8x4 pixel block at ((w&1)*8, (w>>1)*4) of the tile
  const uint32_t px = blockIdx.x * GSR_TILE_X + (warp & 1) * 8 + (lane & 7);
  const uint32_t py = blockIdx.y * GSR_TILE_Y + (warp >> 1) * 4 + (lane >> 3);
  const bool inside = px < (uint32_t)W && py < (uint32_t)H;
  const uint32_t pix_id = (uint32_t)W * py + px;
  const float2 pixf = {(float)px, (float)py};

  const uint2 range = ranges[tile_id];
  const int total = (int)(range.y - range.x);
  const int rounds = (total + kFwdBatch - 1) / kFwdBatch;

  if (tid == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](int round) {
    const int buf = round & 1;
    const int start = range.x + round * kFwdBatch;
    const int cnt = min(kFwdBatch, total - round * kFwdBatch);
    const uint32_t bytes = (uint32_t)cnt * 16u;
    mbar_expect_tx(&s_full[buf], 3u * bytes);
    bulk_g2s(&sA[buf][0], rec_a + start, bytes, &s_full[buf]);
    bulk_g2s(&sB[buf][0], rec_b + start, bytes, &s_full[buf]);
    bulk_g2s(&sC[buf][0], rec_c + start, bytes, &s_full[buf]);
  };
  if (tid == 0 && rounds > 0) issue(0);

  bool done = !inside;
  float T = 1.0f;
  uint32_t contributor = 0;
  uint32_t last_contributor = 0;
  float C[3] = {0.f, 0.f, 0.f};

  for (int i = 0; i < rounds; ++i) {
    const int buf = i & 1;
    mbar_wait(&s_full[buf], (uint32_t)((i >> 1) & 1));
    // all threads are past the previous batch here; stop when every pixel is saturated
    const int num_done = __syncthreads_count(done);
    if (num_done == GSR_TILE_PIX) break;
    if (tid == 0 && i + 1 < rounds) issue(i + 1);

    const int cnt = min(kFwdBatch, total - i * kFwdBatch);
    for (int j = 0; !done && j < cnt; j++) {
      contributor++;
      const float4 A = sA[buf][j];
      const float4 B = sB[buf][j];
      float2 xy = {A.x, A.y};
      float2 d = {xy.x - pixf.x, xy.y - pixf.y};
      float4 con_o = {A.z, A.w, B.x, B.y};
      float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
      if (power > 0.0f) continue;
      if (power < B.z) continue;  // certainly alpha < 1/255

      float alpha = min(0.99f, con_o.w * exp(power));
      if (alpha < 1.0f / 255.0f) continue;
      float test_T = T * (1 - alpha);
      if (test_T < 0.0001f) {
        done = true;
        continue;
      }
      const float4 col = sC[buf][j];
      C[0] += col.x * alpha * T;
      C[1] += col.y * alpha * T;
      C[2] += col.z * alpha * T;
      T = test_T;
      last_contributor = contributor;
    }
  }

  if (inside) {
    const size_t N = (size_t)W * H;
    final_T[pix_id] = T;
    n_contrib[pix_id] = last_contributor;
    out_color[0 * N + pix_id] = C[0] + T * bg[0];
    out_color[1 * N + pix_id] = C[1] + T * bg[1];
    out_color[2 * N + pix_id] = C[2] + T * bg[2];
    accum_c[0 * N + pix_id] = C[0];
    accum_c[1 * N + pix_id] = C[1];
    accum_c[2 * N + pix_id] = C[2];
  }
  // deepest list position any pixel of the tile used (bounds the backward replay)
  const uint32_t wmax = __reduce_max_sync(0xffffffffu, last_contributor);
  if (lane == 0) s_max[warp] = wmax;
  __syncthreads();
  if (tid == 0) {
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) m = max(m, s_max[w]);
    max_contrib[tile_id] = m;
  }
}

int launch_blend_fwd(int W, int H, const uint2* ranges, const float4* rec_a, const float4* rec_b,
                     const float4* rec_c, const float* bg, float* out_color, float* final_T,
                     uint32_t* n_contrib, float* accum_c, uint32_t* max_contrib, cudaStream_t s) {
  dim3 grid((W + GSR_TILE_X - 1) / GSR_TILE_X, (H + GSR_TILE_Y - 1) / GSR_TILE_Y, 1);
  blend_fwd_kernel<<<grid, GSR_TILE_PIX, 0, s>>>(ranges, rec_a, rec_b, rec_c, W, H, bg,
                                                 out_color, final_T,
                                                 n_contrib, accum_c, max_contrib);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
