// Forward per-Gaussian pass: frustum cull, projection, 3D->2D covariance, conic, 3-sigma radius, SH->RGB, EXACT
// opacity-aware tile counts (per Gaussian and per tile), and the compacted list of rendered Gaussians.  Replaces the
// reference's preprocessCUDA (forward.cu:197-312) plus the counting half of duplicateWithKeys
// (rasterizer_impl.cu:139-212): by running the tile test here, only tiles that will really be rasterised are
// counted, so no padding keys are ever generated.
//
// Three launches (round 2; measured on B200 against the single-kernel version of round 1, whole stage):
//   cfg2 500 k Gaussians 59 -> 54 us, cfg5 3 M 227 -> 181 us, cfg4 2 M (all visible) 725 -> 686 us,
//   cfg1 100 k (all visible) 52 -> 70 us (two more launches for nothing to cull).
//  1. precull_kernel      every Gaussian: z test + conservative screen-space bound -> candidate list (a mapping view
//                         keeps ~ 1/6 of the map; 40 B read per Gaussian as 16-byte loads)
//  2. preprocess_kernel   the reference's arithmetic, one candidate per thread, dense warps; SH rows by one TMA bulk
//                         copy per lane; tile walk 32 (Gaussian, tile) pairs per step, per-Gaussian counts from the
//                         hit ballots
//  3. count_large_kernel  rectangles of more than kCountQueueTiles tiles, one warp per 128-tile chunk
// What the profile of round 1's kernel showed (ncu source page, 2.9 k samples): it was neither bandwidth- nor
// issue-bound (42 % issue utilisation) but waited for its LONGEST WARP — the one holding a near-camera splat whose
// rectangle covers 1900 tiles walks 60 steps of 32 tiles while the other warps need one or two (cfg2; 2967 tiles
// at cfg5).  Hence stage 3.  Things measured and dropped: fetching scale / rotation only for the
// z-cull survivors with barrier-separated phases inside one CTA (111 vs 59 us); appending the (Gaussian, tile)
// hits to a list for the emit stage (+22 us here, -16 us there); 16 or 8 candidates per warp; 4-byte loads at a
// 12-byte stride in stage 1 (45 vs 42 us at 3 M); a per-warp atomic on the list counter in stage 1 (81 us at 3 M:
// 94 k same-address atomics serialise in L2).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "gsr_common.cuh"
#include "gsr_params.cuh"
#include "tma.cuh"

namespace gsr {


constexpr float kSH_C0 = 0.28209479177387814f;
constexpr float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                -0.5900435899266435f};

// world-space covariance from scale + (caller-normalised) quaternion
// (reference computeCov3D, forward.cu:154-193): Sigma = (S R)^T (S R)
__device__ __forceinline__ void cov3d_from_scale_rot(const float3 scale, float mod, const float4 q,
                                                     float* cov3D) {
  M3 S = m3_make(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
  S[0][0] = mod * scale.x;
  S[1][1] = mod * scale.y;
  S[2][2] = mod * scale.z;
  const float r = q.x, x = q.y, y = q.z, z = q.w;
  M3 R = m3_make(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                 2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                 2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
  M3 Mm = m3_mul(S, R);
  M3 Sigma = m3_mul(m3_transpose(Mm), Mm);
  cov3D[0] = Sigma[0][0];
  cov3D[1] = Sigma[0][1];
  cov3D[2] = Sigma[0][2];
  cov3D[3] = Sigma[1][1];
  cov3D[4] = Sigma[1][2];
  cov3D[5] = Sigma[2][2];
}

// EWA projection of the covariance (reference computeCov2D, forward.cu:104-149)
__device__ __forceinline__ float3 cov2d_project(const float3& mean, float focal_x, float focal_y,
                                                float tan_fovx, float tan_fovy, const float* cov3D,
                                                const float* view) {
  float3 t = xform_4x3(mean, view);
  const float limx = 1.3f * tan_fovx;
  const float limy = 1.3f * tan_fovy;
  const float txtz = t.x / t.z;
  const float tytz = t.y / t.z;
  t.x = min(limx, max(-limx, txtz)) * t.z;
  t.y = min(limy, max(-limy, tytz)) * t.z;

  M3 J = m3_make(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z), 0.0f, focal_y / t.z,
                 -(focal_y * t.y) / (t.z * t.z), 0, 0, 0);
  M3 Wm = m3_make(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
  M3 T = m3_mul(Wm, J);
  M3 Vrk = m3_make(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4],
                   cov3D[5]);
  M3 cov = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);
  cov[0][0] += 0.3f;
  cov[1][1] += 0.3f;
  return {float(cov[0][0]), float(cov[0][1]), float(cov[1][1])};
}

// One colour channel of the SH expansion (reference computeColorFromSH,
// forward.cu:32-101; there it is evaluated on 3-vectors, i.e. per channel with
// exactly this association).  sh[k] = coefficient k (1..15 -> index k-1) of this channel.
__device__ __forceinline__ float sh_channel(int deg, float x, float y, float z, float dc0,
                                            const float* sh, int stride) {
  float result = kSH_C0 * dc0;
  if (deg > 0) {
    result = result - kSH_C1 * y * sh[0 * stride] + kSH_C1 * z * sh[1 * stride] -
             kSH_C1 * x * sh[2 * stride];
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z;
      float xy = x * y, yz = y * z, xz = x * z;
      result = result + kSH_C2[0] * xy * sh[3 * stride] + kSH_C2[1] * yz * sh[4 * stride] +
               kSH_C2[2] * (2.0f * zz - xx - yy) * sh[5 * stride] +
               kSH_C2[3] * xz * sh[6 * stride] + kSH_C2[4] * (xx - yy) * sh[7 * stride];
      if (deg > 2) {
        result = result + kSH_C3[0] * y * (3.0f * xx - yy) * sh[8 * stride] +
                 kSH_C3[1] * xy * z * sh[9 * stride] +
                 kSH_C3[2] * y * (4.0f * zz - xx - yy) * sh[10 * stride] +
                 kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[11 * stride] +
                 kSH_C3[4] * x * (4.0f * zz - xx - yy) * sh[12 * stride] +
                 kSH_C3[5] * z * (xx - yy) * sh[13 * stride] +
                 kSH_C3[6] * x * (xx - 3.0f * yy) * sh[14 * stride];
      }
    }
  }
  result += 0.5f;
  return result;
}

constexpr int kPreThreads = 128;
// SH rows of the rendered Gaussians of a warp are staged kShSlots at a time, each by ONE bulk copy (TMA, SASS
// UBLKCP) that its own lane issues: the 16-byte aligned window around the 12*M-byte row (rows start on 4-byte
// boundaries only), completing on the warp's mbarrier
constexpr int kShSlots = 16;
constexpr int kShRowBytes = 208;   // 12 bytes of alignment slack + 180 bytes, rounded up to 16

// ---- stage 1: conservative cull -----------------------------------------------------------------------
// One thread per Gaussian reads the xyz and scale rows (24 B; + the 16 B rotation row when the caller's
// quaternions are not normalised in here) and decides whether the Gaussian CAN be rendered: view-space
// z > 0.2 (the reference's only early test, forward.cu:216 / auxiliary.h:160-194 — same expression) and a
// screen-space disc whose radius is an upper bound of the reference's 3-sigma radius,
//     lambda_max(cov2D) <= |W|^2 |J|_F^2 lambda_max(Sigma) + 0.3,  lambda1 <= lambda_max + sqrt(0.1),
// (2 % + 2 px of slack for rounding) overlapping the image.  A Gaussian this test rejects has an empty tile
// rectangle in the reference (getRect is monotone in the radius), i.e. radii = 0 and nothing else written —
// which is what this kernel writes for it.  Everything else (a mapping view keeps ~ 1/5 of the map) is
// appended to the candidate list, so the expensive stage below runs on dense warps.
constexpr int kCullThreads = 256;
constexpr int kCullRows = 4;   // CONSECUTIVE rows per thread: 48-byte xyz / scale groups and 64 bytes of quaternions as
                               // 16-byte loads (4-byte loads at a 12-byte stride kept this stage at a third of the
                               // HBM rate)
__device__ __forceinline__ void load_rows4(const float* __restrict__ base, int i0, int P, bool vec, float (&out)[12]) {
  if (vec && i0 + 3 < P) {
    const float4* p = reinterpret_cast<const float4*>(base + 3 * (size_t)i0);
    const float4 x0 = p[0], x1 = p[1], x2 = p[2];
    out[0] = x0.x; out[1] = x0.y; out[2] = x0.z; out[3] = x0.w; out[4] = x1.x; out[5] = x1.y;
    out[6] = x1.z; out[7] = x1.w; out[8] = x2.x; out[9] = x2.y; out[10] = x2.z; out[11] = x2.w;
  } else {
#pragma unroll
    for (int k = 0; k < 12; ++k) out[k] = (i0 + k / 3 < P) ? base[3 * (size_t)i0 + k] : 0.f;
  }
}

__global__ void __launch_bounds__(kCullThreads, 4) precull_kernel(const PreprocessParams a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i0 = (blockIdx.x * kCullThreads + threadIdx.x) * kCullRows;
  // ---- all input rows of this thread first ----
  float xyz[12];
  load_rows4(a.means3D, i0, a.P, a.vec_ok != 0, xyz);
  float var3[kCullRows];   // upper bound of the largest eigenvalue of the world-space covariance
  if (a.cov3D_precomp != nullptr) {
#pragma unroll
    for (int r = 0; r < kCullRows; ++r) {
      var3[r] = 0.f;
      if (i0 + r < a.P) {
        const float* c = a.cov3D_precomp + (size_t)(i0 + r) * 6;
        const float fro2 = c[0] * c[0] + c[3] * c[3] + c[5] * c[5] + 2.f * (c[1] * c[1] + c[2] * c[2] + c[4] * c[4]);
        var3[r] = 1.001f * fro2 * rsqrtf(fmaxf(fro2, 1e-37f));   // Frobenius >= spectral norm (approximate sqrt + slack)
      }
    }
  } else {
    float sc[12];
    load_rows4(a.scales, i0, a.P, a.vec_ok != 0, sc);
#pragma unroll
    for (int r = 0; r < kCullRows; ++r) {
      float smax;
      if (a.activate) smax = 1.001f * __expf(fmaxf(sc[3 * r], fmaxf(sc[3 * r + 1], sc[3 * r + 2])));   // exp is monotone
      else smax = fmaxf(fabsf(sc[3 * r]), fmaxf(fabsf(sc[3 * r + 1]), fabsf(sc[3 * r + 2])));
      smax *= a.scale_modifier;
      var3[r] = smax * smax;
    }
    if (!a.activate) {   // |R|_2 of the matrix built from a quaternion of norm n: (1 - n^2) I + n^2 Rot
#pragma unroll
      for (int r = 0; r < kCullRows; ++r) {
        if (i0 + r < a.P) {
          const float4 q = reinterpret_cast<const float4*>(a.rotations)[i0 + r];
          const float n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
          const float rn = fabsf(1.0f - n2) + n2;
          var3[r] *= rn * rn;
        }
      }
    }
  }
  // |W|_2^2: 1 for a rigid view matrix, else bounded by the Frobenius norm
  const float* v = a.view;
  const float g00 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], g11 = v[4] * v[4] + v[5] * v[5] + v[6] * v[6],
              g22 = v[8] * v[8] + v[9] * v[9] + v[10] * v[10];
  const float g01 = v[0] * v[4] + v[1] * v[5] + v[2] * v[6], g02 = v[0] * v[8] + v[1] * v[9] + v[2] * v[10],
              g12 = v[4] * v[8] + v[5] * v[9] + v[6] * v[10];
  const bool rigid = fabsf(g00 - 1.f) < 1e-3f && fabsf(g11 - 1.f) < 1e-3f && fabsf(g22 - 1.f) < 1e-3f &&
                     fabsf(g01) < 1e-3f && fabsf(g02) < 1e-3f && fabsf(g12) < 1e-3f;
  const float wn = rigid ? 1.01f : 1.01f * (g00 + g11 + g22);
  const float limx = 1.3f * a.tan_fovx, limy = 1.3f * a.tan_fovy;

  unsigned keep = 0;   // bit r: row i0 + r is a candidate
#pragma unroll
  for (int r = 0; r < kCullRows; ++r) {
    const int idx = i0 + r;
    if (idx < a.P) {
      const float3 p_orig = {xyz[3 * r], xyz[3 * r + 1], xyz[3 * r + 2]};
      const float3 p_view = xform_4x3(p_orig, a.view);
      bool cand = false;
      if (p_view.z <= 0.2f) {
        if (a.prefiltered) {
          printf("Point is filtered although prefiltered is set. This shouldn't happen!");
          __trap();
        }
      } else {
        // approximate reciprocals / square roots (MUFU, ~1e-6 relative): the 2 % + 2 px of slack cover them
        const float itz = __fdividef(1.0f, p_view.z);
        const float txc = fminf(limx, fmaxf(-limx, p_view.x * itz)), tyc = fminf(limy, fmaxf(-limy, p_view.y * itz));
        const float jf2 = (a.focal_x * a.focal_x * (1.f + txc * txc) + a.focal_y * a.focal_y * (1.f + tyc * tyc)) * itz * itz;
        const float lam = var3[r] * jf2 * wn + 0.62f;
        const float rb = 3.f * 1.02f * lam * rsqrtf(lam) + 2.f;
        const float4 p_hom = xform_4x4(p_orig, a.proj);
        const float p_w = __fdividef(1.0f, p_hom.w + 0.0000001f);
        const float px = ((p_hom.x * p_w + 1.0f) * a.W - 1.0f) * 0.5f, py = ((p_hom.y * p_w + 1.0f) * a.H - 1.0f) * 0.5f;
        const bool off = (px + rb < 0.f) || (px - rb > (float)(a.grid_x * GSR_TILE_X)) || (py + rb < 0.f) ||
                         (py - rb > (float)(a.grid_y * GSR_TILE_Y));   // NaN / Inf anywhere: not culled
        cand = !off;
      }
      if (cand) keep |= 1u << r;
    }
  }
  // what the reference leaves behind for a culled Gaussian
  if (i0 + 3 < a.P && keep == 0 && (reinterpret_cast<uintptr_t>(a.radii) & 15) == 0) {
    reinterpret_cast<int4*>(a.radii)[i0 >> 2] = make_int4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(a.tiles_touched)[i0 >> 2] = make_uint4(0, 0, 0, 0);
  } else {
#pragma unroll
    for (int r = 0; r < kCullRows; ++r) {
      if (i0 + r < a.P && !((keep >> r) & 1u)) {
        a.radii[i0 + r] = 0;
        a.tiles_touched[i0 + r] = 0;
      }
    }
  }
  // append: ONE atomic per CTA (a per-warp atomic on the single counter serialises in L2: 94 k of them cost
  // 80 us at 3 M Gaussians); the CTA's candidates stay in index order
  __shared__ uint32_t s_off[kCullThreads / 32];
  __shared__ uint32_t s_base;
  const uint32_t mine = (uint32_t)__popc(keep);
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_off[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t c = lane < kCullThreads / 32 ? s_off[lane] : 0u;
    uint32_t wincl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wincl, o);
      if (lane >= o) wincl += t;
    }
    if (lane < kCullThreads / 32) s_off[lane] = wincl - c;
    if (lane == 31 && wincl) s_base = atomicAdd(a.num_cand, wincl);
  }
  __syncthreads();
  if (keep) {
    uint32_t pos = s_base + s_off[warp] + incl - mine;
#pragma unroll
    for (int r = 0; r < kCullRows; ++r)
      if ((keep >> r) & 1u) a.cand[pos++] = (uint32_t)(i0 + r);
  }
}

// ---- stage 2: the reference's preprocess on the candidates, one per thread -------------------------------
#ifndef GSR_PRE_MINBLOCKS
#define GSR_PRE_MINBLOCKS 8
#endif
// rectangles of more tiles than this leave the warp's walk for the work queue (16 steps of 32 tiles; the
// chunks of the queue are kQueueRectTiles tiles, as in the emit stage)
#ifndef GSR_COUNT_QUEUE_TILES
#define GSR_COUNT_QUEUE_TILES 512
#endif
constexpr int kCountQueueTiles = GSR_COUNT_QUEUE_TILES;
// candidates per warp: fewer than 32 spreads a short candidate list over more warps (latency hiding when the
// list is shorter than one wave of full warps); the tile walk below is warp-cooperative either way
#ifndef GSR_PRE_CAND_PER_WARP
#define GSR_PRE_CAND_PER_WARP 32
#endif
constexpr int kCandPerWarp = GSR_PRE_CAND_PER_WARP;
__global__ void __launch_bounds__(kPreThreads, GSR_PRE_MINBLOCKS) preprocess_kernel(const PreprocessParams a) {
  __shared__ __align__(16) unsigned char s_sh[kPreThreads / 32][kShSlots * kShRowBytes];  // per-warp SH rows
  __shared__ __align__(8) uint64_t s_bar[kPreThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_cand = a.cand != nullptr ? *a.num_cand : (uint32_t)a.P;   // no list: every Gaussian (cull skipped)
  uint32_t sh_phase = 0;   // parity of the warp's mbarrier
  if (lane == 0) {
    mbar_init(&s_bar[warp], 1);
    fence_mbar_init();
  }
  __syncwarp();
  // warps stride over the candidate list; everything below synchronises per warp only
  // chunk c of kCandPerWarp candidates -> CTA c % grid: a list shorter than one wave still loads every SM evenly
  const uint32_t n_chunks = (n_cand + kCandPerWarp - 1) / kCandPerWarp;
  for (uint32_t chunk = blockIdx.x + gridDim.x * warp; chunk < n_chunks; chunk += gridDim.x * (kPreThreads / 32)) {
  const uint32_t w0 = chunk * (uint32_t)kCandPerWarp;
  const bool in_range = lane < kCandPerWarp && w0 + lane < n_cand;
  const int idx = in_range ? (a.cand != nullptr ? (int)a.cand[w0 + lane] : (int)(w0 + lane)) : 0;

  // ---- this thread's input rows, all loads in flight at once --------------------------------------------
  float3 p_orig = {0.f, 0.f, 0.f}, sc = {0.f, 0.f, 0.f}, dc = {0.f, 0.f, 0.f};
  float4 q = {0.f, 0.f, 0.f, 0.f};
  if (in_range) {
    p_orig = {a.means3D[3 * (size_t)idx], a.means3D[3 * (size_t)idx + 1], a.means3D[3 * (size_t)idx + 2]};
    if (a.scales != nullptr) {
      sc = {a.scales[3 * (size_t)idx], a.scales[3 * (size_t)idx + 1], a.scales[3 * (size_t)idx + 2]};
      q = reinterpret_cast<const float4*>(a.rotations)[idx];
    }
    if (a.colors_precomp == nullptr) dc = {a.dc[3 * (size_t)idx], a.dc[3 * (size_t)idx + 1], a.dc[3 * (size_t)idx + 2]};
  }
  // raw-parameter entry point (row f-2): activate first, so the geometry code below is the same
  // instruction stream for both entry points
  if (a.activate && a.scales != nullptr && in_range) {
    sc = act_exp3(sc);
    q = act_normalize4(q);
  }

  // ---- per-Gaussian geometry ---------------------------------------------------
  bool visible = false;
  float3 p_view = {0.f, 0.f, 0.f};
  float my_radius = 0.f;
  float2 point_image = {0.f, 0.f};
  float3 conic = {0.f, 0.f, 0.f};
  uint2 rect_min = {0, 0}, rect_max = {0, 0};
  float opacity = 0.f;

  if (in_range) {
    p_view = xform_4x3(p_orig, a.view);
    if (p_view.z <= 0.2f) {   // only without stage 1 (its candidates all pass)
      if (a.prefiltered) {
        printf("Point is filtered although prefiltered is set. This shouldn't happen!");
        __trap();
      }
    } else {
      float4 p_hom = xform_4x4(p_orig, a.proj);
      float p_w = 1.0f / (p_hom.w + 0.0000001f);
      float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};

      float cov3D[6];
      if (a.cov3D_precomp != nullptr) {
#pragma unroll
        for (int k = 0; k < 6; ++k) cov3D[k] = a.cov3D_precomp[(size_t)idx * 6 + k];
      } else {
        cov3d_from_scale_rot(sc, a.scale_modifier, q, cov3D);
      }
      float3 cov = cov2d_project(p_orig, a.focal_x, a.focal_y, a.tan_fovx, a.tan_fovy, cov3D, a.view);

      float det = (cov.x * cov.z - cov.y * cov.y);
      if (det != 0.0f) {
        float det_inv = 1.f / det;
        conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};
        float mid = 0.5f * (cov.x + cov.z);
        float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
        float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
        my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
        point_image = {ndc_to_pix(p_proj.x, a.W), ndc_to_pix(p_proj.y, a.H)};
        tile_rect(point_image, (int)my_radius, a.grid_x, a.grid_y, rect_min, rect_max);
        if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) != 0) {
          visible = true;
          opacity = a.opacities[idx];
          if (a.activate) opacity = act_sigmoid(opacity);
        }
      }
    }
  }

  // ---- SH -> RGB for the rendered ones: every lane fetches its own row with one bulk copy ----
  float3 rgb = {0.f, 0.f, 0.f};
  uint8_t clamp_bits = 0;
  if (a.colors_precomp == nullptr) {
    const int deg = (a.shs != nullptr) ? min(a.D, a.M >= 15 ? 3 : (a.M >= 8 ? 2 : (a.M >= 3 ? 1 : 0))) : 0;
    const unsigned vis_mask = __ballot_sync(0xffffffffu, visible);
    const int my_rank = __popc(vis_mask & ((1u << lane) - 1u));  // position among the warp's rendered Gaussians
    const int n_vis = __popc(vis_mask);
    float3 dir = {0.f, 0.f, 0.f};
    if (visible) {
      dir = {p_orig.x - a.campos[0], p_orig.y - a.campos[1], p_orig.z - a.campos[2]};
      const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
      dir = {dir.x / len, dir.y / len, dir.z / len};
    }
    const int ncoef = min(a.M, 15) * 3;
    for (int g0 = 0; g0 < (deg > 0 ? n_vis : 1); g0 += kShSlots) {
      const bool mine = visible && (deg == 0 || (my_rank >= g0 && my_rank < g0 + kShSlots));
      const float* sh = reinterpret_cast<const float*>(s_sh[warp]);
      if (deg > 0) {
        uint32_t bytes = 0;
        unsigned char* slot = s_sh[warp];
        uintptr_t row0 = 0;
        if (mine) {
          slot += (my_rank - g0) * kShRowBytes;
          const uintptr_t row = reinterpret_cast<uintptr_t>(a.shs + (size_t)idx * (size_t)a.M * 3);
          row0 = row & ~(uintptr_t)15;
          const uint32_t off = (uint32_t)(row - row0);
          const uint32_t want = (off + (uint32_t)ncoef * 4u + 15u) & ~15u;
          sh = reinterpret_cast<const float*>(slot + off);
          if (row0 + want <= reinterpret_cast<uintptr_t>(a.shs + (size_t)a.P * (size_t)a.M * 3)) {
            bytes = want;
          } else {   // the window of the tensor's last row would end past the tensor: element-wise copy
            const float* g = reinterpret_cast<const float*>(row);
            float* d = reinterpret_cast<float*>(slot + off);
            for (int k = 0; k < ncoef; ++k) d[k] = g[k];
          }
        }
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        if (total) {
          if (lane == 0) mbar_expect_tx(&s_bar[warp], total);
          __syncwarp();
          if (bytes) bulk_g2s(slot, reinterpret_cast<const void*>(row0), bytes, &s_bar[warp]);
          mbar_wait(&s_bar[warp], sh_phase);
          sh_phase ^= 1u;
        }
        __syncwarp();
      }
      if (mine) {
        float r = sh_channel(deg, dir.x, dir.y, dir.z, dc.x, sh + 0, 3);
        float g = sh_channel(deg, dir.x, dir.y, dir.z, dc.y, sh + 1, 3);
        float b = sh_channel(deg, dir.x, dir.y, dir.z, dc.z, sh + 2, 3);
        clamp_bits = (r < 0 ? 1 : 0) | (g < 0 ? 2 : 0) | (b < 0 ? 4 : 0);
        rgb = {fmaxf(r, 0.0f), fmaxf(g, 0.0f), fmaxf(b, 0.0f)};
      }
      __syncwarp();  // the slots are overwritten by the next round
    }
  }

  // ---- compacted list of the rendered Gaussians (radii > 0), one warp-aggregated atomic per warp.  Its order is
  // irrelevant: the emit stage and the backward preprocess only need the set, the depth order is established
  // per tile (binning.cu).  Every Gaussian with radii > 0 is listed (also the few whose tiles all fail the
  // opacity-aware test): the list is exactly the set of rows the backward pass writes, i.e. the radii mask.
  const unsigned vm = __ballot_sync(0xffffffffu, visible);
  uint32_t list_pos = 0;
  if (vm) {
    if (lane == 0) list_pos = atomicAdd(a.num_visible, (uint32_t)__popc(vm));
    list_pos = __shfl_sync(0xffffffffu, list_pos, 0) + __popc(vm & ((1u << lane) - 1u));
  }

  // ---- exact tile count (opacity-aware test) -----------------------------------
  const float4 co = {conic.x, conic.y, conic.z, opacity};
  // Very large rectangles are not walked here (one warp walking a 2000-tile rectangle was this kernel's critical
  // path: 60 steps behind everything else): they are cut into chunks of about kQueueRectTiles tiles, queued as
  // (chunk << 24 | list position) and counted by count_large_kernel, spread over the whole GPU — the same queue
  // layout the emit stage builds for itself later (binning.cu)
  bool large = false;
  if (visible) {
    const uint32_t area = (rect_max.x - rect_min.x) * (rect_max.y - rect_min.y);
    if (area > (uint32_t)kCountQueueTiles) {
      const uint32_t nchunks = queue_chunks(area);
      const uint32_t q0 = atomicAdd(a.large_count, nchunks);
      if (q0 + nchunks <= a.large_capacity) {
        for (uint32_t c = 0; c < nchunks; ++c) a.large_list[q0 + c] = (c << 24) | list_pos;
        large = true;
      } else {   // queue full (pathological input): poison the reserved entries and walk it right here
        for (uint32_t c = 0; c < nchunks && q0 + c < a.large_capacity; ++c) a.large_list[q0 + c] = 0xffffffffu;
      }
    }
  }
  // all other (Gaussian, tile) pairs of the warp, 32 per step; per-tile list lengths by RED (ranges, R and the
  // tile-sort classes all follow from them, binning.cu: tile_meta_kernel), per-Gaussian counts from the ballots
  const uint32_t ntiles = warp_count_rects_flat(visible && !large, co, point_image, rect_min, rect_max,
                                                visible ? opacity_power_threshold(opacity) : 0.f, a.tile_count, a.grid_x);

  // ---- outputs -----------------------------------------------------------------
  if (in_range) {
    a.radii[idx] = visible ? (int)my_radius : 0;
    a.tiles_touched[idx] = ntiles;   // queued rectangles: 0 here, count_large_kernel adds their tiles
    if (visible) {
      a.depths[idx] = p_view.z;
      a.means2D[idx] = point_image;
      a.conic_opacity[idx] = co;
      if (a.colors_precomp == nullptr) {
        a.rgb[3 * (size_t)idx + 0] = rgb.x;
        a.rgb[3 * (size_t)idx + 1] = rgb.y;
        a.rgb[3 * (size_t)idx + 2] = rgb.z;
        a.clamped[idx] = clamp_bits;
      }
      a.vis_items[list_pos] = ((uint64_t)__float_as_uint(p_view.z) << 32) | (uint32_t)idx;
      // the packed record the binning stage copies once per (Gaussian, tile) duplicate
      float3 c = rgb;
      if (a.colors_precomp != nullptr)
        c = {a.colors_precomp[3 * (size_t)idx], a.colors_precomp[3 * (size_t)idx + 1],
             a.colors_precomp[3 * (size_t)idx + 2]};
      // power below which opacity*exp(power) is certainly < 1/255 (1 % safety margin): lets the
      // blend kernels skip the exp for far-away pixels without changing any result
      const float reject = -logf(255.0f * opacity) - 0.01f;
      float4* r = a.splat_rec + 3 * (size_t)idx;
      r[0] = make_float4(point_image.x, point_image.y, conic.x, conic.y);
      r[1] = make_float4(conic.z, opacity, reject, __uint_as_float((uint32_t)idx));
      r[2] = make_float4(c.x, c.y, c.z, 0.f);
    }
  }
  }  // candidate loop
}

// ---- stage 3: the queued large rectangles, one warp per chunk (grid-stride over the queue, 32 tiles per step) ----
__global__ void __launch_bounds__(256) count_large_kernel(const PreprocessParams a) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t num_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t count = min(*a.large_count, a.large_capacity);
  for (uint32_t k = warp_global; k < count; k += num_warps) {
    const uint32_t item = a.large_list[k];
    if (item == 0xffffffffu) continue;
    const uint32_t j = item & ((1u << 24) - 1u);
    const uint32_t chunk = item >> 24;
    const uint32_t g = (uint32_t)(a.vis_items[j] & 0xffffffffu);
    const float2 xy = a.means2D[g];
    const float4 co = a.conic_opacity[g];
    uint2 rect_min, rect_max;
    tile_rect(xy, a.radii[g], a.grid_x, a.grid_y, rect_min, rect_max);
    const float thr = opacity_power_threshold(co.w);
    const float rcp_x = __frcp_rn(15 * 15 * co.x), rcp_z = __frcp_rn(15 * 15 * co.z);
    const int w = (int)(rect_max.x - rect_min.x), h = (int)(rect_max.y - rect_min.y);
    const int area = w * h;
    const int per_chunk = (area + (int)queue_chunks((uint32_t)area) - 1) / (int)queue_chunks((uint32_t)area);
    const int t_begin = (int)chunk * per_chunk;
    const int t_end = min(area, t_begin + per_chunk);
    const uint32_t w_magic = w > 1 ? 0xffffffffu / (uint32_t)w + 1u : 0u;   // t / w by multiplication (gsr_common.cuh)
    uint32_t n = 0;
    for (int t0 = t_begin; t0 < t_end; t0 += 32) {
      const int t = t0 + lane;
      bool hit = false;
      if (t < t_end) {
        const int q = w > 1 ? (int)__umulhi((uint32_t)t, w_magic) : t;
        const int y = (int)rect_min.y + q;
        const int x = (int)rect_min.x + t - q * w;
        hit = tile_min_power_pre(co, xy, x, y, rcp_x, rcp_z) <= thr;
        if (hit) atomicAdd(&a.tile_count[y * a.grid_x + x], 1u);
      }
      n += (uint32_t)__popc(__ballot_sync(0xffffffffu, hit));
    }
    if (lane == 0 && n) atomicAdd(&a.tiles_touched[g], n);
  }
}

int launch_precull(const PreprocessParams& p, cudaStream_t s) {
  if (p.P <= 0) return 0;
  precull_kernel<<<(p.P + kCullThreads * kCullRows - 1) / (kCullThreads * kCullRows), kCullThreads, 0, s>>>(p);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

int launch_preprocess(const PreprocessParams& p, cudaStream_t s) {
  if (p.P <= 0) return 0;
  // warps stride over the candidate list, whose length only the device knows: a grid that fills the machine
  const int ctas = min((p.P + kPreThreads - 1) / kPreThreads, 148 * GSR_PRE_MINBLOCKS);
  preprocess_kernel<<<ctas, kPreThreads, 0, s>>>(p);
  GSR_LAUNCHED();
  // 148 SMs x 4 CTAs x 8 warps count the queued large rectangles (a handful per view; an empty queue costs the launch)
  count_large_kernel<<<148 * 4, 256, 0, s>>>(p);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

// present[i] = view-space z > 0.2   (reference checkFrustum / in_frustum,
// rasterizer_impl.cu:122-135, auxiliary.h:160-194)
__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D,
                                    const float* __restrict__ view, uint8_t* __restrict__ present) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  const float3 p = {means3D[3 * (size_t)idx], means3D[3 * (size_t)idx + 1], means3D[3 * (size_t)idx + 2]};
  const float3 v = xform_4x3(p, view);
  present[idx] = (v.z <= 0.2f) ? 0 : 1;
}

int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present,
                        cudaStream_t s) {
  if (P <= 0) return 0;
  mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view, present);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
