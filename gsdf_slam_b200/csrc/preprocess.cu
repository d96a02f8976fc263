// Forward per-Gaussian pass: frustum cull, projection, 3D->2D covariance, conic,
// 3-sigma radius, SH->RGB, EXACT opacity-aware tile counts (per Gaussian and per
// tile), and the compacted list of rendered Gaussians.  One kernel replaces the reference's preprocessCUDA (forward.cu:197-312)
// plus the counting half of duplicateWithKeys (rasterizer_impl.cu:139-212): by
// running the tile test here, only tiles that will really be rasterised are
// counted, so no padding keys are ever generated.
//
// Memory plan (HBM-bound, SURVEY.md §8d: 236 B read + ~53 B written per visible
// Gaussian at SH-3): a CTA owns 128 consecutive Gaussians; the xyz / scale /
// rotation / opacity / dc rows of the CTA are fetched with coalesced float4
// loads into shared memory and each thread picks its own row from there; the
// 180-byte SH rows are fetched only for Gaussians that survive culling, one
// coalesced warp-wide read per surviving Gaussian.  Outputs are written SoA.
// The stage is LATENCY-bound, not bandwidth-bound (measured, round 2): a variant that fetched
// scale / rotation only for the z-cull survivors and compacted the survivors between its stages (256-Gaussian
// CTAs, barrier-separated phases, each with its own dependent global loads) moved fewer bytes and ran dense
// warps, but took 111 us instead of 59 us at cfg2 (410 vs 228 us at 3 M Gaussians); appending the
// (Gaussian, tile) hits found here to a list, so that the emit stage only streams them, cost this kernel more
// (+22 us) than it saved there (-16 us).  All input rows in flight at once is what matters here.
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
constexpr int kShFloats = 45;  // 15 coefficients x 3 channels
// SH rows of the surviving Gaussians of a warp are staged kShGroup at a time (a mapping view keeps ~5 of
// 32): 1.4 KB instead of 5.6 KB of shared memory per warp, which was what limited the resident CTAs
constexpr int kShGroup = 8;

#ifndef GSR_PRE_MINBLOCKS
#define GSR_PRE_MINBLOCKS 8
#endif
__global__ void __launch_bounds__(kPreThreads, GSR_PRE_MINBLOCKS) preprocess_kernel(const PreprocessParams a) {
  // CTA-staged input rows
  __shared__ __align__(16) float s_xyz[kPreThreads * 3];
  __shared__ __align__(16) float s_scale[kPreThreads * 3];
  __shared__ __align__(16) float s_rot[kPreThreads * 4];
  __shared__ __align__(16) float s_dc[kPreThreads * 3];
  __shared__ float s_sh[kPreThreads / 32][kShGroup * kShFloats];  // per-warp SH rows, kShGroup at a time
  __shared__ uint32_t s_tiles[kPreThreads / 32][32];        // per-warp tile counters

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int base = blockIdx.x * kPreThreads;
  const int count = min(kPreThreads, a.P - base);
  const int idx = base + tid;
  const bool in_range = tid < count;

  // ---- coalesced staging (rows of this CTA are contiguous in every tensor) ----
  {
    // base*3 and base*4 floats are multiples of 4 floats (base % 128 == 0): 16 B aligned rows
    const int n3 = count * 3, n4 = count * 4;
    const int v3 = a.vec_ok ? (n3 >> 2) : 0, v4 = a.vec_ok ? (n4 >> 2) : 0;
    const float* gx = a.means3D + (size_t)base * 3;
    for (int i = tid; i < v3; i += kPreThreads)
      reinterpret_cast<float4*>(s_xyz)[i] = reinterpret_cast<const float4*>(gx)[i];
    for (int k = v3 * 4 + tid; k < n3; k += kPreThreads) s_xyz[k] = gx[k];
    if (a.scales != nullptr) {
      const float* gs = a.scales + (size_t)base * 3;
      for (int i = tid; i < v3; i += kPreThreads)
        reinterpret_cast<float4*>(s_scale)[i] = reinterpret_cast<const float4*>(gs)[i];
      for (int k = v3 * 4 + tid; k < n3; k += kPreThreads) s_scale[k] = gs[k];
      const float* gr = a.rotations + (size_t)base * 4;
      for (int i = tid; i < v4; i += kPreThreads)
        reinterpret_cast<float4*>(s_rot)[i] = reinterpret_cast<const float4*>(gr)[i];
      for (int k = v4 * 4 + tid; k < n4; k += kPreThreads) s_rot[k] = gr[k];
    }
    if (a.colors_precomp == nullptr) {
      const float* gd = a.dc + (size_t)base * 3;
      for (int i = tid; i < v3; i += kPreThreads)
        reinterpret_cast<float4*>(s_dc)[i] = reinterpret_cast<const float4*>(gd)[i];
      for (int k = v3 * 4 + tid; k < n3; k += kPreThreads) s_dc[k] = gd[k];
    }
  }
  __syncthreads();
  // raw-parameter entry point (row f-2): every thread activates its own staged row in place, so
  // the geometry code below is the same instruction stream for both entry points
  if (a.activate && a.scales != nullptr && tid < count) {
    const float3 s3 = act_exp3({s_scale[3 * tid], s_scale[3 * tid + 1], s_scale[3 * tid + 2]});
    s_scale[3 * tid] = s3.x; s_scale[3 * tid + 1] = s3.y; s_scale[3 * tid + 2] = s3.z;
    reinterpret_cast<float4*>(s_rot)[tid] = act_normalize4(reinterpret_cast<const float4*>(s_rot)[tid]);
  }

  // ---- per-Gaussian geometry ---------------------------------------------------
  bool visible = false;
  float3 p_orig = {0.f, 0.f, 0.f};
  float3 p_view = {0.f, 0.f, 0.f};
  float my_radius = 0.f;
  float2 point_image = {0.f, 0.f};
  float3 conic = {0.f, 0.f, 0.f};
  uint2 rect_min = {0, 0}, rect_max = {0, 0};
  float opacity = 0.f;

  if (in_range) {
    p_orig = {s_xyz[3 * tid], s_xyz[3 * tid + 1], s_xyz[3 * tid + 2]};
    p_view = xform_4x3(p_orig, a.view);
    if (p_view.z <= 0.2f) {
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
        const float3 sc = {s_scale[3 * tid], s_scale[3 * tid + 1], s_scale[3 * tid + 2]};
        const float4 q = reinterpret_cast<const float4*>(s_rot)[tid];
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

  // ---- SH -> RGB for the survivors: coalesced warp-wide fetch of each 180 B row ----
  float3 rgb = {0.f, 0.f, 0.f};
  uint8_t clamp_bits = 0;
  if (a.colors_precomp == nullptr) {
    const int deg = (a.shs != nullptr) ? min(a.D, a.M >= 15 ? 3 : (a.M >= 8 ? 2 : (a.M >= 3 ? 1 : 0))) : 0;
    const unsigned vis_mask = __ballot_sync(0xffffffffu, visible);
    const int my_rank = __popc(vis_mask & ((1u << lane) - 1u));  // position among the warp's survivors
    const int n_vis = __popc(vis_mask);
    float3 dir = {0.f, 0.f, 0.f};
    if (visible) {
      dir = {p_orig.x - a.campos[0], p_orig.y - a.campos[1], p_orig.z - a.campos[2]};
      const float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
      dir = {dir.x / len, dir.y / len, dir.z / len};
    }
    for (int g0 = 0; g0 < (deg > 0 ? n_vis : 1); g0 += kShGroup) {
      if (deg > 0) {
        const int ncoef = min(a.M, 15) * 3;
        // the survivors with rank g0 .. g0 + kShGroup - 1: asynchronous copies, all rows in flight at once
        unsigned m = vis_mask;
        int rank = 0;
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          if (rank >= g0 && rank < g0 + kShGroup) {
            const float* row = a.shs + (size_t)(base + warp * 32 + src) * (size_t)a.M * 3;
            float* dst = &s_sh[warp][(rank - g0) * kShFloats];
            if (lane < ncoef) cp_async_f32(dst + lane, row + lane);
            if (lane + 32 < ncoef) cp_async_f32(dst + lane + 32, row + lane + 32);
          }
          ++rank;
        }
        cp_async_wait_all();
        __syncwarp();
      }
      if (visible && (deg == 0 || (my_rank >= g0 && my_rank < g0 + kShGroup))) {
        const float* sh = &s_sh[warp][(deg > 0 ? my_rank - g0 : 0) * kShFloats];
        float r = sh_channel(deg, dir.x, dir.y, dir.z, s_dc[3 * tid + 0], sh + 0, 3);
        float g = sh_channel(deg, dir.x, dir.y, dir.z, s_dc[3 * tid + 1], sh + 1, 3);
        float b = sh_channel(deg, dir.x, dir.y, dir.z, s_dc[3 * tid + 2], sh + 2, 3);
        clamp_bits = (r < 0 ? 1 : 0) | (g < 0 ? 2 : 0) | (b < 0 ? 4 : 0);
        rgb = {fmaxf(r, 0.0f), fmaxf(g, 0.0f), fmaxf(b, 0.0f)};
      }
      __syncwarp();  // the group's rows are overwritten by the next group
    }
  }

  // ---- exact tile count (opacity-aware test) -----------------------------------
  uint32_t ntiles = 0;
  const float4 co = {conic.x, conic.y, conic.z, opacity};
  {
    // all (Gaussian, tile) pairs of the warp, 32 per step (gsr_common.cuh)
    const float thr = visible ? opacity_power_threshold(opacity) : 0.f;
    s_tiles[warp][lane] = 0;
    __syncwarp();
    uint32_t* cnt = s_tiles[warp];
    uint32_t* tile_count = a.tile_count;
    const int grid_x = a.grid_x;
    warp_walk_rects_flat(visible, co, point_image, rect_min, rect_max, thr, 0u, 0u,
                         [cnt, tile_count, grid_x](int owner, int x, int y, uint32_t, uint32_t) {
                           atomicAdd(&cnt[owner], 1u);
                           // exact per-tile list length: ranges, R and the tile-sort histograms
                           // all follow from it (binning.cu: tile_meta_kernel)
                           atomicAdd(&tile_count[y * grid_x + x], 1u);
                         });
    __syncwarp();
    ntiles = s_tiles[warp][lane];
  }

  // ---- outputs -----------------------------------------------------------------
  // every Gaussian with radii > 0 is listed (also the few whose tiles all fail the opacity-aware
  // test): the list is exactly the set of rows the backward pass writes, i.e. the radii mask
  const bool rendered = visible;
  if (in_range) {
    a.radii[idx] = visible ? (int)my_radius : 0;
    a.tiles_touched[idx] = ntiles;
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
    }
  }
  // compacted list of the visible Gaussians (radii > 0), one warp-aggregated atomic per warp.  Its order is irrelevant: the emit stage and the backward
  // preprocess only need the set, the depth order is established per tile (binning.cu).
  const unsigned vm = __ballot_sync(0xffffffffu, rendered);
  if (vm) {
    uint32_t base_pos = 0;
    if (lane == 0) base_pos = atomicAdd(a.num_visible, (uint32_t)__popc(vm));
    base_pos = __shfl_sync(0xffffffffu, base_pos, 0);
    if (rendered) {
      a.vis_items[base_pos + __popc(vm & ((1u << lane) - 1u))] =
          ((uint64_t)__float_as_uint(p_view.z) << 32) | (uint32_t)idx;
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
}

int launch_preprocess(const PreprocessParams& p, cudaStream_t s) {
  if (p.P <= 0) return 0;
  preprocess_kernel<<<(p.P + kPreThreads - 1) / kPreThreads, kPreThreads, 0, s>>>(p);
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
