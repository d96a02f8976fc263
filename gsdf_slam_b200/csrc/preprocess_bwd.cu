// Backward per-Gaussian pass, one kernel: chains the 2D gradients produced by the
// blend backward (dL/dmean2D, dL/dconic, dL/dopacity, dL/dcolor) to the 3D mean,
// scale, rotation, SH coefficients and (optionally) the 3D covariance.
// Replaces the reference's computeCov2DCUDA + BACKWARD::preprocessCUDA pair
// (backward.cu:159-313, 402-461) and the shim's blanket zero-fill of ten gradient
// tensors (src/rasterize_points.cu:157-166): every output row is written exactly
// once here — zeros for Gaussians that were not rendered — and dL/dcov3D never
// round-trips through HBM unless the caller asks for it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr_common.cuh"
#include "gsr_params.cuh"
#include "tma.cuh"

namespace gsr {


constexpr float SH_C0 = 0.28209479177387814f;
constexpr float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                   0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                                   -0.5900435899266435f};

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3& operator+=(V3& a, V3 b) {
  a.x += b.x; a.y += b.y; a.z += b.z;
  return a;
}
__device__ __forceinline__ float dotv(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__device__ __forceinline__ void store3(float* p, size_t i, V3 v) {
  p[3 * i + 0] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z;
}

// Everything a Gaussian's backward pass reads from global memory, fetched up front: the loads are in flight
// together with the SH-row gather instead of one dependent round trip after the other (the kernel is bound by
// that latency chain).  With `accm` that includes the row's current parameter gradients (keyframe batches).
struct PreBwdIn {
  float4 acc0, acc1, acc2, cop, q;
  float3 mean, scale;
  uint8_t cl;
  float old_op;
  V3 old_dc, old_mean, old_scale;
  float4 old_rot;
};
__device__ __forceinline__ void preprocess_bwd_load(const PreBwdParams& a, const int idx, const bool accm, PreBwdIn& in) {
  const size_t i = (size_t)idx;
  in.acc0 = reinterpret_cast<const float4*>(a.acc)[3 * i + 0];
  in.acc1 = reinterpret_cast<const float4*>(a.acc)[3 * i + 1];
  in.acc2 = reinterpret_cast<const float4*>(a.acc)[3 * i + 2];
  in.cop = a.conic_opacity[i];
  in.mean = {a.means3D[3 * i], a.means3D[3 * i + 1], a.means3D[3 * i + 2]};
  in.scale = {0.f, 0.f, 0.f};
  in.q = {0.f, 0.f, 0.f, 0.f};
  if (a.scales != nullptr) {
    in.scale = {a.scales[3 * i], a.scales[3 * i + 1], a.scales[3 * i + 2]};
    in.q = reinterpret_cast<const float4*>(a.rotations)[i];
  }
  in.cl = (a.shs != nullptr) ? a.clamped[idx] : (uint8_t)0;
  in.old_op = 0.f;
  in.old_dc = in.old_mean = in.old_scale = {0.f, 0.f, 0.f};
  in.old_rot = {0.f, 0.f, 0.f, 0.f};
  if (accm) {
    in.old_op = a.dL_dopacity[i];
    in.old_dc = {a.dL_ddc[3 * i], a.dL_ddc[3 * i + 1], a.dL_ddc[3 * i + 2]};
    in.old_mean = {a.dL_dmean3D[3 * i], a.dL_dmean3D[3 * i + 1], a.dL_dmean3D[3 * i + 2]};
    if (a.scales != nullptr) {
      in.old_scale = {a.dL_dscale[3 * i], a.dL_dscale[3 * i + 1], a.dL_dscale[3 * i + 2]};
      in.old_rot = reinterpret_cast<const float4*>(a.dL_drot)[i];
    }
  }
}
// parameter gradients: overwrite, or add to the row's previous value (fetched by preprocess_bwd_load)
__device__ __forceinline__ void out3(float* p, size_t i, V3 v, bool acc, V3 old) {
  if (acc) { v.x += old.x; v.y += old.y; v.z += old.z; }
  store3(p, i, v);
}
// dL/dsh element in the shared row: preloaded with the row's previous value when accumulating
__device__ __forceinline__ void sh_put(V3* p, V3 v, bool acc) {
  if (acc) { v.x = __fadd_rn(p->x, v.x); v.y = __fadd_rn(p->y, v.y); v.z = __fadd_rn(p->z, v.z); }
  *p = v;
}

// Per-Gaussian body.  `sh_row` is this Gaussian's dL/dsh row (3*M floats) in SHARED memory — zeros, or with `accm`
// the row's current values; the warp writes the rows out afterwards.  `accm` (keyframe batches): add to the row's
// current parameter gradients instead of overwriting them — for every row, or (accumulate_rows) only for the rows
// an earlier keyframe of the batch has already written.
__device__ __forceinline__ void preprocess_bwd_one(const PreBwdParams& a, const int idx, const bool accm,
                                                   const PreBwdIn& in, float* sh_row, const float* sh_in_row) {
  const size_t i = (size_t)idx;
  const int M = a.M;
  // ---- accumulated 2D gradients ------------------------------------------------
  const float4 acc0 = in.acc0, acc1 = in.acc1, acc2 = in.acc2;
  // blend_bwd.cu accumulated, over every (pixel, this splat) pair with w = G*dL/dalpha and
  // d = mean2D - pixel:  Mx = sum w*dx, My, Mxx = sum w*dx^2, Mxy, Myy, M0 = sum w, colour sums.
  // The reference's per-pair terms (backward.cu:622-637) are linear in these moments:
  //   dL/dmean2D.x = sum dL_dG*(-G*(a*dx+b*dy))*0.5W = -op*0.5W*(a*Mx + b*My)     (dL_dG = op*dL/dalpha)
  //   dL/dconic.x  = sum -0.5*G*dx*dx*dL_dG          = -0.5*op*Mxx   (same for xy, yy)
  //   dL/dopacity  = sum G*dL/dalpha                 = M0
  const float4 cop = in.cop;
  const float Mx = acc0.x, My = acc0.y, Mxx = acc0.z, Mxy = acc0.w, Myy = acc1.x, M0 = acc1.y;
  const float ddelx_dx = 0.5f * a.W, ddely_dy = 0.5f * a.H;
  const float dmean2D_x = -cop.w * ddelx_dx * (cop.x * Mx + cop.y * My);
  const float dmean2D_y = -cop.w * ddely_dy * (cop.z * My + cop.y * Mx);
  const float3 dL_dconic = {-0.5f * cop.w * Mxx, -0.5f * cop.w * Mxy, -0.5f * cop.w * Myy};
  const float dopacity = M0;
  V3 dL_dRGB = {acc1.z, acc1.w, acc2.x};

  store3(a.dL_dmean2D, i, {dmean2D_x, dmean2D_y, 0.f});
  if (a.dL_dconic)
    reinterpret_cast<float4*>(a.dL_dconic)[i] = make_float4(dL_dconic.x, dL_dconic.y, 0.f, dL_dconic.z);
  // raw-parameter entry point: chain through sigmoid (sigma' = o (1 - o), o = the activated opacity)
  {
    const float dop = a.activate ? dopacity * cop.w * (1.0f - cop.w) : dopacity;
    a.dL_dopacity[i] = accm ? in.old_op + dop : dop;
  }
  store3(a.dL_dcolor, i, dL_dRGB);

  const float3 mean = in.mean;

  // ---- 3D covariance (recomputed rather than stored by the forward) -------------
  float cov3D[6];
  float3 scale = {0.f, 0.f, 0.f};
  float4 q = {0.f, 0.f, 0.f, 0.f};
  float q_norm = 1.0f;
  M3 R, Mm;
  if (a.scales != nullptr) {
    scale = in.scale;
    q = in.q;
    if (a.activate) {
      q_norm = quat_norm(q);
      scale = act_exp3(scale);
      q = {q.x / q_norm, q.y / q_norm, q.z / q_norm, q.w / q_norm};
    }
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    R = m3_make(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    M3 S = m3_make(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
    S[0][0] = a.scale_modifier * scale.x;
    S[1][1] = a.scale_modifier * scale.y;
    S[2][2] = a.scale_modifier * scale.z;
    Mm = m3_mul(S, R);
    M3 Sigma = m3_mul(m3_transpose(Mm), Mm);
    cov3D[0] = Sigma[0][0]; cov3D[1] = Sigma[0][1]; cov3D[2] = Sigma[0][2];
    cov3D[3] = Sigma[1][1]; cov3D[4] = Sigma[1][2]; cov3D[5] = Sigma[2][2];
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) cov3D[k] = a.cov3D_precomp[6 * i + k];
  }

  // ---- conic -> cov2D -> cov3D / mean  (reference computeCov2DCUDA, backward.cu:159-313)
  float3 t = xform_4x3(mean, a.view);
  const float h_x = a.focal_x, h_y = a.focal_y;
  const float limx = 1.3f * a.tan_fovx;
  const float limy = 1.3f * a.tan_fovy;
  const float txtz = t.x / t.z;
  const float tytz = t.y / t.z;
  t.x = min(limx, max(-limx, txtz)) * t.z;
  t.y = min(limy, max(-limy, tytz)) * t.z;
  const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
  const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;

  M3 J = m3_make(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z), 0.0f, h_y / t.z,
                 -(h_y * t.y) / (t.z * t.z), 0, 0, 0);
  const float* vm = a.view;
  M3 Wm = m3_make(vm[0], vm[4], vm[8], vm[1], vm[5], vm[9], vm[2], vm[6], vm[10]);
  M3 Vrk = m3_make(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4],
                   cov3D[5]);
  M3 T = m3_mul(Wm, J);
  M3 cov2D = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);

  float aa = cov2D[0][0] += 0.3f;
  float bb = cov2D[0][1];
  float cc = cov2D[1][1] += 0.3f;

  float denom = aa * cc - bb * bb;
  float dL_da = 0, dL_db = 0, dL_dc = 0;
  float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

  float dcov[6];
  if (denom2inv != 0) {
    dL_da = denom2inv * (-cc * cc * dL_dconic.x + 2 * bb * cc * dL_dconic.y + (denom - aa * cc) * dL_dconic.z);
    dL_dc = denom2inv * (-aa * aa * dL_dconic.z + 2 * aa * bb * dL_dconic.y + (denom - aa * cc) * dL_dconic.x);
    dL_db = denom2inv * 2 * (bb * cc * dL_dconic.x - (denom + 2 * bb * bb) * dL_dconic.y + aa * bb * dL_dconic.z);

    dcov[0] = (T[0][0] * T[0][0] * dL_da + T[0][0] * T[1][0] * dL_db + T[1][0] * T[1][0] * dL_dc);
    dcov[3] = (T[0][1] * T[0][1] * dL_da + T[0][1] * T[1][1] * dL_db + T[1][1] * T[1][1] * dL_dc);
    dcov[5] = (T[0][2] * T[0][2] * dL_da + T[0][2] * T[1][2] * dL_db + T[1][2] * T[1][2] * dL_dc);
    dcov[1] = 2 * T[0][0] * T[0][1] * dL_da + (T[0][0] * T[1][1] + T[0][1] * T[1][0]) * dL_db +
              2 * T[1][0] * T[1][1] * dL_dc;
    dcov[2] = 2 * T[0][0] * T[0][2] * dL_da + (T[0][0] * T[1][2] + T[0][2] * T[1][0]) * dL_db +
              2 * T[1][0] * T[1][2] * dL_dc;
    dcov[4] = 2 * T[0][2] * T[0][1] * dL_da + (T[0][1] * T[1][2] + T[0][2] * T[1][1]) * dL_db +
              2 * T[1][1] * T[1][2] * dL_dc;
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) dcov[k] = 0;
  }
  if (a.dL_dcov3D) {
#pragma unroll
    for (int k = 0; k < 6; ++k) a.dL_dcov3D[6 * i + k] = dcov[k];
  }

  float dL_dT00 = 2 * (T[0][0] * Vrk[0][0] + T[0][1] * Vrk[0][1] + T[0][2] * Vrk[0][2]) * dL_da +
                  (T[1][0] * Vrk[0][0] + T[1][1] * Vrk[0][1] + T[1][2] * Vrk[0][2]) * dL_db;
  float dL_dT01 = 2 * (T[0][0] * Vrk[1][0] + T[0][1] * Vrk[1][1] + T[0][2] * Vrk[1][2]) * dL_da +
                  (T[1][0] * Vrk[1][0] + T[1][1] * Vrk[1][1] + T[1][2] * Vrk[1][2]) * dL_db;
  float dL_dT02 = 2 * (T[0][0] * Vrk[2][0] + T[0][1] * Vrk[2][1] + T[0][2] * Vrk[2][2]) * dL_da +
                  (T[1][0] * Vrk[2][0] + T[1][1] * Vrk[2][1] + T[1][2] * Vrk[2][2]) * dL_db;
  float dL_dT10 = 2 * (T[1][0] * Vrk[0][0] + T[1][1] * Vrk[0][1] + T[1][2] * Vrk[0][2]) * dL_dc +
                  (T[0][0] * Vrk[0][0] + T[0][1] * Vrk[0][1] + T[0][2] * Vrk[0][2]) * dL_db;
  float dL_dT11 = 2 * (T[1][0] * Vrk[1][0] + T[1][1] * Vrk[1][1] + T[1][2] * Vrk[1][2]) * dL_dc +
                  (T[0][0] * Vrk[1][0] + T[0][1] * Vrk[1][1] + T[0][2] * Vrk[1][2]) * dL_db;
  float dL_dT12 = 2 * (T[1][0] * Vrk[2][0] + T[1][1] * Vrk[2][1] + T[1][2] * Vrk[2][2]) * dL_dc +
                  (T[0][0] * Vrk[2][0] + T[0][1] * Vrk[2][1] + T[0][2] * Vrk[2][2]) * dL_db;

  float dL_dJ00 = Wm[0][0] * dL_dT00 + Wm[0][1] * dL_dT01 + Wm[0][2] * dL_dT02;
  float dL_dJ02 = Wm[2][0] * dL_dT00 + Wm[2][1] * dL_dT01 + Wm[2][2] * dL_dT02;
  float dL_dJ11 = Wm[1][0] * dL_dT10 + Wm[1][1] * dL_dT11 + Wm[1][2] * dL_dT12;
  float dL_dJ12 = Wm[2][0] * dL_dT10 + Wm[2][1] * dL_dT11 + Wm[2][2] * dL_dT12;

  float tz = 1.f / t.z;
  float tz2 = tz * tz;
  float tz3 = tz2 * tz;

  float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
  float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
  float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                 (2 * h_y * t.y) * tz3 * dL_dJ12;

  // transformVec4x3Transpose (auxiliary.h:101-110)
  V3 dL_dmean_total = {vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz,
                       vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz,
                       vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz};

  // ---- screen-space mean gradient -> 3D mean (reference preprocessCUDA, backward.cu:427-449)
  {
    const float* proj = a.proj;
    const float3 m = mean;
    float4 m_hom = xform_4x4(m, proj);
    float m_w = 1.0f / (m_hom.w + 0.0000001f);
    float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
    float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
    V3 dL_dmean;
    dL_dmean.x = (proj[0] * m_w - proj[3] * mul1) * dmean2D_x + (proj[1] * m_w - proj[3] * mul2) * dmean2D_y;
    dL_dmean.y = (proj[4] * m_w - proj[7] * mul1) * dmean2D_x + (proj[5] * m_w - proj[7] * mul2) * dmean2D_y;
    dL_dmean.z = (proj[8] * m_w - proj[11] * mul1) * dmean2D_x + (proj[9] * m_w - proj[11] * mul2) * dmean2D_y;
    dL_dmean_total += dL_dmean;
  }

  // ---- colour -> SH coefficients and view direction (backward.cu:27-154) --------
  V3 out_ddc = {0.f, 0.f, 0.f};
  if (a.shs != nullptr) {
    // never index past the coefficients the caller provided
    const int deg = min(a.D, M >= 15 ? 3 : (M >= 8 ? 2 : (M >= 3 ? 1 : 0)));
    const V3 campos = {a.campos[0], a.campos[1], a.campos[2]};
    const V3 dir_orig = {mean.x - campos.x, mean.y - campos.y, mean.z - campos.z};
    const float len = sqrtf(dotv(dir_orig, dir_orig));
    const V3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};
    const V3* sh = reinterpret_cast<const V3*>(sh_in_row);  // this Gaussian's SH row (staged)
    V3* dL_dsh = reinterpret_cast<V3*>(sh_row);

    const uint8_t cl = in.cl;
    dL_dRGB.x *= (cl & 1) ? 0 : 1;
    dL_dRGB.y *= (cl & 2) ? 0 : 1;
    dL_dRGB.z *= (cl & 4) ? 0 : 1;

    V3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
    const float x = dir.x, y = dir.y, z = dir.z;

    float dRGBdsh0 = SH_C0;
    out_ddc = dRGBdsh0 * dL_dRGB;
    int written = 0;
    if (deg > 0) {
      float dRGBdsh1 = -SH_C1 * y;
      float dRGBdsh2 = SH_C1 * z;
      float dRGBdsh3 = -SH_C1 * x;
      sh_put(&dL_dsh[0], dRGBdsh1 * dL_dRGB, accm);
      sh_put(&dL_dsh[1], dRGBdsh2 * dL_dRGB, accm);
      sh_put(&dL_dsh[2], dRGBdsh3 * dL_dRGB, accm);
      written = 3;
      dRGBdx = -SH_C1 * sh[2];
      dRGBdy = -SH_C1 * sh[0];
      dRGBdz = SH_C1 * sh[1];

      if (deg > 1) {
        float xx = x * x, yy = y * y, zz = z * z;
        float xy = x * y, yz = y * z, xz = x * z;
        float dRGBdsh4 = SH_C2[0] * xy;
        float dRGBdsh5 = SH_C2[1] * yz;
        float dRGBdsh6 = SH_C2[2] * (2.f * zz - xx - yy);
        float dRGBdsh7 = SH_C2[3] * xz;
        float dRGBdsh8 = SH_C2[4] * (xx - yy);
        sh_put(&dL_dsh[3], dRGBdsh4 * dL_dRGB, accm);
        sh_put(&dL_dsh[4], dRGBdsh5 * dL_dRGB, accm);
        sh_put(&dL_dsh[5], dRGBdsh6 * dL_dRGB, accm);
        sh_put(&dL_dsh[6], dRGBdsh7 * dL_dRGB, accm);
        sh_put(&dL_dsh[7], dRGBdsh8 * dL_dRGB, accm);
        written = 8;

        dRGBdx += SH_C2[0] * y * sh[3] + SH_C2[2] * 2.f * -x * sh[5] + SH_C2[3] * z * sh[6] +
                  SH_C2[4] * 2.f * x * sh[7];
        dRGBdy += SH_C2[0] * x * sh[3] + SH_C2[1] * z * sh[4] + SH_C2[2] * 2.f * -y * sh[5] +
                  SH_C2[4] * 2.f * -y * sh[7];
        dRGBdz += SH_C2[1] * y * sh[4] + SH_C2[2] * 2.f * 2.f * z * sh[5] + SH_C2[3] * x * sh[6];

        if (deg > 2) {
          float dRGBdsh9 = SH_C3[0] * y * (3.f * xx - yy);
          float dRGBdsh10 = SH_C3[1] * xy * z;
          float dRGBdsh11 = SH_C3[2] * y * (4.f * zz - xx - yy);
          float dRGBdsh12 = SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy);
          float dRGBdsh13 = SH_C3[4] * x * (4.f * zz - xx - yy);
          float dRGBdsh14 = SH_C3[5] * z * (xx - yy);
          float dRGBdsh15 = SH_C3[6] * x * (xx - 3.f * yy);
          sh_put(&dL_dsh[8], dRGBdsh9 * dL_dRGB, accm);
          sh_put(&dL_dsh[9], dRGBdsh10 * dL_dRGB, accm);
          sh_put(&dL_dsh[10], dRGBdsh11 * dL_dRGB, accm);
          sh_put(&dL_dsh[11], dRGBdsh12 * dL_dRGB, accm);
          sh_put(&dL_dsh[12], dRGBdsh13 * dL_dRGB, accm);
          sh_put(&dL_dsh[13], dRGBdsh14 * dL_dRGB, accm);
          sh_put(&dL_dsh[14], dRGBdsh15 * dL_dRGB, accm);
          written = 15;

          dRGBdx += (SH_C3[0] * sh[8] * 3.f * 2.f * xy + SH_C3[1] * sh[9] * yz +
                     SH_C3[2] * sh[10] * -2.f * xy + SH_C3[3] * sh[11] * -3.f * 2.f * xz +
                     SH_C3[4] * sh[12] * (-3.f * xx + 4.f * zz - yy) + SH_C3[5] * sh[13] * 2.f * xz +
                     SH_C3[6] * sh[14] * 3.f * (xx - yy));
          dRGBdy += (SH_C3[0] * sh[8] * 3.f * (xx - yy) + SH_C3[1] * sh[9] * xz +
                     SH_C3[2] * sh[10] * (-3.f * yy + 4.f * zz - xx) +
                     SH_C3[3] * sh[11] * -3.f * 2.f * yz + SH_C3[4] * sh[12] * -2.f * xy +
                     SH_C3[5] * sh[13] * -2.f * yz + SH_C3[6] * sh[14] * -3.f * 2.f * xy);
          dRGBdz += (SH_C3[1] * sh[9] * xy + SH_C3[2] * sh[10] * 4.f * 2.f * yz +
                     SH_C3[3] * sh[11] * 3.f * (2.f * zz - xx - yy) +
                     SH_C3[4] * sh[12] * 4.f * 2.f * xz + SH_C3[5] * sh[13] * (xx - yy));
        }
      }
    }
    (void)written;  // rows beyond the active degree stay zero (pre-zeroed shared row)

    V3 dL_ddir = {dotv(dRGBdx, dL_dRGB), dotv(dRGBdy, dL_dRGB), dotv(dRGBdz, dL_dRGB)};
    // dnormvdv (auxiliary.h:122-138)
    {
      const V3 v = dir_orig, dv = dL_ddir;
      float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
      float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
      V3 r;
      r.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
      r.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
      r.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
      dL_dmean_total += r;
    }
  }
  out3(a.dL_ddc, i, out_ddc, accm, in.old_dc);
  out3(a.dL_dmean3D, i, dL_dmean_total, accm, in.old_mean);

  // ---- cov3D -> scale, rotation (reference computeCov3D backward, backward.cu:317-397)
  if (a.scales != nullptr) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const V3 s = {a.scale_modifier * scale.x, a.scale_modifier * scale.y, a.scale_modifier * scale.z};
    M3 dL_dSigma = m3_make(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3],
                           0.5f * dcov[4], 0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
    M3 M2;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 3; ++k) M2[c][k] = Mm[c][k] * 2.0f;
    M3 dL_dM = m3_mul(M2, dL_dSigma);
    M3 Rt = m3_transpose(R);
    M3 dL_dMt = m3_transpose(dL_dM);

    V3 dscale;
    dscale.x = dot3(Rt[0], dL_dMt[0]);
    dscale.y = dot3(Rt[1], dL_dMt[1]);
    dscale.z = dot3(Rt[2], dL_dMt[2]);
    // raw-parameter entry point: d exp(s)/ds = exp(s) = the activated scale
    if (a.activate) dscale = {dscale.x * scale.x, dscale.y * scale.y, dscale.z * scale.z};
    out3(a.dL_dscale, i, dscale, accm, in.old_scale);

#pragma unroll
    for (int k = 0; k < 3; ++k) {
      dL_dMt[0][k] *= s.x;
      dL_dMt[1][k] *= s.y;
      dL_dMt[2][k] *= s.z;
    }
    float4 dL_dq;
    dL_dq.x = 2 * z * (dL_dMt[0][1] - dL_dMt[1][0]) + 2 * y * (dL_dMt[2][0] - dL_dMt[0][2]) +
              2 * x * (dL_dMt[1][2] - dL_dMt[2][1]);
    dL_dq.y = 2 * y * (dL_dMt[1][0] + dL_dMt[0][1]) + 2 * z * (dL_dMt[2][0] + dL_dMt[0][2]) +
              2 * r * (dL_dMt[1][2] - dL_dMt[2][1]) - 4 * x * (dL_dMt[2][2] + dL_dMt[1][1]);
    dL_dq.z = 2 * x * (dL_dMt[1][0] + dL_dMt[0][1]) + 2 * r * (dL_dMt[2][0] - dL_dMt[0][2]) +
              2 * z * (dL_dMt[1][2] + dL_dMt[2][1]) - 4 * y * (dL_dMt[2][2] + dL_dMt[0][0]);
    dL_dq.w = 2 * r * (dL_dMt[0][1] - dL_dMt[1][0]) + 2 * x * (dL_dMt[2][0] + dL_dMt[0][2]) +
              2 * y * (dL_dMt[1][2] + dL_dMt[2][1]) - 4 * z * (dL_dMt[1][1] + dL_dMt[0][0]);
    if (a.activate) {
      // q = raw / |raw|: d raw = (dq - q (q . dq)) / |raw|
      const float qd = q.x * dL_dq.x + q.y * dL_dq.y + q.z * dL_dq.z + q.w * dL_dq.w;
      dL_dq = {(dL_dq.x - q.x * qd) / q_norm, (dL_dq.y - q.y * qd) / q_norm, (dL_dq.z - q.z * qd) / q_norm,
               (dL_dq.w - q.w * qd) / q_norm};
    }
    if (accm) {
      const float4 o = in.old_rot;
      reinterpret_cast<float4*>(a.dL_drot)[i] = make_float4(o.x + dL_dq.x, o.y + dL_dq.y, o.z + dL_dq.z, o.w + dL_dq.w);
    } else {
      reinterpret_cast<float4*>(a.dL_drot)[i] = dL_dq;
    }
  } else if (!accm) {
    store3(a.dL_dscale, i, {0.f, 0.f, 0.f});
    reinterpret_cast<float4*>(a.dL_drot)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

constexpr int kPreBwdThreads = 128;
constexpr int kZeroThreads = 256;
constexpr int kZeroMaxTensors = 12;

// ---- pass 1: dense zero fill of every gradient tensor (coalesced float4 stream) ----------
// The API returns dense [P, ...] gradient tensors while only the rendered Gaussians (~15 % of
// a mapping view) have non-zero rows.  Writing the zeros as one vectorised multi-tensor stream
// and then computing only the rendered rows from the compacted depth-sorted list is much
// cheaper than letting every thread of a P-wide kernel write its own strided zero row.
struct ZeroParams {
  float* ptr[kZeroMaxTensors];
  long long numel[kZeroMaxTensors];
  int first_block[kZeroMaxTensors + 1];
  int n;
};
constexpr int kZeroChunk = kZeroThreads * 4 * 4;  // floats per CTA

__global__ void __launch_bounds__(kZeroThreads) zero_fill_kernel(const ZeroParams z) {
  int t = 0;
#pragma unroll 1
  for (int k = 1; k < z.n; ++k)
    if ((int)blockIdx.x >= z.first_block[k]) t = k;
  float* p = z.ptr[t];
  const long long n = z.numel[t];
  const long long base = (long long)(blockIdx.x - z.first_block[t]) * kZeroChunk;
  const long long end = min(n, base + kZeroChunk);
  if ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && end - base == kZeroChunk) {
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      __stcs(reinterpret_cast<float4*>(p) + base / 4 + k * kZeroThreads + threadIdx.x, zero4);
  } else {
    for (long long e = base + threadIdx.x; e < end; e += kZeroThreads) p[e] = 0.f;
  }
}

// ---- pass 1b: which rendered Gaussians received a gradient -----------------------------------------------
// A rendered Gaussian that no pixel blended (hidden behind opaque ones, or beyond the point where its tiles'
// pixels saturated) has all-zero moment sums, hence all-zero gradients: nothing to compute — 85 % of the rendered
// rows of a 3 M-Gaussian room view.  One thread per entry of the forward's list reads the 48-byte moment row and
// appends the Gaussians with a gradient to a second list (one atomic per CTA); pass 2 runs on that list with
// dense warps.  A row without gradient still gets its per-view outputs (zeros) unless the outputs were
// zero-filled, and is kept in the list when nobody else would define its parameter-gradient rows (rendered_rows_only
// without grad_rows, first writer).
constexpr int kSelThreads = 256;
__global__ void __launch_bounds__(kSelThreads)
select_grad_rows_kernel(const PreBwdParams a, const uint64_t* __restrict__ vis_items, const uint32_t* __restrict__ num_visible,
                        uint32_t* __restrict__ list, uint32_t* __restrict__ num_list) {
  __shared__ uint32_t s_off[kSelThreads / 32];
  __shared__ uint32_t s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t nv = *num_visible;
  for (uint32_t t0 = blockIdx.x * kSelThreads; t0 < nv; t0 += gridDim.x * kSelThreads) {   // uniform per CTA
    const uint32_t t = t0 + threadIdx.x;
    bool keep = false;
    uint32_t idx = 0;
    if (t < nv) {
      idx = (uint32_t)(vis_items[t] & 0xffffffffu);
      const float4 m0 = reinterpret_cast<const float4*>(a.acc)[3 * (size_t)idx + 0];
      const float4 m1 = reinterpret_cast<const float4*>(a.acc)[3 * (size_t)idx + 1];
      const float m2 = a.acc[12 * (size_t)idx + 8];
      const bool has_grad = m0.x != 0.f || m0.y != 0.f || m0.z != 0.f || m0.w != 0.f || m1.x != 0.f || m1.y != 0.f ||
                            m1.z != 0.f || m1.w != 0.f || m2 != 0.f;
      bool accm = false;
      if (a.accumulate != 0)
        accm = a.grad_rows != nullptr ? a.grad_rows[idx] > 0 : (a.accumulate_rows == nullptr || a.accumulate_rows[idx] > 0);
      keep = has_grad || !(a.dense_zeroed || a.grad_rows != nullptr || accm);
      if (!keep && !a.dense_zeroed) {   // the per-view outputs of a rendered row are always defined
        store3(a.dL_dmean2D, (size_t)idx, {0.f, 0.f, 0.f});
        if (a.dL_dconic) reinterpret_cast<float4*>(a.dL_dconic)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        store3(a.dL_dcolor, (size_t)idx, {0.f, 0.f, 0.f});
        if (a.dL_dcov3D)
          for (int k = 0; k < 6; ++k) a.dL_dcov3D[6 * (size_t)idx + k] = 0.f;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_off[warp] = (uint32_t)__popc(m);
    __syncthreads();
    if (warp == 0) {
      const uint32_t c = lane < kSelThreads / 32 ? s_off[lane] : 0u;
      uint32_t incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (lane < kSelThreads / 32) s_off[lane] = incl - c;
      if (lane == 31 && incl) s_base = atomicAdd(num_list, incl);
    }
    __syncthreads();
    if (keep) list[s_base + s_off[warp] + __popc(m & ((1u << lane) - 1u))] = idx;
    __syncthreads();   // s_off / s_base are reused by the next round
  }
}

// ---- pass 2: the rendered Gaussians, taken from the depth-sorted list of the forward ------
// sorted[0 .. *num_valid) are exactly the Gaussians that own at least one tile; every other
// Gaussian received no gradient from the blend stage, so all of its outputs are zero.
// 4 CTAs per SM = at most 128 registers (162 without the bound: 3 CTAs, measured slower on dense views)
#ifndef GSR_PREBWD_MINBLOCKS
#define GSR_PREBWD_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(kPreBwdThreads, GSR_PREBWD_MINBLOCKS)
preprocess_bwd_kernel(const PreBwdParams a, const uint32_t* __restrict__ list,
                      const uint32_t* __restrict__ num_valid) {
  // per-warp staging: SH rows in (coalesced 180 B reads), dL/dsh rows out (coalesced writes)
  __shared__ float s_sh[kPreBwdThreads / 32][32 * 45];
  __shared__ float s_dsh[kPreBwdThreads / 32][32 * 45];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = 3 * (a.M < 15 ? a.M : 15);
  const bool has_sh = a.shs != nullptr && row > 0;
  const uint32_t nv = *num_valid;
  const uint32_t stride = gridDim.x * kPreBwdThreads;
  const uint32_t rounds = (nv + stride - 1) / stride;
  for (uint32_t r = 0; r < rounds; ++r) {
    const uint32_t t = r * stride + blockIdx.x * kPreBwdThreads + tid;
    const bool act = t < nv;
    const int idx = act ? (int)list[t] : -1;
    // keyframe batches: add to the row's current parameter gradients — for every row, or only for the rows an
    // earlier keyframe of the batch has written (grad_rows, kept by these kernels; or the caller's accumulate_rows)
    bool accm = false;
    if (act && a.accumulate != 0)
      accm = a.grad_rows != nullptr ? a.grad_rows[idx] > 0 : (a.accumulate_rows == nullptr || a.accumulate_rows[idx] > 0);
    PreBwdIn in;
    if (act) {
      preprocess_bwd_load(a, idx, accm, in);
      if (a.grad_rows != nullptr && !accm) a.grad_rows[idx] = 1;
    }
    const unsigned accm_mask = __ballot_sync(0xffffffffu, accm);
    const unsigned act_mask = __ballot_sync(0xffffffffu, act);
    // (the grid is sized for the whole map while a view renders a fraction of it: three quarters of the warps have
    // nothing to do, and walking 32 empty sources twice was a third of the kernel's instructions)
    if (act_mask == 0) continue;
    // gather the SH rows — and, for the rows that accumulate, their current dL/dsh rows — with asynchronous
    // copies: all of them are in flight together (and together with the loads above)
    for (unsigned todo = act_mask; todo; todo &= todo - 1) {
      const int src = __ffs(todo) - 1;
      const int g = __shfl_sync(0xffffffffu, idx, src);
      if (has_sh) {
        const float* grow = a.shs + (size_t)g * (size_t)a.M * 3;
        float* dst = &s_sh[warp][src * row];
        if (lane < row) cp_async_f32(dst + lane, grow + lane);
        if (lane + 32 < row) cp_async_f32(dst + lane + 32, grow + lane + 32);
      }
      float* drow = &s_dsh[warp][src * row];
      if (((accm_mask >> src) & 1u) && a.dL_dsh != nullptr) {
        const float* old = a.dL_dsh + (size_t)g * row;
        if (lane < row) cp_async_f32(drow + lane, old + lane);
        if (lane + 32 < row) cp_async_f32(drow + lane + 32, old + lane + 32);
      } else {
        if (lane < row) drow[lane] = 0.f;
        if (lane + 32 < row) drow[lane + 32] = 0.f;
      }
    }
    cp_async_wait_all();
    __syncwarp();
    float* my_row = &s_dsh[warp][lane * row];
    if (act) preprocess_bwd_one(a, idx, accm, in, my_row, &s_sh[warp][lane * row]);
    __syncwarp();
    if (a.dL_dsh != nullptr && row > 0) {
      // each rendered row leaves as one contiguous 12*M-byte segment (sums included: plain stores)
      for (unsigned todo = act_mask; todo; todo &= todo - 1) {
        const int src = __ffs(todo) - 1;
        const int g = __shfl_sync(0xffffffffu, idx, src);
        float* dst = a.dL_dsh + (size_t)g * row;
        const float* srow = &s_dsh[warp][src * row];
        if (lane < row) dst[lane] = srow[lane];
        if (lane + 32 < row) dst[lane + 32] = srow[lane + 32];
      }
    }
    __syncwarp();
  }
}

int launch_preprocess_bwd(const PreBwdParams& p, const uint64_t* vis_items, const uint32_t* num_visible,
                          uint32_t* list, uint32_t* num_list, bool zero_fill, cudaStream_t s) {
  if (p.P <= 0) return 0;
  if (p.M > 15) return GSR_ERR_BAD_ARG;
  const long long P = p.P;
  ZeroParams z;
  z.n = 0;
  int blocks = 0;
  auto add = [&](float* ptr, long long numel) {
    if (ptr == nullptr || numel <= 0) return;
    z.ptr[z.n] = ptr;
    z.numel[z.n] = numel;
    z.first_block[z.n] = blocks;
    blocks += (int)((numel + kZeroChunk - 1) / kZeroChunk);
    ++z.n;
  };
  add(p.dL_dmean2D, 3 * P); add(p.dL_dconic, 4 * P); add(p.dL_dopacity, P); add(p.dL_dcolor, 3 * P);
  add(p.dL_dmean3D, 3 * P); add(p.dL_dcov3D, 6 * P); add(p.dL_ddc, 3 * P);
  add(p.dL_dsh, 3LL * (p.M < 15 ? p.M : 15) * P); add(p.dL_dscale, 3 * P); add(p.dL_drot, 4 * P);
  z.first_block[z.n] = blocks;
  if (blocks > 0 && zero_fill) {
    zero_fill_kernel<<<blocks, kZeroThreads, 0, s>>>(z);
    GSR_LAUNCHED();
  }
  // the Gaussians that received a gradient (*num_list, zeroed by the caller) ...
  const int want_sel = (p.P + kSelThreads - 1) / kSelThreads;
  select_grad_rows_kernel<<<want_sel < 148 * 8 ? want_sel : 148 * 8, kSelThreads, 0, s>>>(p, vis_items, num_visible, list,
                                                                                         num_list);
  GSR_LAUNCHED();
  // ... and their gradients: enough CTAs for one round in the common case (<= 300 k rows); grid-stride beyond
  const int want = (p.P + kPreBwdThreads - 1) / kPreBwdThreads;
  preprocess_bwd_kernel<<<want < 148 * 16 ? want : 148 * 16, kPreBwdThreads, 0, s>>>(p, list, num_list);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // namespace gsr
