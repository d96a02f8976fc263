// TEST INFRASTRUCTURE ONLY — CPU ORACLE.
//
// A scalar C++ restatement of the reference's rasterizer hot path, because the reference
// itself has no CPU implementation (BASELINE.json north_star).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
// product (gsdf_slam_b200/) never does.
//
// PINNING: the reference has no tests, golden vectors or fixtures for this path
// (SURVEY.md §4, §8c).  This oracle is therefore pinned against OUTPUTS OF THE REFERENCE
// ITSELF: tests/golden/*.npz were produced by the unmodified reference CUDA code compiled
// here (oracle/_ref, oracle/build_ref.sh) and run on a B200 by tests/golden/make_golden.py;
// tests/test_oracle_vs_golden.py checks this file against them.  CPU libm exp/log differ
// from CUDA's by <= 2 ulp, so discrete outputs (n_contrib) are compared with a small
// mismatch allowance and images/gradients with tolerances; the BIT-EXACT gates are held by
// the GPU oracle (oracle/_ref) in tests/test_gpu_parity.py.
//
// Each function cites the reference lines it restates (paths under /root/reference).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int BX = 16, BY = 16;  // cuda_rasterizer/config.h:15-16

const float SH_C0 = 0.28209479177387814f;
const float SH_C1 = 0.4886025119029199f;
const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                       -1.0925484305920792f, 0.5462742152960396f};
const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                       0.3731763325901154f,  -0.4570457994644658f, 1.445305721320277f,
                       -0.5900435899266435f};

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// column-major 3x3 like glm: m[c][r]
struct M3 {
  float m[3][3];
  float* operator[](int c) { return m[c]; }
  const float* operator[](int c) const { return m[c]; }
};
inline M3 mk(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
  M3 r; r[0][0]=a; r[0][1]=b; r[0][2]=c; r[1][0]=d; r[1][1]=e; r[1][2]=f; r[2][0]=g; r[2][1]=h; r[2][2]=i;
  return r;
}
inline M3 tr(const M3& a) {
  return mk(a[0][0], a[1][0], a[2][0], a[0][1], a[1][1], a[2][1], a[0][2], a[1][2], a[2][2]);
}
inline M3 mul(const M3& a, const M3& b) {
  M3 r;
  for (int c = 0; c < 3; ++c)
    for (int q = 0; q < 3; ++q) r[c][q] = a[0][q] * b[c][0] + a[1][q] * b[c][1] + a[2][q] * b[c][2];
  return r;
}

// auxiliary.h:68-88
inline V3 xf43(V3 p, const float* m) {
  return {m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12], m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
          m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]};
}
inline void xf44(V3 p, const float* m, float out[4]) {
  out[0] = m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12];
  out[1] = m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13];
  out[2] = m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14];
  out[3] = m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15];
}
// auxiliary.h:33-36 (double arithmetic)
inline float ndc2pix(float v, int S) { return (float)(((v + 1.0) * S - 1.0) * 0.5); }

// auxiliary.h:38-51
inline void get_rect(float px, float py, int r, int gx, int gy, int& x0, int& y0, int& x1, int& y1) {
  x0 = std::min(gx, std::max(0, (int)((px - r) / BX)));
  y0 = std::min(gy, std::max(0, (int)((py - r) / BY)));
  x1 = std::min(gx, std::max(0, (int)((px + r + BX - 1) / BX)));
  y1 = std::min(gy, std::max(0, (int)((py + r + BY - 1) / BY)));
}

inline float sat(float x) { return std::min(1.0f, std::max(0.0f, x)); }

// rasterizer_impl.cu:59-118 (max_contrib_power_rect_gaussian_float<15,15> + evaluate_opacity_factor)
inline float tile_min_power(const float co[4], float mx, float my, int tx, int ty) {
  const float rminx = (float)(tx * BX), rminy = (float)(ty * BY);
  const float rmaxx = (float)((tx + 1) * BX - 1), rmaxy = (float)((ty + 1) * BY - 1);
  const float x_min_diff = rminx - mx;
  const float x_left = x_min_diff > 0.0f;
  const float not_in_x = x_left + (mx > rmaxx);
  const float y_min_diff = rminy - my;
  const float y_above = y_min_diff > 0.0f;
  const float not_in_y = y_above + (my > rmaxy);
  float res = 0.0f;
  if ((not_in_y + not_in_x) > 0.0f) {
    const float px = x_left * rminx + (1.0f - x_left) * rmaxx;
    const float py = y_above * rminy + (1.0f - y_above) * rmaxy;
    const float dx = std::copysign(15.0f, x_min_diff);
    const float dy = std::copysign(15.0f, y_min_diff);
    const float diffx = mx - px, diffy = my - py;
    const float rx = 1.0f / (225.0f * co[0]);
    const float ry = 1.0f / (225.0f * co[2]);
    const float tx_ = not_in_y * sat((dx * co[0] * diffx + dx * co[1] * diffy) * rx);
    const float ty_ = not_in_x * sat((dy * co[1] * diffx + dy * co[2] * diffy) * ry);
    const float qx = px + tx_ * dx, qy = py + ty_ * dy;
    const float ex = mx - qx, ey = my - qy;
    res = 0.5f * (co[0] * ex * ex + co[2] * ey * ey) + co[1] * ex * ey;
  }
  return res;
}

struct State {
  int P = 0, W = 0, H = 0, gx = 0, gy = 0;
  std::vector<float> depths, means2D, conic_opacity, rgb, cov3D;
  std::vector<uint8_t> clamped;
  std::vector<int> radii;
  std::vector<uint32_t> tiles;
  std::vector<uint64_t> keys;
  std::vector<uint32_t> ids;
  std::vector<uint32_t> ranges;  // 2*T
  std::vector<float> final_T, accum_c;
  std::vector<uint32_t> n_contrib;
};

// forward.cu:154-193
void cov3d(const float* s, float mod, const float* q, float* out) {
  M3 S = mk(1, 0, 0, 0, 1, 0, 0, 0, 1);
  S[0][0] = mod * s[0]; S[1][1] = mod * s[1]; S[2][2] = mod * s[2];
  const float r = q[0], x = q[1], y = q[2], z = q[3];
  M3 R = mk(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
            2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
            2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
  M3 Mm = mul(S, R);
  M3 Sg = mul(tr(Mm), Mm);
  out[0] = Sg[0][0]; out[1] = Sg[0][1]; out[2] = Sg[0][2]; out[3] = Sg[1][1]; out[4] = Sg[1][2]; out[5] = Sg[2][2];
}

struct Cov2DTmp { M3 T, Vrk, Wm; float tx, ty, tz, txtz, tytz, limx, limy; };

// forward.cu:104-149
void cov2d(V3 mean, float fx, float fy, float tanx, float tany, const float* c3, const float* view,
           float out[3], Cov2DTmp* tmp = nullptr) {
  V3 t = xf43(mean, view);
  const float limx = 1.3f * tanx, limy = 1.3f * tany;
  const float txtz = t.x / t.z, tytz = t.y / t.z;
  t.x = std::min(limx, std::max(-limx, txtz)) * t.z;
  t.y = std::min(limy, std::max(-limy, tytz)) * t.z;
  M3 J = mk(fx / t.z, 0.0f, -(fx * t.x) / (t.z * t.z), 0.0f, fy / t.z, -(fy * t.y) / (t.z * t.z), 0, 0, 0);
  M3 Wm = mk(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
  M3 T = mul(Wm, J);
  M3 Vrk = mk(c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]);
  M3 cov = mul(mul(tr(T), tr(Vrk)), T);
  out[0] = cov[0][0] + 0.3f; out[1] = cov[0][1]; out[2] = cov[1][1] + 0.3f;
  if (tmp) { tmp->T = T; tmp->Vrk = Vrk; tmp->Wm = Wm; tmp->tx = t.x; tmp->ty = t.y; tmp->tz = t.z;
             tmp->txtz = txtz; tmp->tytz = tytz; tmp->limx = limx; tmp->limy = limy; }
}

// forward.cu:32-101, one channel
float sh_eval(int deg, float x, float y, float z, float dc0, const float* sh /*stride 3*/) {
  float r = SH_C0 * dc0;
  if (deg > 0) {
    r = r - SH_C1 * y * sh[0] + SH_C1 * z * sh[3] - SH_C1 * x * sh[6];
    if (deg > 1) {
      float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      r = r + SH_C2[0] * xy * sh[9] + SH_C2[1] * yz * sh[12] + SH_C2[2] * (2.0f * zz - xx - yy) * sh[15] +
          SH_C2[3] * xz * sh[18] + SH_C2[4] * (xx - yy) * sh[21];
      if (deg > 2) {
        r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[24] + SH_C3[1] * xy * z * sh[27] +
            SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[30] +
            SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[33] +
            SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[36] + SH_C3[5] * z * (xx - yy) * sh[39] +
            SH_C3[6] * x * (xx - 3.0f * yy) * sh[42];
      }
    }
  }
  return r + 0.5f;
}

inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

}  // namespace

extern "C" {

void* oracle_create() { return new State(); }
void oracle_destroy(void* h) { delete static_cast<State*>(h); }
int oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}

// Rasterizer::forward (rasterizer_impl.cu:366-538): preprocess (forward.cu:197-312),
// duplicateWithKeys (rasterizer_impl.cu:139-212; only the valid keys are generated),
// stable sort on (tile, depth bits) (== cub SortPairs on bits [0, 32+bit), :486-491),
// identifyTileRanges (:217-246) and renderCUDA (forward.cu:317-477).
int oracle_forward(void* hh, int P, int D, int M, const float* bg, int W, int H, const float* means3D,
                   const float* dc, const float* shs, const float* colors_precomp, const float* opac,
                   const float* scales, float mod, const float* rots, const float* cov3D_precomp,
                   const float* view, const float* proj, const float* campos, float tanx, float tany,
                   float* out_color, int* radii_out, long long* R_out) {
  State& S = *static_cast<State*>(hh);
  const float fy = H / (2.0f * tany), fx = W / (2.0f * tanx);
  const int gx = (W + BX - 1) / BX, gy = (H + BY - 1) / BY, T = gx * gy;
  const size_t N = (size_t)W * H;
  S.P = P; S.W = W; S.H = H; S.gx = gx; S.gy = gy;
  S.depths.assign(P, 0.f); S.means2D.assign(2 * (size_t)P, 0.f); S.conic_opacity.assign(4 * (size_t)P, 0.f);
  S.rgb.assign(3 * (size_t)P, 0.f); S.cov3D.assign(6 * (size_t)P, 0.f); S.clamped.assign(3 * (size_t)P, 0);
  S.radii.assign(P, 0); S.tiles.assign(P, 0);
  const int deg = (shs && M > 0) ? std::min(D, M >= 15 ? 3 : (M >= 8 ? 2 : (M >= 3 ? 1 : 0))) : 0;

#pragma omp parallel for schedule(dynamic, 512)
  for (int i = 0; i < P; ++i) {
    const V3 p = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
    const V3 pv = xf43(p, view);
    if (pv.z <= 0.2f) continue;  // auxiliary.h:160-194
    float ph[4];
    xf44(p, proj, ph);
    const float pw = 1.0f / (ph[3] + 0.0000001f);
    const float prx = ph[0] * pw, pry = ph[1] * pw;
    float* c3 = &S.cov3D[6 * (size_t)i];
    if (cov3D_precomp) std::memcpy(c3, cov3D_precomp + 6 * (size_t)i, 24);
    else cov3d(scales + 3 * (size_t)i, mod, rots + 4 * (size_t)i, c3);
    float cov[3];
    cov2d(p, fx, fy, tanx, tany, c3, view, cov);
    const float det = cov[0] * cov[2] - cov[1] * cov[1];
    if (det == 0.0f) continue;
    const float di = 1.f / det;
    const float conic[3] = {cov[2] * di, -cov[1] * di, cov[0] * di};
    const float mid = 0.5f * (cov[0] + cov[2]);
    const float l1 = mid + std::sqrt(std::max(0.1f, mid * mid - det));
    const float l2 = mid - std::sqrt(std::max(0.1f, mid * mid - det));
    const float rad = std::ceil(3.f * std::sqrt(std::max(l1, l2)));
    const float px = ndc2pix(prx, W), py = ndc2pix(pry, H);
    int x0, y0, x1, y1;
    get_rect(px, py, (int)rad, gx, gy, x0, y0, x1, y1);
    if ((x1 - x0) * (y1 - y0) == 0) continue;
    if (!colors_precomp) {
      V3 dir = p - V3{campos[0], campos[1], campos[2]};
      const float len = std::sqrt(dot(dir, dir));
      dir = {dir.x / len, dir.y / len, dir.z / len};
      const float* sh = shs ? shs + (size_t)i * M * 3 : nullptr;
      for (int c = 0; c < 3; ++c) {
        float v = sh_eval(deg, dir.x, dir.y, dir.z, dc[3 * (size_t)i + c], sh ? sh + c : nullptr);
        S.clamped[3 * (size_t)i + c] = v < 0;
        S.rgb[3 * (size_t)i + c] = std::max(v, 0.0f);
      }
    } else {
      for (int c = 0; c < 3; ++c) S.rgb[3 * (size_t)i + c] = colors_precomp[3 * (size_t)i + c];
    }
    S.depths[i] = pv.z;
    S.radii[i] = (int)rad;
    S.means2D[2 * (size_t)i] = px; S.means2D[2 * (size_t)i + 1] = py;
    float* co = &S.conic_opacity[4 * (size_t)i];
    co[0] = conic[0]; co[1] = conic[1]; co[2] = conic[2]; co[3] = opac[i];
    // exact tile count = number of keys duplicateWithKeys keeps (rasterizer_impl.cu:169-199)
    const float thr = std::log(co[3] / (1.0f / 255.0f));
    uint32_t cnt = 0;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) cnt += tile_min_power(co, px, py, x, y) <= thr;
    S.tiles[i] = cnt;
  }
  if (radii_out) std::memcpy(radii_out, S.radii.data(), sizeof(int) * (size_t)P);

  // keys in Gaussian-index order, then stable sort by (tile, depth bits)
  std::vector<uint64_t> off(P + 1, 0);
  for (int i = 0; i < P; ++i) off[i + 1] = off[i] + S.tiles[i];
  const size_t R = off[P];
  std::vector<uint64_t> keys(R);
  std::vector<uint32_t> ids(R);
#pragma omp parallel for schedule(dynamic, 512)
  for (int i = 0; i < P; ++i) {
    if (S.tiles[i] == 0) continue;
    const float* co = &S.conic_opacity[4 * (size_t)i];
    const float px = S.means2D[2 * (size_t)i], py = S.means2D[2 * (size_t)i + 1];
    int x0, y0, x1, y1;
    get_rect(px, py, S.radii[i], gx, gy, x0, y0, x1, y1);
    const float thr = std::log(co[3] / (1.0f / 255.0f));
    size_t o = off[i];
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x)
        if (tile_min_power(co, px, py, x, y) <= thr) {
          keys[o] = ((uint64_t)(y * gx + x) << 32) | fbits(S.depths[i]);
          ids[o] = (uint32_t)i;
          ++o;
        }
  }
  std::vector<uint32_t> perm(R);
  std::iota(perm.begin(), perm.end(), 0u);
  std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  S.keys.resize(R); S.ids.resize(R);
  for (size_t k = 0; k < R; ++k) { S.keys[k] = keys[perm[k]]; S.ids[k] = ids[perm[k]]; }
  S.ranges.assign(2 * (size_t)T, 0);
  for (size_t k = 0; k < R; ++k) {
    const uint32_t t = (uint32_t)(S.keys[k] >> 32);
    if (k == 0 || (uint32_t)(S.keys[k - 1] >> 32) != t) S.ranges[2 * t] = (uint32_t)k;
    S.ranges[2 * t + 1] = (uint32_t)k + 1;
  }
  if (R_out) *R_out = (long long)R;

  // per-pixel front-to-back blend (forward.cu:378-464)
  S.final_T.assign(N, 1.f); S.n_contrib.assign(N, 0); S.accum_c.assign(3 * N, 0.f);
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < T; ++t) {
    const int tx = t % gx, ty = t / gx;
    const uint32_t r0 = S.ranges[2 * t], r1 = S.ranges[2 * t + 1];
    for (int ly = 0; ly < BY; ++ly)
      for (int lx = 0; lx < BX; ++lx) {
        const int px = tx * BX + lx, py = ty * BY + ly;
        if (px >= W || py >= H) continue;
        const float pxf = (float)px, pyf = (float)py;
        float Tt = 1.0f, C[3] = {0, 0, 0};
        uint32_t contributor = 0, last = 0;
        for (uint32_t k = r0; k < r1; ++k) {
          ++contributor;
          const uint32_t id = S.ids[k];
          const float* co = &S.conic_opacity[4 * (size_t)id];
          const float dx = S.means2D[2 * (size_t)id] - pxf, dy = S.means2D[2 * (size_t)id + 1] - pyf;
          const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
          if (power > 0.0f) continue;
          const float alpha = std::min(0.99f, co[3] * std::exp(power));
          if (alpha < 1.0f / 255.0f) continue;
          const float test_T = Tt * (1 - alpha);
          if (test_T < 0.0001f) break;
          for (int c = 0; c < 3; ++c) C[c] += S.rgb[3 * (size_t)id + c] * alpha * Tt;
          Tt = test_T;
          last = contributor;
        }
        const size_t pix = (size_t)py * W + px;
        S.final_T[pix] = Tt;
        S.n_contrib[pix] = last;
        for (int c = 0; c < 3; ++c) {
          S.accum_c[c * N + pix] = C[c];
          out_color[c * N + pix] = C[c] + Tt * bg[c];
        }
      }
  }
  return 0;
}

// Rasterizer::backward (rasterizer_impl.cu:542-623): per-pair gradient terms of
// PerGaussianRenderCUDA (backward.cu:597-637) accumulated per pixel front to back, then
// computeCov2DCUDA (backward.cu:159-313) and preprocessCUDA (backward.cu:402-461) with the
// SH (backward.cu:27-154) and cov3D (backward.cu:317-397) chains.
int oracle_backward(void* hh, int P, int D, int M, const float* bg, const float* means3D, const float* dc,
                    const float* shs, const float* colors_precomp, const float* scales, float mod,
                    const float* rots, const float* cov3D_precomp, const float* view, const float* proj,
                    const float* campos, float tanx, float tany, const float* dL_dpix, float* dL_dmean2D,
                    float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                    float* dL_dcov3D, float* dL_ddc, float* dL_dsh, float* dL_dscale, float* dL_drot) {
  State& S = *static_cast<State*>(hh);
  const int W = S.W, H = S.H, gx = S.gx, gy = S.gy, T = gx * gy;
  const size_t N = (size_t)W * H;
  const float fy = H / (2.0f * tany), fx = W / (2.0f * tanx);
  (void)dc; (void)colors_precomp;
  std::vector<double> a_m2(2 * (size_t)P, 0.0), a_con(3 * (size_t)P, 0.0), a_op(P, 0.0), a_col(3 * (size_t)P, 0.0);
  const float ddelx = 0.5f * W, ddely = 0.5f * H;

#pragma omp parallel
  {
    std::vector<double> l_m2(2 * (size_t)P, 0.0), l_con(3 * (size_t)P, 0.0), l_op(P, 0.0), l_col(3 * (size_t)P, 0.0);
#pragma omp for schedule(dynamic, 1) nowait
    for (int t = 0; t < T; ++t) {
      const int tx = t % gx, ty = t / gx;
      const uint32_t r0 = S.ranges[2 * t];
      for (int ly = 0; ly < BY; ++ly)
        for (int lx = 0; lx < BX; ++lx) {
          const int px = tx * BX + lx, py = ty * BY + ly;
          if (px >= W || py >= H) continue;
          const size_t pix = (size_t)py * W + px;
          const float pxf = (float)px, pyf = (float)py;
          const float T_final = S.final_T[pix];
          const uint32_t last = S.n_contrib[pix];
          float Tt = 1.0f, ar[3], dL[3];
          float bg_dot = 0.f;
          for (int c = 0; c < 3; ++c) {
            ar[c] = -S.accum_c[c * N + pix];
            dL[c] = dL_dpix[c * N + pix];
            bg_dot += bg[c] * dL[c];
          }
          for (uint32_t j = 0; j < last; ++j) {
            const uint32_t id = S.ids[r0 + j];
            const float* co = &S.conic_opacity[4 * (size_t)id];
            const float dx = S.means2D[2 * (size_t)id] - pxf, dy = S.means2D[2 * (size_t)id + 1] - pyf;
            const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
            if (power > 0.0f) continue;
            const float G = std::exp(power);
            const float alpha = std::min(0.99f, co[3] * G);
            if (alpha < 1.0f / 255.0f) continue;
            const float dch = alpha * Tt;
            float dL_dalpha = 0.f;
            for (int c = 0; c < 3; ++c) {
              const float col = S.rgb[3 * (size_t)id + c];
              ar[c] += Tt * alpha * col;
              l_col[3 * (size_t)id + c] += dch * dL[c];
              dL_dalpha += ((col * Tt) - (1.0f / (1.0f - alpha)) * (-ar[c])) * dL[c];
            }
            dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
            Tt *= (1.0f - alpha);
            const float dL_dG = co[3] * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * co[0] - gdy * co[1];
            const float dG_ddely = -gdy * co[2] - gdx * co[1];
            l_m2[2 * (size_t)id] += dL_dG * dG_ddelx * ddelx;
            l_m2[2 * (size_t)id + 1] += dL_dG * dG_ddely * ddely;
            l_con[3 * (size_t)id] += -0.5f * gdx * dx * dL_dG;
            l_con[3 * (size_t)id + 1] += -0.5f * gdx * dy * dL_dG;
            l_con[3 * (size_t)id + 2] += -0.5f * gdy * dy * dL_dG;
            l_op[id] += G * dL_dalpha;
          }
        }
    }
#pragma omp critical
    {
      for (size_t k = 0; k < a_m2.size(); ++k) a_m2[k] += l_m2[k];
      for (size_t k = 0; k < a_con.size(); ++k) a_con[k] += l_con[k];
      for (size_t k = 0; k < a_op.size(); ++k) a_op[k] += l_op[k];
      for (size_t k = 0; k < a_col.size(); ++k) a_col[k] += l_col[k];
    }
  }

  const int deg = (shs && M > 0) ? std::min(D, M >= 15 ? 3 : (M >= 8 ? 2 : (M >= 3 ? 1 : 0))) : 0;
#pragma omp parallel for schedule(dynamic, 512)
  for (int i = 0; i < P; ++i) {
    const size_t si = (size_t)i;
    float g_m2[3] = {(float)a_m2[2 * si], (float)a_m2[2 * si + 1], 0.f};
    const float dcon[3] = {(float)a_con[3 * si], (float)a_con[3 * si + 1], (float)a_con[3 * si + 2]};
    V3 dRGB = {(float)a_col[3 * si], (float)a_col[3 * si + 1], (float)a_col[3 * si + 2]};
    for (int c = 0; c < 3; ++c) { dL_dmean2D[3 * si + c] = g_m2[c]; dL_dmean3D[3 * si + c] = 0; dL_ddc[3 * si + c] = 0;
                                  dL_dscale[3 * si + c] = 0; }
    dL_dcolor[3 * si] = dRGB.x; dL_dcolor[3 * si + 1] = dRGB.y; dL_dcolor[3 * si + 2] = dRGB.z;
    if (dL_dconic) { dL_dconic[4 * si] = dcon[0]; dL_dconic[4 * si + 1] = dcon[1]; dL_dconic[4 * si + 2] = 0; dL_dconic[4 * si + 3] = dcon[2]; }
    dL_dopacity[si] = (float)a_op[si];
    for (int c = 0; c < 4; ++c) dL_drot[4 * si + c] = 0;
    if (dL_dcov3D) for (int c = 0; c < 6; ++c) dL_dcov3D[6 * si + c] = 0;
    if (dL_dsh) for (int c = 0; c < 3 * M; ++c) dL_dsh[3 * (size_t)M * si + c] = 0;
    if (!(S.radii[i] > 0)) continue;

    const V3 mean = {means3D[3 * si], means3D[3 * si + 1], means3D[3 * si + 2]};
    const float* c3 = cov3D_precomp ? cov3D_precomp + 6 * si : &S.cov3D[6 * si];
    float cov[3];
    Cov2DTmp q;
    cov2d(mean, fx, fy, tanx, tany, c3, view, cov, &q);
    const M3& Tm = q.T; const M3& Vrk = q.Vrk; const M3& Wm = q.Wm;
    const float a = cov[0], b = cov[1], c = cov[2];
    const float xg = (q.txtz < -q.limx || q.txtz > q.limx) ? 0.f : 1.f;
    const float yg = (q.tytz < -q.limy || q.tytz > q.limy) ? 0.f : 1.f;
    const float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float d2i = 1.0f / ((denom * denom) + 0.0000001f);
    float dcov[6] = {0, 0, 0, 0, 0, 0};
    if (d2i != 0) {
      dL_da = d2i * (-c * c * dcon[0] + 2 * b * c * dcon[1] + (denom - a * c) * dcon[2]);
      dL_dc = d2i * (-a * a * dcon[2] + 2 * a * b * dcon[1] + (denom - a * c) * dcon[0]);
      dL_db = d2i * 2 * (b * c * dcon[0] - (denom + 2 * b * b) * dcon[1] + a * b * dcon[2]);
      dcov[0] = Tm[0][0] * Tm[0][0] * dL_da + Tm[0][0] * Tm[1][0] * dL_db + Tm[1][0] * Tm[1][0] * dL_dc;
      dcov[3] = Tm[0][1] * Tm[0][1] * dL_da + Tm[0][1] * Tm[1][1] * dL_db + Tm[1][1] * Tm[1][1] * dL_dc;
      dcov[5] = Tm[0][2] * Tm[0][2] * dL_da + Tm[0][2] * Tm[1][2] * dL_db + Tm[1][2] * Tm[1][2] * dL_dc;
      dcov[1] = 2 * Tm[0][0] * Tm[0][1] * dL_da + (Tm[0][0] * Tm[1][1] + Tm[0][1] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][1] * dL_dc;
      dcov[2] = 2 * Tm[0][0] * Tm[0][2] * dL_da + (Tm[0][0] * Tm[1][2] + Tm[0][2] * Tm[1][0]) * dL_db + 2 * Tm[1][0] * Tm[1][2] * dL_dc;
      dcov[4] = 2 * Tm[0][2] * Tm[0][1] * dL_da + (Tm[0][1] * Tm[1][2] + Tm[0][2] * Tm[1][1]) * dL_db + 2 * Tm[1][1] * Tm[1][2] * dL_dc;
    }
    if (dL_dcov3D) for (int k = 0; k < 6; ++k) dL_dcov3D[6 * si + k] = dcov[k];
    auto rowdot = [&](int r, int vr) { return Tm[r][0] * Vrk[vr][0] + Tm[r][1] * Vrk[vr][1] + Tm[r][2] * Vrk[vr][2]; };
    const float dT00 = 2 * rowdot(0, 0) * dL_da + rowdot(1, 0) * dL_db;
    const float dT01 = 2 * rowdot(0, 1) * dL_da + rowdot(1, 1) * dL_db;
    const float dT02 = 2 * rowdot(0, 2) * dL_da + rowdot(1, 2) * dL_db;
    const float dT10 = 2 * rowdot(1, 0) * dL_dc + rowdot(0, 0) * dL_db;
    const float dT11 = 2 * rowdot(1, 1) * dL_dc + rowdot(0, 1) * dL_db;
    const float dT12 = 2 * rowdot(1, 2) * dL_dc + rowdot(0, 2) * dL_db;
    const float dJ00 = Wm[0][0] * dT00 + Wm[0][1] * dT01 + Wm[0][2] * dT02;
    const float dJ02 = Wm[2][0] * dT00 + Wm[2][1] * dT01 + Wm[2][2] * dT02;
    const float dJ11 = Wm[1][0] * dT10 + Wm[1][1] * dT11 + Wm[1][2] * dT12;
    const float dJ12 = Wm[2][0] * dT10 + Wm[2][1] * dT11 + Wm[2][2] * dT12;
    const float tz = 1.f / q.tz, tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = xg * -fx * tz2 * dJ02;
    const float dty = yg * -fy * tz2 * dJ12;
    const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * q.tx) * tz3 * dJ02 + (2 * fy * q.ty) * tz3 * dJ12;
    V3 dmean = {view[0] * dtx + view[1] * dty + view[2] * dtz, view[4] * dtx + view[5] * dty + view[6] * dtz,
                view[8] * dtx + view[9] * dty + view[10] * dtz};
    {  // backward.cu:427-449
      float mh[4];
      xf44(mean, proj, mh);
      const float mw = 1.0f / (mh[3] + 0.0000001f);
      const float mul1 = (proj[0] * mean.x + proj[4] * mean.y + proj[8] * mean.z + proj[12]) * mw * mw;
      const float mul2 = (proj[1] * mean.x + proj[5] * mean.y + proj[9] * mean.z + proj[13]) * mw * mw;
      V3 d;
      d.x = (proj[0] * mw - proj[3] * mul1) * g_m2[0] + (proj[1] * mw - proj[3] * mul2) * g_m2[1];
      d.y = (proj[4] * mw - proj[7] * mul1) * g_m2[0] + (proj[5] * mw - proj[7] * mul2) * g_m2[1];
      d.z = (proj[8] * mw - proj[11] * mul1) * g_m2[0] + (proj[9] * mw - proj[11] * mul2) * g_m2[1];
      dmean = dmean + d;
    }
    if (shs && M > 0) {  // backward.cu:27-154 (guarded by `if (shs)`, :452)
      const V3 dir_o = mean - V3{campos[0], campos[1], campos[2]};
      const float len = std::sqrt(dot(dir_o, dir_o));
      const V3 dir = {dir_o.x / len, dir_o.y / len, dir_o.z / len};
      const V3* sh = reinterpret_cast<const V3*>(shs) + si * M;
      V3* dsh = reinterpret_cast<V3*>(dL_dsh) + si * M;
      dRGB.x *= S.clamped[3 * si] ? 0 : 1; dRGB.y *= S.clamped[3 * si + 1] ? 0 : 1; dRGB.z *= S.clamped[3 * si + 2] ? 0 : 1;
      V3 dx = {0, 0, 0}, dy = {0, 0, 0}, dz = {0, 0, 0};
      const float x = dir.x, y = dir.y, z = dir.z;
      const V3 ddc = SH_C0 * dRGB;
      dL_ddc[3 * si] = ddc.x; dL_ddc[3 * si + 1] = ddc.y; dL_ddc[3 * si + 2] = ddc.z;
      if (deg > 0) {
        dsh[0] = (-SH_C1 * y) * dRGB; dsh[1] = (SH_C1 * z) * dRGB; dsh[2] = (-SH_C1 * x) * dRGB;
        dx = -SH_C1 * sh[2]; dy = -SH_C1 * sh[0]; dz = SH_C1 * sh[1];
        if (deg > 1) {
          const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
          dsh[3] = (SH_C2[0] * xy) * dRGB; dsh[4] = (SH_C2[1] * yz) * dRGB;
          dsh[5] = (SH_C2[2] * (2.f * zz - xx - yy)) * dRGB; dsh[6] = (SH_C2[3] * xz) * dRGB;
          dsh[7] = (SH_C2[4] * (xx - yy)) * dRGB;
          dx = dx + (SH_C2[0] * y) * sh[3] + (SH_C2[2] * 2.f * -x) * sh[5] + (SH_C2[3] * z) * sh[6] + (SH_C2[4] * 2.f * x) * sh[7];
          dy = dy + (SH_C2[0] * x) * sh[3] + (SH_C2[1] * z) * sh[4] + (SH_C2[2] * 2.f * -y) * sh[5] + (SH_C2[4] * 2.f * -y) * sh[7];
          dz = dz + (SH_C2[1] * y) * sh[4] + (SH_C2[2] * 2.f * 2.f * z) * sh[5] + (SH_C2[3] * x) * sh[6];
          if (deg > 2) {
            dsh[8] = (SH_C3[0] * y * (3.f * xx - yy)) * dRGB; dsh[9] = (SH_C3[1] * xy * z) * dRGB;
            dsh[10] = (SH_C3[2] * y * (4.f * zz - xx - yy)) * dRGB;
            dsh[11] = (SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dRGB;
            dsh[12] = (SH_C3[4] * x * (4.f * zz - xx - yy)) * dRGB; dsh[13] = (SH_C3[5] * z * (xx - yy)) * dRGB;
            dsh[14] = (SH_C3[6] * x * (xx - 3.f * yy)) * dRGB;
            dx = dx + (SH_C3[0] * 3.f * 2.f * xy) * sh[8] + (SH_C3[1] * yz) * sh[9] + (SH_C3[2] * -2.f * xy) * sh[10] +
                 (SH_C3[3] * -3.f * 2.f * xz) * sh[11] + (SH_C3[4] * (-3.f * xx + 4.f * zz - yy)) * sh[12] +
                 (SH_C3[5] * 2.f * xz) * sh[13] + (SH_C3[6] * 3.f * (xx - yy)) * sh[14];
            dy = dy + (SH_C3[0] * 3.f * (xx - yy)) * sh[8] + (SH_C3[1] * xz) * sh[9] +
                 (SH_C3[2] * (-3.f * yy + 4.f * zz - xx)) * sh[10] + (SH_C3[3] * -3.f * 2.f * yz) * sh[11] +
                 (SH_C3[4] * -2.f * xy) * sh[12] + (SH_C3[5] * -2.f * yz) * sh[13] + (SH_C3[6] * -3.f * 2.f * xy) * sh[14];
            dz = dz + (SH_C3[1] * xy) * sh[9] + (SH_C3[2] * 4.f * 2.f * yz) * sh[10] +
                 (SH_C3[3] * 3.f * (2.f * zz - xx - yy)) * sh[11] + (SH_C3[4] * 4.f * 2.f * xz) * sh[12] +
                 (SH_C3[5] * (xx - yy)) * sh[13];
          }
        }
      }
      const V3 ddir = {dot(dx, dRGB), dot(dy, dRGB), dot(dz, dRGB)};
      const V3 v = dir_o;  // dnormvdv, auxiliary.h:122-138
      const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
      const float inv = 1.0f / std::sqrt(sum2 * sum2 * sum2);
      V3 r;
      r.x = ((+sum2 - v.x * v.x) * ddir.x - v.y * v.x * ddir.y - v.z * v.x * ddir.z) * inv;
      r.y = (-v.x * v.y * ddir.x + (sum2 - v.y * v.y) * ddir.y - v.z * v.y * ddir.z) * inv;
      r.z = (-v.x * v.z * ddir.x - v.y * v.z * ddir.y + (sum2 - v.z * v.z) * ddir.z) * inv;
      dmean = dmean + r;
    }
    dL_dmean3D[3 * si] = dmean.x; dL_dmean3D[3 * si + 1] = dmean.y; dL_dmean3D[3 * si + 2] = dmean.z;
    if (scales && !cov3D_precomp) {  // backward.cu:317-397
      const float* qd = rots + 4 * si;
      const float r = qd[0], x = qd[1], y = qd[2], z = qd[3];
      M3 R = mk(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
      const float s[3] = {mod * scales[3 * si], mod * scales[3 * si + 1], mod * scales[3 * si + 2]};
      M3 Sm = mk(s[0], 0, 0, 0, s[1], 0, 0, 0, s[2]);
      M3 Mm = mul(Sm, R);
      M3 dSig = mk(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                   0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
      M3 M2;
      for (int cc = 0; cc < 3; ++cc) for (int k = 0; k < 3; ++k) M2[cc][k] = 2.0f * Mm[cc][k];
      M3 dM = mul(M2, dSig);
      M3 Rt = tr(R), dMt = tr(dM);
      for (int k = 0; k < 3; ++k)
        dL_dscale[3 * si + k] = Rt[k][0] * dMt[k][0] + Rt[k][1] * dMt[k][1] + Rt[k][2] * dMt[k][2];
      for (int k = 0; k < 3; ++k) { dMt[0][k] *= s[0]; dMt[1][k] *= s[1]; dMt[2][k] *= s[2]; }
      float* dq = dL_drot + 4 * si;
      dq[0] = 2 * z * (dMt[0][1] - dMt[1][0]) + 2 * y * (dMt[2][0] - dMt[0][2]) + 2 * x * (dMt[1][2] - dMt[2][1]);
      dq[1] = 2 * y * (dMt[1][0] + dMt[0][1]) + 2 * z * (dMt[2][0] + dMt[0][2]) + 2 * r * (dMt[1][2] - dMt[2][1]) - 4 * x * (dMt[2][2] + dMt[1][1]);
      dq[2] = 2 * x * (dMt[1][0] + dMt[0][1]) + 2 * r * (dMt[2][0] - dMt[0][2]) + 2 * z * (dMt[1][2] + dMt[2][1]) - 4 * y * (dMt[2][2] + dMt[0][0]);
      dq[3] = 2 * r * (dMt[0][1] - dMt[1][0]) + 2 * x * (dMt[2][0] + dMt[0][2]) + 2 * y * (dMt[1][2] + dMt[2][1]) - 4 * z * (dMt[1][1] + dMt[0][0]);
    }
  }
  return 0;
}

// ---- state getters ---------------------------------------------------------------------
long long oracle_num_rendered(void* hh) { return (long long)static_cast<State*>(hh)->keys.size(); }
void oracle_get_binning(void* hh, uint64_t* keys, uint32_t* ids) {
  State& S = *static_cast<State*>(hh);
  if (keys) std::memcpy(keys, S.keys.data(), 8 * S.keys.size());
  if (ids) std::memcpy(ids, S.ids.data(), 4 * S.ids.size());
}
void oracle_get_image(void* hh, float* final_T, uint32_t* n_contrib, uint32_t* ranges) {
  State& S = *static_cast<State*>(hh);
  if (final_T) std::memcpy(final_T, S.final_T.data(), 4 * S.final_T.size());
  if (n_contrib) std::memcpy(n_contrib, S.n_contrib.data(), 4 * S.n_contrib.size());
  if (ranges) std::memcpy(ranges, S.ranges.data(), 4 * S.ranges.size());
}
void oracle_get_geometry(void* hh, float* depths, float* means2D, float* conic_opacity, float* rgb,
                         uint32_t* tiles, uint8_t* clamped) {
  State& S = *static_cast<State*>(hh);
  if (depths) std::memcpy(depths, S.depths.data(), 4 * S.depths.size());
  if (means2D) std::memcpy(means2D, S.means2D.data(), 4 * S.means2D.size());
  if (conic_opacity) std::memcpy(conic_opacity, S.conic_opacity.data(), 4 * S.conic_opacity.size());
  if (rgb) std::memcpy(rgb, S.rgb.data(), 4 * S.rgb.size());
  if (tiles) std::memcpy(tiles, S.tiles.data(), 4 * S.tiles.size());
  if (clamped) std::memcpy(clamped, S.clamped.data(), S.clamped.size());
}

// torch::optim::Adam::step, single tensor (LibTorch adam.cpp; call site
// src/gaussian_mapper.cpp:465, options src/gaussian_model.cpp:319-352)
void oracle_adam(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1,
                 double beta2, double eps, long long step) {
  const double bc1 = 1 - std::pow(beta1, (double)step), bc2 = 1 - std::pow(beta2, (double)step);
  const float b1 = (float)beta1, b2 = (float)beta2, o1 = (float)(1 - beta1), o2 = (float)(1 - beta2);
  const float inv = (float)(1.0 / std::sqrt(bc2)), ss = (float)(-(lr / bc1)), e = (float)eps;
#pragma omp parallel for
  for (long long i = 0; i < n; ++i) {
    const float m2 = std::fmaf(o1, g[i], m[i] * b1);
    const float v2 = std::fmaf(o2, g[i] * g[i], v[i] * b2);
    const float d = std::sqrt(v2) * inv + e;
    p[i] = std::fmaf(ss, m2 / d, p[i]);
    m[i] = m2; v[i] = v2;
  }
}

// SimpleKNN::knn result (simple_knn.cu:147-183): mean of the 3 smallest squared distances,
// by brute force (the Morton/box machinery only prunes, it does not change the result).
void oracle_knn(int P, const float* pts, float* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < P; ++i) {
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int j = 0; j < P; ++j) {
      if (j == i) continue;
      const float dx = pts[3 * j] - pts[3 * i], dy = pts[3 * j + 1] - pts[3 * i + 1], dz = pts[3 * j + 2] - pts[3 * i + 2];
      float d = dx * dx + dy * dy + dz * dz;
      for (int k = 0; k < 3; ++k)
        if (best[k] > d) std::swap(best[k], d);
    }
    out[i] = (best[0] + best[1] + best[2]) / 3.0f;
  }
}

// checkFrustum (rasterizer_impl.cu:122-135)
void oracle_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
  for (int i = 0; i < P; ++i) {
    const V3 p = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
    present[i] = xf43(p, view).z > 0.2f;
  }
}


// colorize_kernel + the compaction of colorize_launcher, cuda_colorize/colorize_pointcloud.cu:3-107.
// points [n,3] (LiDAR frame), image [H,W,3] float, proj row-major 3x4, Tr / pose row-major 4x4.
// Writes the surviving points (world frame) and their colours compacted in input order; returns the count.
static inline void oracle_transform_point(const float* m, const float* p, float* o) {
  o[0] = m[0] * p[0] + m[1] * p[1] + m[2] * p[2] + m[3];
  o[1] = m[4] * p[0] + m[5] * p[1] + m[6] * p[2] + m[7];
  o[2] = m[8] * p[0] + m[9] * p[1] + m[10] * p[2] + m[11];
}
long long oracle_colorize(long long n, const float* points, const float* pose, const float* proj,
                          const float* Tr, const float* image, int W, int H, float* points_out,
                          float* colors_out) {
  long long k = 0;
  for (long long i = 0; i < n; ++i) {
    const float* p = points + 3 * i;
    float cam[3], pr[3], world[3];
    oracle_transform_point(Tr, p, cam);     // :31
    oracle_transform_point(proj, cam, pr);  // :32
    if (cam[2] <= 0.2f) continue;           // :33-37
    const int px = static_cast<int>(pr[0] / pr[2]);  // :40-41
    const int py = static_cast<int>(pr[1] / pr[2]);
    if (px < 0 || px >= W || py < 0 || py >= H) continue;  // :43-47
    oracle_transform_point(pose, p, world);  // :50
    for (int c = 0; c < 3; ++c) {
      points_out[3 * k + c] = world[c];
      colors_out[3 * k + c] = image[3 * ((size_t)py * W + px) + c];  // :49
    }
    ++k;
  }
  return k;
}
}  // extern "C"
