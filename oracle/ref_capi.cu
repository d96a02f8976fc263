// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI wrapper around the UNMODIFIED
// reference CUDA rasterizer, compiled from the sources where they lie under
// /root/reference (see oracle/build_ref.sh).  It is the "GPU oracle": the
// bit-exact parity tests feed identical inputs to this library and to the
// product library (gsdf_slam_b200/csrc) and compare.  It is also the
// "reference CUDA recompiled for the same B200" baseline of bench.py
// (--impl reference).  Nothing in the product path links or loads it.
//
// Wraps: CudaRasterizer::Rasterizer::{forward,backward,markVisible}
//        (/root/reference/cuda_rasterizer/rasterizer.h:20-94) and
//        SimpleKNN::knn (/root/reference/submodules/simple-knn/simple_knn.h).
// The std::function<char*(size_t)> scratch callbacks that the reference's
// torch shim implements with Tensor::resize_ (src/rasterize_points.cu:28-36)
// are implemented here with grow-only cudaMalloc buffers held in a handle.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <tuple>

#include "rasterizer.h"
#include "rasterizer_impl.h"
#include "simple_knn.h"

namespace {
struct Buf {
  char* p = nullptr;
  size_t cap = 0;
  size_t size = 0;
  char* resize(size_t n) {
    if (n > cap) {
      if (p) cudaFree(p);
      size_t want = n + n / 4 + 256;
      if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cap = 0; return nullptr; }
      cap = want;
    }
    size = n;
    return p;
  }
  ~Buf() { if (p) cudaFree(p); }
};
struct Handle {
  Buf geom, binning, img, sample;
  int P = 0, R = 0, B = 0, W = 0, H = 0;
};
}  // namespace

extern "C" {

void* ref_create() { return new Handle(); }
void ref_destroy(void* h) { delete static_cast<Handle*>(h); }

// Returns 0 on success; -1 on exception / CUDA error.
int ref_forward(void* hh, int P, int D, int M, const float* background, int W,
                int H, const float* means3D, const float* dc, const float* shs,
                const float* colors_precomp, const float* opacities,
                const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix,
                const float* campos, float tan_fovx, float tan_fovy,
                int prefiltered, float* out_color, int* radii, int debug,
                int* num_rendered, int* num_buckets) {
  Handle* h = static_cast<Handle*>(hh);
  try {
    auto tup = CudaRasterizer::Rasterizer::forward(
        [h](size_t n) { return h->geom.resize(n); },
        [h](size_t n) { return h->binning.resize(n); },
        [h](size_t n) { return h->img.resize(n); },
        [h](size_t n) { return h->sample.resize(n); }, P, D, M, background, W,
        H, means3D, dc, shs, colors_precomp, opacities, scales, scale_modifier,
        rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx,
        tan_fovy, prefiltered != 0, out_color, radii, debug != 0);
    h->P = P; h->W = W; h->H = H;
    h->R = std::get<0>(tup);
    h->B = std::get<1>(tup);
    *num_rendered = h->R;
    *num_buckets = h->B;
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_forward: %s\n", e.what());
    return -1;
  }
  return 0;
}

int ref_backward(void* hh, int P, int D, int M, int R, int B,
                 const float* background, int W, int H, const float* means3D,
                 const float* dc, const float* shs, const float* colors_precomp,
                 const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp,
                 const float* viewmatrix, const float* projmatrix,
                 const float* campos, float tan_fovx, float tan_fovy,
                 const int* radii, const float* dL_dpix, float* dL_dmean2D,
                 float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                 float* dL_dmean3D, float* dL_dcov3D, float* dL_ddc,
                 float* dL_dsh, float* dL_dscale, float* dL_drot, int debug) {
  Handle* h = static_cast<Handle*>(hh);
  try {
    CudaRasterizer::Rasterizer::backward(
        P, D, M, R, B, background, W, H, means3D, dc, shs, colors_precomp,
        scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
        projmatrix, campos, tan_fovx, tan_fovy, radii, h->geom.p, h->binning.p,
        h->img.p, h->sample.p, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity,
        dL_dcolor, dL_dmean3D, dL_dcov3D, dL_ddc, dL_dsh, dL_dscale, dL_drot,
        debug != 0);
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_backward: %s\n", e.what());
    return -1;
  }
  return 0;
}

// ---- decoders of the reference's opaque byte buffers (device -> device) ----
#define D2D(dst, src, n)                                                     \
  if ((dst) != nullptr &&                                                    \
      cudaMemcpy((dst), (src), (n), cudaMemcpyDeviceToDevice) != cudaSuccess) \
    return -1;

int ref_get_geometry(void* hh, float* depths, float* means2D,
                     float* conic_opacity, float* rgb, uint32_t* tiles_touched,
                     uint8_t* clamped, float* cov3D, uint32_t* point_offsets) {
  Handle* h = static_cast<Handle*>(hh);
  char* chunk = h->geom.p;
  size_t P = h->P;
  auto g = CudaRasterizer::GeometryState::fromChunk(chunk, P);
  D2D(depths, g.depths, P * 4);
  D2D(means2D, g.means2D, P * 8);
  D2D(conic_opacity, g.conic_opacity, P * 16);
  D2D(rgb, g.rgb, P * 12);
  D2D(tiles_touched, g.tiles_touched, P * 4);
  D2D(clamped, g.clamped, P * 3);
  D2D(cov3D, g.cov3D, P * 24);
  D2D(point_offsets, g.point_offsets, P * 4);
  return 0;
}

int ref_get_binning(void* hh, uint64_t* keys_sorted, uint32_t* vals_sorted,
                    uint64_t* keys_unsorted, uint32_t* vals_unsorted) {
  Handle* h = static_cast<Handle*>(hh);
  char* chunk = h->binning.p;
  size_t R = h->R;
  if (R == 0) return 0;
  auto b = CudaRasterizer::BinningState::fromChunk(chunk, R);
  D2D(keys_sorted, b.point_list_keys, R * 8);
  D2D(vals_sorted, b.point_list, R * 4);
  D2D(keys_unsorted, b.point_list_keys_unsorted, R * 8);
  D2D(vals_unsorted, b.point_list_unsorted, R * 4);
  return 0;
}

// ranges/max_contrib/bucket_offsets: T = tiles entries are meaningful.
int ref_get_image(void* hh, float* final_T, uint32_t* n_contrib,
                  uint32_t* ranges /*2*T*/, uint32_t* max_contrib /*T*/,
                  uint32_t* bucket_offsets /*T*/) {
  Handle* h = static_cast<Handle*>(hh);
  char* chunk = h->img.p;
  size_t N = (size_t)h->W * h->H;
  size_t T = (size_t)((h->W + 15) / 16) * ((h->H + 15) / 16);
  auto s = CudaRasterizer::ImageState::fromChunk(chunk, N);
  D2D(final_T, s.accum_alpha, N * 4);
  D2D(n_contrib, s.n_contrib, N * 4);
  D2D(ranges, s.ranges, T * 8);
  D2D(max_contrib, s.max_contrib, T * 4);
  D2D(bucket_offsets, s.bucket_offsets, T * 4);
  return 0;
}

int ref_mark_visible(int P, float* means3D, float* viewmatrix,
                     float* projmatrix, uint8_t* present) {
  CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix,
                                          reinterpret_cast<bool*>(present));
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int ref_knn(int P, float* points, float* mean_dists) {
  SimpleKNN::knn(P, reinterpret_cast<float3*>(points), mean_dists);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
