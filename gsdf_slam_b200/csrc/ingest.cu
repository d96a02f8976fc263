// Frame ingest: colourise a LiDAR scan from the camera image and keep the points
// that project into it.  Replaces colorize_kernel + colorize_launcher
// (cuda_colorize/colorize_pointcloud.cu:14-107): the reference writes a flag per point, calls
// cudaDeviceSynchronize, then builds the result with nonzero + two index_select passes over
// uncompacted arrays.  Here: one flag pass, this library's scan (gsr_select_rows), one pass that
// transforms + colours ONLY the surviving points straight into the compacted outputs; no device-wide
// synchronisation (the count is a 4-byte stream-ordered read-back).
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr_capi.h"
#include "gsr_common.cuh"

namespace gsr {
namespace {

// the reference's transformPoint (:3-12), same expression so that nvcc contracts it the same way
__device__ __forceinline__ float3 transform_point(const float3& p, const float* __restrict__ m) {
  float3 t = {
      m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3],
      m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
      m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11],
  };
  return t;
}

// pixel the point projects to, or -1 (colorize_kernel :32-48)
__device__ __forceinline__ int project_pixel(const float3& p, const float* __restrict__ Tr,
                                             const float* __restrict__ proj, int W, int H) {
  const float3 p_cam = transform_point(p, Tr);
  const float3 p_proj = transform_point(p_cam, proj);
  if (p_cam.z <= 0.2f) return -1;
  const int px = static_cast<int>(p_proj.x / p_proj.z);
  const int py = static_cast<int>(p_proj.y / p_proj.z);
  if (px < 0 || px >= W || py < 0 || py >= H) return -1;
  return py * W + px;
}

__global__ void __launch_bounds__(256)
colorize_flags_kernel(const float* __restrict__ points, int64_t n, const float* __restrict__ Tr,
                      const float* __restrict__ proj, int W, int H, uint8_t* __restrict__ valid) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float3 p = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
  valid[i] = project_pixel(p, Tr, proj, W, H) >= 0 ? 1 : 0;
}

__global__ void __launch_bounds__(256)
colorize_gather_kernel(const float* __restrict__ points, const uint32_t* __restrict__ idx,
                       const uint32_t* __restrict__ count, const float* __restrict__ pose,
                       const float* __restrict__ Tr, const float* __restrict__ proj,
                       const float* __restrict__ image, int W, int H, float* __restrict__ points_out,
                       float* __restrict__ colors_out) {
  const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (j >= (int64_t)*count) return;
  const size_t i = idx[j];
  const float3 p = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
  const int pix = project_pixel(p, Tr, proj, W, H);  // >= 0 by construction
  const float3 w = transform_point(p, pose);
  points_out[3 * j] = w.x;
  points_out[3 * j + 1] = w.y;
  points_out[3 * j + 2] = w.z;
  colors_out[3 * j] = image[3 * (size_t)pix];
  colors_out[3 * j + 1] = image[3 * (size_t)pix + 1];
  colors_out[3 * j + 2] = image[3 * (size_t)pix + 2];
}

}  // namespace
}  // namespace gsr

namespace gsr {
namespace {
// point records with a stride (KITTI .bin: x, y, z, intensity) -> dense [n,3]
__global__ void __launch_bounds__(256)
pack_xyz_kernel(const float* __restrict__ src, int64_t n, int stride, float* __restrict__ dst) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;   // one thread per output float: coalesced writes
  if (e >= 3 * n) return;
  const int64_t i = e / 3;
  dst[e] = src[i * stride + (e - 3 * i)];
}
}  // namespace
}  // namespace gsr

using namespace gsr;

extern "C" {

int gsr_pack_xyz(const float* records, int64_t n, int stride_floats, float* xyz_out, void* stream) {
  if (n < 0 || stride_floats < 3) return GSR_ERR_BAD_ARG;
  if (n == 0) return GSR_OK;
  if (!records || !xyz_out) return GSR_ERR_BAD_ARG;
  pack_xyz_kernel<<<(unsigned)((3 * n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(records, n, stride_floats,
                                                                                                 xyz_out);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

size_t gsr_colorize_workspace_bytes(int64_t n) {
  const size_t a = ((size_t)(n > 0 ? n : 0) + 255) / 256 * 256;  // flags
  return a + a * 4 + 256 + gsr_store_workspace_bytes(n);          // + index list + count + scan scratch
}

int gsr_colorize_points(const float* points, int64_t n, const float* lidar_pose, const float* proj_mat,
                        const float* Tr_velo_to_cam, const float* image, int img_width, int img_height,
                        float* points_out, float* colors_out, char* workspace, int64_t* count_out_host,
                        void* stream) {
  if (n < 0 || n >= (1ll << 31) || img_width <= 0 || img_height <= 0) return GSR_ERR_BAD_ARG;
  if (count_out_host) *count_out_host = 0;
  if (n == 0) return GSR_OK;
  if (!points || !lidar_pose || !proj_mat || !Tr_velo_to_cam || !image || !points_out || !colors_out || !workspace)
    return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t a = ((size_t)n + 255) / 256 * 256;
  uint8_t* valid = reinterpret_cast<uint8_t*>(workspace);
  uint32_t* idx = reinterpret_cast<uint32_t*>(workspace + a);
  uint32_t* count = reinterpret_cast<uint32_t*>(workspace + a + a * 4);
  char* scan_ws = workspace + a + a * 4 + 256;
  const unsigned blocks = (unsigned)((n + 255) / 256);
  colorize_flags_kernel<<<blocks, 256, 0, s>>>(points, n, Tr_velo_to_cam, proj_mat, img_width, img_height, valid);
  GSR_LAUNCHED();
  int rc = gsr_select_rows(valid, n, 0, idx, count, scan_ws, nullptr, stream);
  if (rc) return rc;
  colorize_gather_kernel<<<blocks, 256, 0, s>>>(points, idx, count, lidar_pose, Tr_velo_to_cam, proj_mat, image,
                                                img_width, img_height, points_out, colors_out);
  GSR_LAUNCHED();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (count_out_host) {
    uint32_t h = 0;
    e = cudaMemcpyAsync(&h, count, sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return (int)e;
    *count_out_host = h;
  }
  return GSR_OK;
}

}  // extern "C"
