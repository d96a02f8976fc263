// Drop-in replacement of the reference's torch<->CUDA shim header
// (/root/reference/include/rasterize_points.h:18-39, 62-87, 112-115): the SAME C++ symbols
// and signatures, so src/gaussian_rasterizer.cpp and src/gaussian_renderer.cpp of the
// reference compile and link unchanged against libgsr_torch_shim.so + libgsr.so.
// LibTorch is only the tensor container here; all arithmetic is behind the C ABI
// (include/gsr_capi.h).
#pragma once
#include <torch/torch.h>

#include <tuple>

std::tuple<int, int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                       const torch::Tensor& colors, const torch::Tensor& opacity,
                       const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                       const float tan_fovx, const float tan_fovy, const int image_height,
                       const int image_width, const torch::Tensor& dc, const torch::Tensor& sh,
                       const int degree, const torch::Tensor& campos, const bool prefiltered,
                       const bool debug);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                               const torch::Tensor& radii, const torch::Tensor& colors,
                               const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp,
                               const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                               const float tan_fovx, const float tan_fovy,
                               const torch::Tensor& dL_dout_color, const torch::Tensor& dc,
                               const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                               const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
                               const int B, const torch::Tensor& sampleBuffer, const bool debug);

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix);

// reference: submodules/simple-knn/spatial.h:14
torch::Tensor distCUDA2(const torch::Tensor& points);

// Fused Adam over parameter / gradient / moment tensors (replaces optimizer_->step(),
// src/gaussian_mapper.cpp:465, for callers that adopt it).  `steps[i]` is the step count after
// this update.
void FusedAdamStep(std::vector<torch::Tensor>& params, const std::vector<torch::Tensor>& grads,
                   std::vector<torch::Tensor>& exp_avgs, std::vector<torch::Tensor>& exp_avg_sqs,
                   const std::vector<double>& lrs, const std::vector<int64_t>& steps, double beta1,
                   double beta2, double eps);

// reference: cuda_colorize/colorize_pointcloud.h:4-11.  points / colors are re-bound to
// the compacted survivors, as the reference does.
void colorize_launcher(torch::Tensor& points, torch::Tensor& colors, torch::Tensor& proj, torch::Tensor& Tr,
                       torch::Tensor& pose, torch::Tensor& image);
