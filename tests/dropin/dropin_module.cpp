// Test harness of the drop-in boundary: the reference's UNMODIFIED
// autograd glue — include/gaussian_rasterizer.h + src/gaussian_rasterizer.cpp (GaussianRasterizationSettings,
// GaussianRasterizerFunction::forward/backward with its 16 saved tensors, GaussianRasterizer::forward) — compiled
// where it lies and linked against THIS repository's libgsr_torch_shim.so instead of the reference's
// src/rasterize_points.cu.  The functions below call it the way GaussianRenderer::render does
// (/root/reference/src/gaussian_renderer.cpp:44-70, 119-130) and hand the results to Python.
#include <torch/extension.h>

#include <vector>

#include "include/gaussian_rasterizer.h"

// -> {color, radii, dL_dmeans3D, dL_dmeans2D (viewspace points), dL_ddc, dL_dsh, dL_dopacity, dL_dscales, dL_drotations}
static std::vector<torch::Tensor> render_forward_backward(
    torch::Tensor means3D_in, torch::Tensor opacity_in, torch::Tensor dc_in, torch::Tensor sh_in,
    torch::Tensor scales_in, torch::Tensor rotations_in, torch::Tensor bg, torch::Tensor viewmatrix,
    torch::Tensor projmatrix, torch::Tensor campos, int64_t H, int64_t W, double tanfovx, double tanfovy,
    int64_t sh_degree, torch::Tensor dL_dcolor, bool debug) {
  GaussianRasterizationSettings raster_settings((int)H, (int)W, (float)tanfovx, (float)tanfovy, bg, 1.0f, viewmatrix,
                                                projmatrix, (int)sh_degree, campos, false, debug);
  GaussianRasterizer rasterizer(raster_settings);
  auto leaf = [](const torch::Tensor& t) { return t.detach().clone().requires_grad_(true); };
  torch::Tensor means3D = leaf(means3D_in), opacity = leaf(opacity_in), dc = leaf(dc_in), sh = leaf(sh_in),
                scales = leaf(scales_in), rotations = leaf(rotations_in);
  // "screenspace points": zero tensor whose gradient receives dL/dmean2D (src/gaussian_renderer.cpp:44-51)
  torch::Tensor screenspace_points = torch::zeros_like(means3D).requires_grad_(true);
  torch::Tensor none;  // undefined == Python None: colors_precomp / cov3D_precomp absent
  auto out = rasterizer.forward(means3D, screenspace_points, opacity, dc, sh, none, scales, rotations, none);
  torch::Tensor color = std::get<0>(out), radii = std::get<1>(out);
  auto loss = (color * dL_dcolor).sum();
  {
    pybind11::gil_scoped_release no_gil;
    loss.backward();
  }
  return {color.detach(), radii,        means3D.grad(), screenspace_points.grad(), dc.grad(), sh.grad(),
          opacity.grad(), scales.grad(), rotations.grad()};
}

static torch::Tensor mark_visible(torch::Tensor positions, torch::Tensor bg, torch::Tensor viewmatrix,
                                  torch::Tensor projmatrix, torch::Tensor campos, int64_t H, int64_t W, double tanfovx,
                                  double tanfovy) {
  GaussianRasterizationSettings raster_settings((int)H, (int)W, (float)tanfovx, (float)tanfovy, bg, 1.0f, viewmatrix,
                                                projmatrix, 3, campos, false, false);
  GaussianRasterizer rasterizer(raster_settings);
  return rasterizer.markVisibleGaussians(positions);
}

// the error behaviour of GaussianRasterizer::forward (src/gaussian_rasterizer.cpp:272-287)
static void forward_without_colours(torch::Tensor means3D, torch::Tensor opacity, torch::Tensor scales,
                                    torch::Tensor rotations, torch::Tensor bg, torch::Tensor viewmatrix,
                                    torch::Tensor projmatrix, torch::Tensor campos) {
  GaussianRasterizationSettings raster_settings(16, 16, 1.0f, 1.0f, bg, 1.0f, viewmatrix, projmatrix, 3, campos, false,
                                                false);
  GaussianRasterizer rasterizer(raster_settings);
  torch::Tensor none, m2 = torch::zeros_like(means3D);
  rasterizer.forward(means3D, m2, opacity, none, none, none, scales, rotations, none);
}

PYBIND11_MODULE(gsdf_dropin, m) {
  m.def("render_forward_backward", &render_forward_backward);
  m.def("mark_visible", &mark_visible);
  m.def("forward_without_colours", &forward_without_colours);
}
