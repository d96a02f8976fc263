// LibTorch shim over the C ABI (include/gsr_capi.h): unpacks tensors, provides the resize
// callbacks (the reference's resizeFunctional, src/rasterize_points.cu:28-36) and translates
// error codes into the exceptions the reference throws.
#include "rasterize_points.h"

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <stdexcept>
#include <string>

#include "gsr_capi.h"

namespace {
char* resize_cb(void* user, size_t n) {
  try {
    auto* t = static_cast<torch::Tensor*>(user);
    t->resize_({(long long)n});
    return n ? reinterpret_cast<char*>(t->data_ptr()) : nullptr;
  } catch (...) {
    return nullptr;  // never let an exception cross the C ABI
  }
}
const float* fptr(const torch::Tensor& t) { return t.numel() ? t.data_ptr<float>() : nullptr; }
void check(int rc, const char* what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + " failed with code " + std::to_string(rc));
}
int coeffs(const torch::Tensor& sh) { return (sh.dim() >= 2 && sh.size(0) != 0) ? (int)sh.size(1) : 0; }
}  // namespace

std::tuple<int, int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                       const torch::Tensor& colors, const torch::Tensor& opacity,
                       const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                       const float tan_fovx, const float tan_fovy, const int image_height,
                       const int image_width, const torch::Tensor& dc, const torch::Tensor& sh,
                       const int degree, const torch::Tensor& campos, const bool prefiltered,
                       const bool debug) {
  if (means3D.ndimension() != 2 || means3D.size(1) != 3) {
    TORCH_CHECK(false, "means3D must have dimensions (num_points, 3)");
  }
  const int P = means3D.size(0), H = image_height, W = image_width;
  auto fopt = means3D.options().dtype(torch::kFloat32);
  torch::Tensor out_color = torch::empty({3, H, W}, fopt);  // fully written by the callee
  torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
  auto bopt = torch::TensorOptions(torch::kByte).device(means3D.device());
  torch::Tensor geom = torch::empty({0}, bopt), binning = torch::empty({0}, bopt),
                img = torch::empty({0}, bopt), sample = torch::empty({0}, bopt);
  auto bg_c = background.contiguous(), m_c = means3D.contiguous(), col_c = colors.contiguous(),
       op_c = opacity.contiguous(), sc_c = scales.contiguous(), rot_c = rotations.contiguous(),
       cov_c = cov3D_precomp.contiguous(), v_c = viewmatrix.contiguous(), p_c = projmatrix.contiguous(),
       dc_c = dc.contiguous(), sh_c = sh.contiguous(), cam_c = campos.contiguous();
  int rendered = 0, buckets = 0;
  gsr_forward_args a{};
  a.P = P; a.D = degree; a.M = coeffs(sh); a.W = W; a.H = H;
  a.background = fptr(bg_c); a.means3D = fptr(m_c); a.dc = fptr(dc_c); a.shs = fptr(sh_c);
  a.colors_precomp = fptr(col_c); a.opacities = fptr(op_c); a.scales = fptr(sc_c);
  a.scale_modifier = scale_modifier; a.rotations = fptr(rot_c); a.cov3D_precomp = fptr(cov_c);
  a.viewmatrix = fptr(v_c); a.projmatrix = fptr(p_c); a.campos = fptr(cam_c);
  a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy; a.prefiltered = prefiltered; a.debug = debug;
  a.out_color = out_color.data_ptr<float>(); a.radii = P ? radii.data_ptr<int>() : nullptr;
  a.geom_resize = resize_cb; a.geom_user = &geom;
  a.binning_resize = resize_cb; a.binning_user = &binning;
  a.image_resize = resize_cb; a.image_user = &img;
  a.sample_resize = resize_cb; a.sample_user = &sample;
  a.stream = c10::cuda::getCurrentCUDAStream(means3D.device().index()).stream();
  a.num_rendered = &rendered; a.num_buckets = &buckets;
  check(gsr_forward(&a), "gsr_forward");
  return std::make_tuple(rendered, buckets, out_color, radii, geom, binning, img, sample);
}

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
                               const int B, const torch::Tensor& sampleBuffer, const bool debug) {
  const int P = means3D.size(0);
  const int H = dL_dout_color.size(1), W = dL_dout_color.size(2);
  const int M = coeffs(sh);
  auto opt = means3D.options();
  // the callee writes every row (zeros for Gaussians that were not rendered): no memsets here
  auto mk = [&](std::initializer_list<int64_t> s) { return P ? torch::empty(s, opt) : torch::zeros(s, opt); };
  torch::Tensor dL_dmeans3D = mk({P, 3}), dL_dmeans2D = mk({P, 3}), dL_dcolors = mk({P, 3}),
                dL_dopacity = mk({P, 1}), dL_dcov3D = mk({P, 6}), dL_ddc = mk({P, 1, 3}),
                dL_dsh = mk({P, M, 3}), dL_dscales = mk({P, 3}), dL_drotations = mk({P, 4});
  if (P != 0) {
    auto bg_c = background.contiguous(), m_c = means3D.contiguous(), col_c = colors.contiguous(),
         sc_c = scales.contiguous(), rot_c = rotations.contiguous(), cov_c = cov3D_precomp.contiguous(),
         v_c = viewmatrix.contiguous(), p_c = projmatrix.contiguous(), dc_c = dc.contiguous(),
         sh_c = sh.contiguous(), cam_c = campos.contiguous(), dpix_c = dL_dout_color.contiguous(),
         radii_c = radii.contiguous();
    gsr_backward_args a{};
    a.P = P; a.D = degree; a.M = M; a.R = R; a.B = B; a.W = W; a.H = H;
    a.background = fptr(bg_c); a.means3D = fptr(m_c); a.dc = fptr(dc_c); a.shs = fptr(sh_c);
    a.colors_precomp = fptr(col_c); a.scales = fptr(sc_c); a.scale_modifier = scale_modifier;
    a.rotations = fptr(rot_c); a.cov3D_precomp = fptr(cov_c); a.viewmatrix = fptr(v_c);
    a.projmatrix = fptr(p_c); a.campos = fptr(cam_c); a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
    a.radii = radii_c.data_ptr<int>();
    a.geom_buffer = reinterpret_cast<char*>(geomBuffer.data_ptr());
    a.binning_buffer = binningBuffer.numel() ? reinterpret_cast<char*>(binningBuffer.data_ptr()) : nullptr;
    a.image_buffer = reinterpret_cast<char*>(imageBuffer.data_ptr());
    a.sample_buffer = sampleBuffer.numel() ? reinterpret_cast<char*>(sampleBuffer.data_ptr()) : nullptr;
    a.dL_dpix = fptr(dpix_c);
    a.dL_dmean2D = dL_dmeans2D.data_ptr<float>(); a.dL_dconic = nullptr;
    a.dL_dopacity = dL_dopacity.data_ptr<float>(); a.dL_dcolor = dL_dcolors.data_ptr<float>();
    a.dL_dmean3D = dL_dmeans3D.data_ptr<float>(); a.dL_dcov3D = dL_dcov3D.data_ptr<float>();
    a.dL_ddc = dL_ddc.data_ptr<float>(); a.dL_dsh = M ? dL_dsh.data_ptr<float>() : nullptr;
    a.dL_dscale = dL_dscales.data_ptr<float>(); a.dL_drot = dL_drotations.data_ptr<float>();
    a.debug = debug;
    a.stream = c10::cuda::getCurrentCUDAStream(means3D.device().index()).stream();
    check(gsr_backward(&a), "gsr_backward");
  }
  return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_ddc, dL_dsh,
                         dL_dscales, dL_drotations);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix) {
  const int P = means3D.size(0);
  torch::Tensor present = torch::full({P}, false, means3D.options().dtype(at::kBool));
  if (P != 0) {
    auto m = means3D.contiguous(), v = viewmatrix.contiguous(), p = projmatrix.contiguous();
    check(gsr_mark_visible(P, m.data_ptr<float>(), v.data_ptr<float>(), p.data_ptr<float>(),
                           reinterpret_cast<uint8_t*>(present.data_ptr<bool>()),
                           c10::cuda::getCurrentCUDAStream(means3D.device().index()).stream()),
          "gsr_mark_visible");
  }
  return present;
}

torch::Tensor distCUDA2(const torch::Tensor& points) {
  const int P = points.size(0);
  torch::Tensor means = torch::full({P}, 0.0, points.options().dtype(torch::kFloat32));
  if (P != 0) {
    auto pts = points.contiguous();
    check(gsr_knn_mean_dist2(P, pts.data_ptr<float>(), means.data_ptr<float>(),
                             c10::cuda::getCurrentCUDAStream(points.device().index()).stream()),
          "gsr_knn_mean_dist2");
  }
  return means;
}

void FusedAdamStep(std::vector<torch::Tensor>& params, const std::vector<torch::Tensor>& grads,
                   std::vector<torch::Tensor>& exp_avgs, std::vector<torch::Tensor>& exp_avg_sqs,
                   const std::vector<double>& lrs, const std::vector<int64_t>& steps, double beta1,
                   double beta2, double eps) {
  std::vector<gsr_adam_group> g(params.size());
  for (size_t i = 0; i < params.size(); ++i) {
    TORCH_CHECK(params[i].is_contiguous() && grads[i].is_contiguous(), "FusedAdamStep needs contiguous tensors");
    g[i].param = params[i].data_ptr<float>(); g[i].grad = grads[i].data_ptr<float>();
    g[i].exp_avg = exp_avgs[i].data_ptr<float>(); g[i].exp_avg_sq = exp_avg_sqs[i].data_ptr<float>();
    g[i].numel = params[i].numel(); g[i].lr = lrs[i]; g[i].step = steps[i];
  }
  if (!g.empty())
    check(gsr_adam_step(g.data(), (int)g.size(), beta1, beta2, eps, 1.0f,
                        c10::cuda::getCurrentCUDAStream(params[0].device().index()).stream()),
          "gsr_adam_step");
}

// ---- frame ingest: same symbol as the reference's colorize_launcher ----
// (cuda_colorize/colorize_pointcloud.h:4-11, implementation colorize_pointcloud.cu:57-107, called from
// submodules/data_loader/data_loader.cpp:189): `points` and `colors` are re-bound to the compacted
// survivors (world-frame positions, pixel colours); no device-wide synchronisation.
void colorize_launcher(torch::Tensor& points, torch::Tensor& colors, torch::Tensor& proj, torch::Tensor& Tr,
                       torch::Tensor& pose, torch::Tensor& image) {
  if (!points.is_cuda()) points = points.to(torch::kCUDA);
  const auto dev = points.device();
  auto pts = points.to(torch::kFloat32).contiguous();
  auto img = image.to(dev).to(torch::kFloat32).contiguous();
  auto proj_mat = proj.to(dev).to(torch::kFloat32).contiguous();
  auto Tr_velo_to_cam = Tr.to(dev).to(torch::kFloat32).contiguous();
  auto lidar_pose = pose.to(dev).to(torch::kFloat32).contiguous();
  TORCH_CHECK(pts.dim() == 2 && pts.size(1) == 3, "points must have dimensions (num_points, 3)");
  TORCH_CHECK(img.dim() == 3 && img.size(2) == 3, "image must have dimensions (H, W, 3)");
  TORCH_CHECK(proj_mat.numel() == 12 && Tr_velo_to_cam.numel() == 16 && lidar_pose.numel() == 16,
              "proj must have 12 elements, Tr and pose 16");
  const int64_t n = pts.size(0);
  auto out_p = torch::empty({n, 3}, pts.options());
  auto out_c = torch::empty({n, 3}, pts.options());
  int64_t count = 0;
  if (n > 0) {
    c10::cuda::CUDAGuard guard(dev);
    auto ws = torch::empty({(int64_t)gsr_colorize_workspace_bytes(n)}, pts.options().dtype(torch::kByte));
    check(gsr_colorize_points(pts.data_ptr<float>(), n, lidar_pose.data_ptr<float>(), proj_mat.data_ptr<float>(),
                              Tr_velo_to_cam.data_ptr<float>(), img.data_ptr<float>(), (int)img.size(1),
                              (int)img.size(0), out_p.data_ptr<float>(), out_c.data_ptr<float>(),
                              reinterpret_cast<char*>(ws.data_ptr()), &count,
                              c10::cuda::getCurrentCUDAStream(dev.index()).stream()),
          "gsr_colorize_points");
  }
  points = out_p.slice(0, 0, count).contiguous();
  colors = out_c.slice(0, 0, count).contiguous();
}

// Also reachable from Python (torch.ops.gsr_shim.*) so the shim itself is covered by tests.
TORCH_LIBRARY(gsr_shim, m) {
  m.def("rasterize_forward", [](const torch::Tensor& bg, const torch::Tensor& means3D, const torch::Tensor& colors,
                                const torch::Tensor& opacity, const torch::Tensor& scales,
                                const torch::Tensor& rotations, double scale_modifier,
                                const torch::Tensor& cov3D, const torch::Tensor& view, const torch::Tensor& proj,
                                double tanx, double tany, int64_t H, int64_t W, const torch::Tensor& dc,
                                const torch::Tensor& sh, int64_t degree, const torch::Tensor& campos,
                                bool prefiltered, bool debug) {
    auto r = RasterizeGaussiansCUDA(bg, means3D, colors, opacity, scales, rotations, (float)scale_modifier, cov3D,
                                    view, proj, (float)tanx, (float)tany, (int)H, (int)W, dc, sh, (int)degree,
                                    campos, prefiltered, debug);
    return std::make_tuple((int64_t)std::get<0>(r), (int64_t)std::get<1>(r), std::get<2>(r), std::get<3>(r),
                           std::get<4>(r), std::get<5>(r), std::get<6>(r), std::get<7>(r));
  });
  m.def("rasterize_backward", [](const torch::Tensor& bg, const torch::Tensor& means3D, const torch::Tensor& radii,
                                 const torch::Tensor& colors, const torch::Tensor& scales,
                                 const torch::Tensor& rotations, double scale_modifier, const torch::Tensor& cov3D,
                                 const torch::Tensor& view, const torch::Tensor& proj, double tanx, double tany,
                                 const torch::Tensor& dL, const torch::Tensor& dc, const torch::Tensor& sh,
                                 int64_t degree, const torch::Tensor& campos, const torch::Tensor& geom, int64_t R,
                                 const torch::Tensor& binning, const torch::Tensor& img, int64_t B,
                                 const torch::Tensor& sample, bool debug) {
    return RasterizeGaussiansBackwardCUDA(bg, means3D, radii, colors, scales, rotations, (float)scale_modifier,
                                          cov3D, view, proj, (float)tanx, (float)tany, dL, dc, sh, (int)degree,
                                          campos, geom, (int)R, binning, img, (int)B, sample, debug);
  });
  m.def("dist_cuda2", [](const torch::Tensor& pts) { return distCUDA2(pts); });
  m.def("colorize", [](const torch::Tensor& points, const torch::Tensor& proj, const torch::Tensor& Tr,
                       const torch::Tensor& pose, const torch::Tensor& image) {
    torch::Tensor p = points, c = torch::zeros_like(points), pr = proj, t = Tr, po = pose, im = image;
    colorize_launcher(p, c, pr, t, po, im);
    return std::make_tuple(p, c);
  });
}
