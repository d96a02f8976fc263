// TEST INFRASTRUCTURE ONLY.  pybind11 wrapper around the UNMODIFIED reference loss header
// /root/reference/include/loss_utils.h (l1_loss :28-31, ssim :110-124, _ssim :77-108), included where it lies and
// compiled by oracle/build_ref_loss.py into oracle/_ref/gsdf_ref_loss.so.  `mapping_loss` drives those functions with
// the statement sequence of GaussianMapper::trainForOneIteration (/root/reference/src/gaussian_mapper.cpp:415-431:
// white-fill masking with torch::where, L1 on the two masked images, SSIM against the UNMASKED ground truth,
// (1 - lambda) * L1 + lambda * (1 - SSIM), loss.backward()) and returns the loss terms and dL/d(rendered image).
// It is the GPU oracle of row f-1 (tests/test_gpu_loss.py) and the loss of the reference arm of bench.py.
#include <torch/extension.h>

#include <vector>

#include "loss_utils.h"

static torch::Tensor ref_l1_loss(torch::Tensor a, torch::Tensor b) { return loss_utils::l1_loss(a, b); }

static torch::Tensor ref_ssim(torch::Tensor a, torch::Tensor b) {
  return loss_utils::ssim(a, b, a.device().type());
}

// -> {loss, L1, SSIM, dL_drendered}
static std::vector<torch::Tensor> mapping_loss(torch::Tensor rendered, torch::Tensor gt_image, torch::Tensor mask,
                                               double lambda_dssim_d, bool use_mask) {
  torch::Tensor rendered_image = rendered.detach().clone().requires_grad_(true);
  torch::Tensor masked_image = rendered_image, masked_gt_image = gt_image;
  if (use_mask) {
    torch::Tensor white = torch::ones_like(rendered_image);
    masked_image = torch::where(mask > 0, rendered_image, white);
    masked_gt_image = torch::where(mask > 0, gt_image, white);
  }
  auto Ll1 = loss_utils::l1_loss(masked_image, masked_gt_image);
  float lambda_dssim = (float)lambda_dssim_d;
  auto ssim_v = loss_utils::ssim(masked_image, gt_image, rendered.device().type());
  auto loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - ssim_v);
  {
    pybind11::gil_scoped_release no_gil;   // the autograd engine must not be entered with the GIL held
    loss.backward();
  }
  return {loss.detach(), Ll1.detach(), ssim_v.detach(), rendered_image.grad()};
}

PYBIND11_MODULE(gsdf_ref_loss, m) {
  m.def("l1_loss", &ref_l1_loss);
  m.def("ssim", &ref_ssim);
  m.def("mapping_loss", &mapping_loss);
}
