// TEST INFRASTRUCTURE ONLY.  pybind11 wrapper around the UNMODIFIED reference colorize_launcher
// (/root/reference/cuda_colorize/colorize_pointcloud.cu:57-107), compiled by oracle/build_ref_colorize.py
// into oracle/_ref/gsdf_ref_colorize.so.  Used by tests/test_gpu_ingest.py as the GPU oracle of row f-4.
#include <torch/extension.h>

#include <vector>

void colorize_launcher(torch::Tensor& points, torch::Tensor& colors, torch::Tensor& proj, torch::Tensor& Tr,
                       torch::Tensor& pose, torch::Tensor& image);

static std::vector<torch::Tensor> ref_colorize(torch::Tensor points, torch::Tensor colors, torch::Tensor proj,
                                               torch::Tensor Tr, torch::Tensor pose, torch::Tensor image) {
  colorize_launcher(points, colors, proj, Tr, pose, image);
  return {points, colors};
}

PYBIND11_MODULE(gsdf_ref_colorize, m) { m.def("colorize", &ref_colorize); }
