// TEST INFRASTRUCTURE ONLY.  The declarations the reference's densification code needs in order to be compiled
// OUT of its translation unit: oracle/build_ref_densify.py extracts the BODIES of GaussianModel::{getXYZ,
// getScalingActivation, getOpacityActivation, percentDense, resetOpacity, replaceTensorToOptimizer, prunePoints,
// densificationPostfix, densifyAndSplit, densifyAndClone, densifyAndPrune, addDensificationStats} verbatim from
// /root/reference/src/gaussian_model.cpp (:48-71, 399-681, 997-1001) into a generated file under oracle/_ref
// (git-ignored) that includes this header first.  The real include/gaussian_model.h pulls Eigen / Sophus / OpenCV
// (absent in this image) and the code was written against LibTorch 1.x, whose optimizer state is keyed by
// c10::guts::to_string(TensorImpl*) strings; what follows stands in for exactly those two things and nothing else:
// the member list of GaussianModel the bodies touch, a minimal optimizer with the 1.x state() / param_groups()
// shape, c10::guts::to_string, and a no-op emptyCache().  general_utils.h (inverse_sigmoid, build_rotation) is the
// reference's own header, included where it lies.
#pragma once
#include <torch/torch.h>

#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "general_utils.h"   // -I /root/reference/include

namespace c10 {
namespace guts {
template <class T>
inline std::string to_string(T* p) {   // LibTorch 1.x had this; 2.x dropped it
  std::ostringstream os;
  os << static_cast<const void*>(p);
  return os.str();
}
}  // namespace guts
namespace cuda {
namespace CUDACachingAllocator {
inline void emptyCache() {}   // allocator trimming is not part of the arithmetic
}  // namespace CUDACachingAllocator
}  // namespace cuda
}  // namespace c10

namespace ref_stub {
struct ParamGroup {
  std::vector<torch::Tensor> p;
  std::vector<torch::Tensor>& params() { return p; }
};
// the part of torch::optim::Adam (LibTorch 1.x) the bodies use
struct Adam {
  std::vector<ParamGroup> groups;
  std::unordered_map<std::string, std::unique_ptr<torch::optim::OptimizerParamState>> st;
  std::vector<ParamGroup>& param_groups() { return groups; }
  std::unordered_map<std::string, std::unique_ptr<torch::optim::OptimizerParamState>>& state() { return st; }
};
}  // namespace ref_stub

@TENSORS_TO_VEC_MACRO@

class GaussianModel {
 public:
  torch::Tensor getScalingActivation();
  torch::Tensor getXYZ();
  torch::Tensor getOpacityActivation();
  float percentDense();
  void resetOpacity();
  torch::Tensor replaceTensorToOptimizer(torch::Tensor& tensor, int tensor_idx);
  void prunePoints(torch::Tensor& mask);
  void densificationPostfix(torch::Tensor& new_xyz, torch::Tensor& new_features_dc, torch::Tensor& new_features_rest,
                            torch::Tensor& new_opacities, torch::Tensor& new_scaling, torch::Tensor& new_rotation,
                            torch::Tensor& new_exist_since_iter);
  void densifyAndSplit(torch::Tensor& grads, float grad_threshold, float scene_extent, int N = 2);
  void densifyAndClone(torch::Tensor& grads, float grad_threshold, float scene_extent);
  void densifyAndPrune(float max_grad, float min_opacity, float extent, int max_screen_size);
  void addDensificationStats(torch::Tensor& viewspace_point_tensor, torch::Tensor& update_filter);

  torch::DeviceType device_type_ = torch::kCPU;
  float percent_dense_ = 0.01f;
  std::mutex mutex_settings_;   // percentDense() locks it
  torch::Tensor xyz_, features_dc_, features_rest_, scaling_, rotation_, opacity_;
  torch::Tensor max_radii2D_, xyz_gradient_accum_, denom_, exist_since_iter_;
  std::vector<torch::Tensor> Tensor_vec_xyz_, Tensor_vec_feature_dc_, Tensor_vec_feature_rest_, Tensor_vec_opacity_,
      Tensor_vec_scaling_, Tensor_vec_rotation_;
  std::shared_ptr<ref_stub::Adam> optimizer_;
};
