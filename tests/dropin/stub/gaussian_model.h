// Stub standing in for /root/reference/include/gaussian_model.h when the reference's autograd glue
// (src/gaussian_rasterizer.cpp) is compiled against the drop-in shim (tests/dropin/build_dropin.py).
// include/gaussian_rasterizer.h:23 includes "gaussian_model.h" but uses nothing from it; the real header pulls
// Eigen / Sophus / OpenCV, which this image does not have.
#pragma once
#include <torch/torch.h>
class GaussianModel;
