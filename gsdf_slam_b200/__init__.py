"""gsdf_slam_b200 — B200-native (sm_100a) implementation of GSDF-SLAM's mapping hot path:
the tile-based differentiable 3D-Gaussian rasterizer behind the reference's
GaussianRasterizer / RasterizeGaussiansCUDA API, its simple-knn initialiser and a fused Adam.

Layout: csrc/ (CUDA kernels + C ABI, built into libgsr.so), rasterizer.py / optim.py
(host-side mirror of the reference interface over the C ABI), scenes.py (seeded inputs),
mapping.py (keyframe-batch data-parallel step), loss.py (fused L1 + SSIM loss).  No CPU fallback exists in this package.
"""
from ._lib import build, lib, LIB_PATH  # noqa: F401
from .rasterizer import (  # noqa: F401
    GaussianRasterizationSettings, GaussianRasterizer, GaussianRasterizerFunction,
    RasterizeGaussiansBackwardCUDA, RasterizeGaussiansCUDA, distCUDA2, markVisible,
    rasterizeGaussians)
from .optim import FusedAdam  # noqa: F401
from .loss import l1_ssim_loss  # noqa: F401

__version__ = "0.1.0"
