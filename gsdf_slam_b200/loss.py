"""Fused L1 + SSIM photometric loss behind the reference's names.

The reference computes, per mapping iteration (/root/reference/src/gaussian_mapper.cpp:414-431,
include/loss_utils.h:28-124)::

    masked_image    = where(mask > 0, rendered_image, 1)
    masked_gt_image = where(mask > 0, gt_image, 1)
    loss = (1 - lambda_dssim) * l1_loss(masked_image, masked_gt_image)
         + lambda_dssim * (1 - ssim(masked_image, gt_image))
    loss.backward()

with ~25 LibTorch kernels and an autograd graph.  ``l1_ssim_loss`` does the forward AND the
backward with two kernels of ``libgsr.so`` (``gsr_l1_ssim_loss``) and returns dL/d(rendered_image)
ready for ``RasterizeGaussiansBackwardCUDA``.  ``l1_loss`` / ``ssim`` below are the plain LibTorch
op sequences of the reference (same names), used by the tests as the fp32 reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib


def l1_ssim_loss(image: torch.Tensor, gt_image: torch.Tensor, mask: Optional[torch.Tensor] = None,
                 lambda_dssim: float = 0.2) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (stats, dL_dimage): stats = device tensor [3] {loss, L1, SSIM}; dL_dimage [3,H,W]."""
    if image.dim() != 3 or image.size(0) != 3 or gt_image.shape != image.shape:
        raise RuntimeError("image and gt_image must be [3,H,W]")
    if image.device.type != "cuda":
        raise RuntimeError("l1_ssim_loss needs CUDA tensors (no CPU fallback)")
    H, W = int(image.size(1)), int(image.size(2))
    img = image.contiguous().float()
    gt = gt_image.contiguous().float()
    m = None
    if mask is not None:
        m = mask.reshape(H, W).contiguous().float()
    L = _lib.lib()
    dev = image.device
    ws = torch.empty(int(L.gsr_l1_ssim_workspace_bytes(H, W)), dtype=torch.uint8, device=dev)
    stats = torch.empty(3, dtype=torch.float32, device=dev)
    grad = torch.empty_like(img)
    with torch.cuda.device(dev):
        _lib.check(L.gsr_l1_ssim_loss(C.c_void_p(img.data_ptr()), C.c_void_p(gt.data_ptr()),
                                      C.c_void_p(m.data_ptr()) if m is not None else None, H, W,
                                      float(lambda_dssim), C.c_void_p(stats.data_ptr()),
                                      C.c_void_p(grad.data_ptr()), C.c_void_p(ws.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "gsr_l1_ssim_loss")
    return stats, grad


# ---- the reference's op sequences (include/loss_utils.h), for tests --------------------------
def l1_loss(network_output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """loss_utils.h:28-31"""
    return torch.abs(network_output - gt).mean()


def _gaussian(window_size: int, sigma: float, device) -> torch.Tensor:
    """loss_utils.h:48-63 (fp32 like the reference's std::exp on float)"""
    import numpy as np
    vals = [float(np.exp(np.float32(-(x - window_size // 2) ** 2) / np.float32(2.0 * sigma * sigma)))
            for x in range(window_size)]
    g = torch.tensor(vals, dtype=torch.float32, device=device)
    return g / g.sum()


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """loss_utils.h:65-124 (size_average=True)"""
    F = torch.nn.functional
    channel = img1.size(-3)
    w1 = _gaussian(window_size, 1.5, img1.device).unsqueeze(1)
    window = (w1 @ w1.t()).float().unsqueeze(0).unsqueeze(0).expand(channel, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    x, y = img1.unsqueeze(0), img2.unsqueeze(0)
    mu1 = F.conv2d(x, window, padding=pad, groups=channel)
    mu2 = F.conv2d(y, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(x * x, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(y * y, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(x * y, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 * 0.01, 0.03 * 0.03
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def reference_loss(rendered: torch.Tensor, gt_image: torch.Tensor, mask: Optional[torch.Tensor],
                   lambda_dssim: float) -> torch.Tensor:
    """src/gaussian_mapper.cpp:414-429, op by op (differentiable; the tests' fp32 reference)."""
    if mask is not None:
        white = torch.ones_like(rendered)
        m = mask.reshape(1, *rendered.shape[1:]) > 0
        masked_image = torch.where(m, rendered, white)
        masked_gt = torch.where(m, gt_image, white)
    else:
        masked_image, masked_gt = rendered, gt_image
    return (1.0 - lambda_dssim) * l1_loss(masked_image, masked_gt) + lambda_dssim * (1.0 - ssim(masked_image, gt_image))
