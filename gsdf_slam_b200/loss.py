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
ready for ``RasterizeGaussiansBackwardCUDA``.  (Its oracle — the reference's own ``loss_utils.h`` compiled
unmodified, and a LibTorch restatement — lives under ``oracle/`` and is only used by tests and the reference arm.)
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
