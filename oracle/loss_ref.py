"""TEST INFRASTRUCTURE ONLY — oracle of row f-1 (fused L1 + SSIM loss).

Two pieces:
  * ``ref_module()``: the reference's UNMODIFIED ``include/loss_utils.h`` compiled behind a pybind11 wrapper
    (oracle/ref_loss.cpp -> oracle/_ref/gsdf_ref_loss.so, built by oracle/build_ref_loss.py).  ``mapping_loss`` runs the
    statement sequence of GaussianMapper::trainForOneIteration (/root/reference/src/gaussian_mapper.cpp:415-431) on it,
    autograd included.  This is what pins parity.
  * ``l1_loss`` / ``ssim`` / ``reference_loss``: a LibTorch-op restatement of the same sequence
    (include/loss_utils.h:28-124), pinned against the compiled header by tests/test_loss_oracle.py; it is what remains
    usable if the compiled module did not travel.
Only tests/, smoke() and bench.py's reference arm may import this file; the product never does.
"""
from __future__ import annotations

import importlib.util
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "gsdf_ref_loss.so")
_mod = None


def ref_module():
    """The compiled reference header (None when oracle/_ref/gsdf_ref_loss.so is absent)."""
    global _mod
    if _mod is None and os.path.exists(_SO):
        spec = importlib.util.spec_from_file_location("gsdf_ref_loss", _SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def reference_mapping_loss(rendered: torch.Tensor, gt_image: torch.Tensor, mask: Optional[torch.Tensor],
                           lambda_dssim: float):
    """-> (loss, L1, SSIM, dL_drendered) from the compiled reference header; falls back to the restatement."""
    m = ref_module()
    if m is not None:
        mk = mask.reshape(1, *rendered.shape[1:]) if mask is not None else torch.empty(0, device=rendered.device)
        return tuple(m.mapping_loss(rendered, gt_image, mk, float(lambda_dssim), mask is not None))
    x = rendered.detach().clone().requires_grad_(True)
    loss = reference_loss(x, gt_image, mask, lambda_dssim)
    loss.backward()
    return loss.detach(), None, None, x.grad


# ---- restatement of the reference's op sequences (include/loss_utils.h) -------------------------
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
