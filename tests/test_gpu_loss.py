"""GPU parity of the fused L1 + SSIM loss (gsr_l1_ssim_loss) against the REFERENCE'S OWN
include/loss_utils.h compiled unmodified (oracle/_ref/gsdf_ref_loss.so; statement sequence of
src/gaussian_mapper.cpp:415-431: torch conv2d with the 11x11 window, where-masking, autograd).  When that module did
not travel, the LibTorch restatement oracle/loss_ref.py::reference_loss (pinned against the compiled header by
tests/test_loss_oracle.py) takes its place.  Tolerances: loss value 1e-5 relative, gradient 1e-4 relative L2 (fp32
summation order of the separable convolution differs from cuDNN's 2-D window)."""
import numpy as np
import pytest
import torch

from util import _lib, rel_l2

pytestmark = pytest.mark.gpu


def _images(H, W, seed, mask_kind):
    g = torch.Generator(device="cuda").manual_seed(seed)
    gt = torch.rand((3, H, W), device="cuda", generator=g)
    # a smooth-ish rendering close to gt plus noise, values slightly outside [0, 1] too
    img = (gt + 0.15 * torch.randn((3, H, W), device="cuda", generator=g)).clamp(-0.1, 1.2)
    img[:, : H // 4] = gt[:, : H // 4]          # exact matches: sign(0) = 0 in the L1 gradient
    mask = None
    if mask_kind == "border":
        mask = torch.ones((H, W), device="cuda")
        mask[:7] = 0
        mask[:, -9:] = 0
        mask[H // 2, W // 3] = 0
    elif mask_kind == "random":
        mask = (torch.rand((H, W), device="cuda", generator=g) > 0.3).float()
    return img, gt, mask


@pytest.mark.parametrize("H,W,mask_kind,lam", [(68, 120, None, 0.2), (97, 131, "border", 0.2), (33, 47, "random", 0.5),
                                               (480, 640, "border", 0.2), (11, 9, None, 0.2), (680, 1200, None, 0.2)])
def test_fused_l1_ssim_matches_the_compiled_reference_header(H, W, mask_kind, lam):
    from gsdf_slam_b200.loss import l1_ssim_loss
    from oracle.loss_ref import reference_mapping_loss
    img, gt, mask = _images(H, W, 100 + H, mask_kind)
    ref, ref_l1, ref_ssim, ref_grad = reference_mapping_loss(img, gt, mask, lam)   # the compiled reference header
    stats, grad = l1_ssim_loss(img, gt, mask, lam)
    torch.cuda.synchronize()
    rv = float(ref)
    assert abs(float(stats[0]) - rv) <= 1e-5 * max(1.0, abs(rv)), (float(stats[0]), rv)
    if ref_l1 is not None:
        assert abs(float(stats[1]) - float(ref_l1)) <= 1e-5 and abs(float(stats[2]) - float(ref_ssim)) <= 1e-5
    assert torch.isfinite(grad).all()
    err = rel_l2(grad, ref_grad)
    assert err <= 1e-4, f"dL/dimage relative L2 {err:.3e}"
    if mask is not None:
        off = (mask <= 0).unsqueeze(0).expand_as(grad)
        assert torch.equal(grad[off], torch.zeros_like(grad[off]))     # masked pixels get no gradient
        assert torch.equal(ref_grad[off], torch.zeros_like(grad[off]))


def test_fused_loss_drives_the_rasterizer_backward():
    """dL/d(image) of the fused loss is what RasterizeGaussiansBackwardCUDA consumes."""
    from util import run_product_backward, run_product_forward, scenes, to_dev
    from gsdf_slam_b200.loss import l1_ssim_loss
    scene = scenes.tiny(seed=5, P=3000, W=200, H=136, degree=3)
    t = to_dev(scene)
    fwd = run_product_forward(scene, t)
    gt = torch.rand_like(fwd[2])
    stats, grad = l1_ssim_loss(fwd[2], gt, None, 0.2)
    g = run_product_backward(scene, t, fwd, dL=grad)
    torch.cuda.synchronize()
    assert 0.0 < float(stats[0]) < 2.0 and -1.0 <= float(stats[2]) <= 1.0
    assert all(torch.isfinite(v).all() for v in g.values())
    assert float(g["dL_dmeans3D"].abs().sum()) > 0.0


def test_loss_rejects_bad_arguments():
    import ctypes as C
    L = _lib.lib()
    x = torch.zeros(3 * 8 * 8, device="cuda")
    p = C.c_void_p(x.data_ptr())
    assert L.gsr_l1_ssim_loss(p, p, None, 0, 8, 0.2, None, p, p, None) != 0
    assert L.gsr_l1_ssim_loss(None, p, None, 8, 8, 0.2, None, p, p, None) != 0
    assert L.gsr_l1_ssim_loss(p, p, None, 8, 8, 0.2, None, p, None, None) != 0
    assert L.gsr_l1_ssim_workspace_bytes(8, 8) >= 9 * 8 * 8 * 4
