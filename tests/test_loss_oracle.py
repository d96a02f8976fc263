"""CPU test: the LibTorch restatement of the reference's photometric loss (oracle/loss_ref.py) is pinned against the
reference's own include/loss_utils.h compiled unmodified (oracle/_ref/gsdf_ref_loss.so, oracle/build_ref_loss.py), on
CPU tensors: loss terms and the autograd gradient w.r.t. the rendered image, with and without a mask."""
import pytest
import torch

from oracle import loss_ref


@pytest.mark.parametrize("H,W,masked,lam", [(40, 56, False, 0.2), (33, 47, True, 0.2), (11, 9, True, 0.5)])
def test_restatement_matches_the_compiled_reference_header(H, W, masked, lam):
    m = loss_ref.ref_module()
    if m is None:
        pytest.skip("oracle/_ref/gsdf_ref_loss.so not built")
    g = torch.Generator().manual_seed(H * 100 + W)
    gt = torch.rand((3, H, W), generator=g)
    img = (gt + 0.2 * torch.randn((3, H, W), generator=g)).clamp(-0.1, 1.2)
    mask = (torch.rand((H, W), generator=g) > 0.3).float() if masked else None
    loss, l1, ssim, grad = loss_ref.reference_mapping_loss(img, gt, mask, lam)
    x = img.clone().requires_grad_(True)
    mine = loss_ref.reference_loss(x, gt, mask, lam)
    mine.backward()
    assert abs(float(mine) - float(loss)) <= 1e-6 * max(1.0, abs(float(loss)))
    assert float((x.grad - grad).abs().max()) <= 1e-7 + 1e-5 * float(grad.abs().max())
    # the individual terms, through the header's own functions
    mi = torch.where(mask.unsqueeze(0) > 0, img, torch.ones_like(img)) if masked else img
    mg = torch.where(mask.unsqueeze(0) > 0, gt, torch.ones_like(gt)) if masked else gt
    assert abs(float(loss_ref.l1_loss(mi, mg)) - float(l1)) <= 1e-6
    assert abs(float(loss_ref.ssim(mi, gt)) - float(ssim)) <= 1e-6
    assert abs(float(m.ssim(mi, gt)) - float(ssim)) <= 1e-7
    if masked:   # masked pixels receive exactly zero gradient in the reference
        off = (mask <= 0).unsqueeze(0).expand_as(grad)
        assert float(grad[off].abs().max()) == 0.0
