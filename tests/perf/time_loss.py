"""Developer tool: fused L1+SSIM loss vs the reference's LibTorch op sequence (fwd+bwd), 1200x680."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gsdf_slam_b200.loss import l1_ssim_loss
from oracle.loss_ref import reference_loss
H, W = 680, 1200
gt = torch.rand((3, H, W), device="cuda"); img = (gt + 0.1 * torch.randn_like(gt)).clamp(0, 1)
mask = torch.ones((H, W), device="cuda"); mask[:8] = 0
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return np.median(ts) * 1000
def ref():
    x = img.clone().requires_grad_(True); l = reference_loss(x, gt, mask, 0.2); l.backward(); return x.grad
print(f"fused l1+ssim fwd+bwd   {t(lambda: l1_ssim_loss(img, gt, mask, 0.2)):8.1f} us")
print(f"LibTorch op sequence    {t(ref):8.1f} us")
