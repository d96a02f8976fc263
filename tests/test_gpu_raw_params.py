"""Row f-2: the raw-parameter entry points (activations fused into the preprocess kernels)
against the reference's formulation — LibTorch activations (torch.sigmoid / torch.exp /
F.normalize, src/gaussian_model.cpp:48-71) feeding the activated-input rasterizer and autograd
carrying the gradients back to the raw parameters.  Activated values may differ by 1 ulp (ATen's
norm reduction order), so the image is compared to 1e-4 max-abs and gradients to 1e-4 rel-L2."""
import numpy as np
import pytest
import torch

import gsdf_slam_b200 as gs
from util import rel_l2, scenes, to_dev

pytestmark = pytest.mark.gpu


def _raw(scene, t):
    op = t["opacity"].clamp(1e-4, 1 - 1e-4)
    raw_op = torch.log(op / (1 - op))
    raw_sc = torch.log(t["scales"])
    gen = torch.Generator(device="cuda").manual_seed(3)
    raw_rot = t["rotations"] * (0.5 + torch.rand((scene.P, 1), device="cuda", generator=gen))  # not unit length
    return raw_op, raw_sc, raw_rot


@pytest.mark.parametrize("seed,degree", [(41, 3), (42, 1)])
def test_raw_parameter_path_matches_libtorch_activations_plus_autograd(seed, degree):
    scene = scenes.tiny(seed=seed, P=4000, W=208, H=144, degree=degree)
    t = to_dev(scene)
    cam = scene.cam
    e = t["empty"]
    raw_op, raw_sc, raw_rot = _raw(scene, t)

    # reference formulation: activations + autograd through them
    ro, rs, rr = (x.clone().requires_grad_(True) for x in (raw_op, raw_sc, raw_rot))
    op_a, sc_a, rot_a = torch.sigmoid(ro), torch.exp(rs), torch.nn.functional.normalize(rr)
    f_ref = gs.RasterizeGaussiansCUDA(t["bg"], t["means3D"], e, op_a.detach(), sc_a.detach(), rot_a.detach(), 1.0, e,
                                      t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W,
                                      t["dc"], t["sh"], degree, t["campos"], False, False)
    g_ref = gs.RasterizeGaussiansBackwardCUDA(t["bg"], t["means3D"], f_ref[3], e, sc_a.detach(), rot_a.detach(), 1.0, e,
                                              t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy,
                                              t["dL_dout"], t["dc"], t["sh"], degree, t["campos"], f_ref[4], f_ref[0],
                                              f_ref[5], f_ref[6], f_ref[1], f_ref[7], False)
    dL_dop, dL_dsc, dL_drot = g_ref[2], g_ref[7], g_ref[8]
    torch.autograd.backward([op_a, sc_a, rot_a], [dL_dop.view_as(op_a), dL_dsc, dL_drot])

    # fused: raw parameters in, raw-parameter gradients out
    f = gs.RasterizeGaussiansCUDA(t["bg"], t["means3D"], e, raw_op, raw_sc, raw_rot, 1.0, e, t["viewmatrix"],
                                  t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W, t["dc"], t["sh"], degree,
                                  t["campos"], False, False, raw_params=True)
    g = gs.RasterizeGaussiansBackwardCUDA(t["bg"], t["means3D"], f[3], e, raw_sc, raw_rot, 1.0, e, t["viewmatrix"],
                                          t["projmatrix"], cam.tanfovx, cam.tanfovy, t["dL_dout"], t["dc"], t["sh"],
                                          degree, t["campos"], f[4], f[0], f[5], f[6], f[1], f[7], False,
                                          raw_opacity=raw_op)
    torch.cuda.synchronize()
    assert (f[2] - f_ref[2]).abs().max().item() <= 1e-4
    assert (f[3] != f_ref[3]).float().mean().item() <= 1e-3          # radii: at most rare ceil() flips
    assert rel_l2(g[2].view_as(ro.grad), ro.grad) <= 1e-4            # d/d raw opacity
    assert rel_l2(g[7], rs.grad) <= 1e-4                             # d/d raw scaling
    assert rel_l2(g[8], rr.grad) <= 1e-4                             # d/d raw rotation
    for k in (0, 1, 3, 5, 6):                                        # the other gradients are unaffected
        assert rel_l2(g[k], g_ref[k]) <= 1e-4


def test_training_on_raw_parameters_reduces_the_photometric_loss():
    """The whole fused mapping iteration: raw parameters -> (activations + rasterizer) -> fused
    L1+SSIM loss -> backward to the raw parameters -> masked Adam.  The loss must go down."""
    from gsdf_slam_b200.mapping import Keyframe, KeyframeBatchMapper
    scene = scenes.tiny(seed=43, P=5000, W=208, H=144, degree=3)
    t = to_dev(scene)
    cam = scene.cam
    e = t["empty"]
    raw_op, raw_sc, raw_rot = _raw(scene, t)
    target = gs.RasterizeGaussiansCUDA(t["bg"], t["means3D"], e, t["opacity"], t["scales"], t["rotations"], 1.0, e,
                                       t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W,
                                       t["dc"], t["sh"], 3, t["campos"], False, False)[2].clone()
    gen = torch.Generator(device="cuda").manual_seed(7)
    params = dict(means3D=t["means3D"].clone(), dc=t["dc"] + 0.3 * torch.randn(t["dc"].shape, device="cuda", generator=gen),
                  sh=t["sh"].clone(), opacity=raw_op.clone(), scales=raw_sc + 0.2, rotations=raw_rot.clone())
    m = KeyframeBatchMapper(params, 3, t["bg"], raw_params=True,
                            lr=dict(dc=0.02, opacity=0.05, scales=0.01))
    kf = Keyframe(t["viewmatrix"], t["projmatrix"], t["campos"], cam.tanfovx, cam.tanfovy, cam.H, cam.W,
                  gt_image=target)
    losses = []
    for _ in range(40):
        m.step([kf])
        losses.append(float(m.last_loss[0]))
    assert all(np.isfinite(losses))
    assert losses[-1] < 0.6 * losses[0], (losses[0], losses[-1])
