"""cfg3-style end-to-end mapping loop through every replaced row of SURVEY.md §8 at once: a map created
from a coloured point cloud (createFromPcd + distCUDA2), then per iteration raw parameters -> rasterizer
(activations fused, row f-2) -> fused L1+SSIM loss (f-1) -> backward -> densification statistics -> Adam on
the store's moments, with densifyAndPrune / increasePcd / resetOpacity changing the Gaussian count between
iterations (f-3) — the order of GaussianMapper::trainForOneIteration (src/gaussian_mapper.cpp:341-466).
Checks what the reference's loop relies on: the loss falls, the row count really changes, every array keeps
the same length, nothing becomes non-finite, and the rasterizer's optimistic per-frame buffers follow a P
that changes every few iterations."""
import numpy as np
import pytest
import torch

import gsdf_slam_b200 as gs
from gsdf_slam_b200.gaussian_model import GaussianModel
from gsdf_slam_b200.loss import l1_ssim_loss
from util import scenes, to_dev

pytestmark = pytest.mark.gpu


def _render(m, t, cam, degree):
    e = t["empty"]
    return gs.RasterizeGaussiansCUDA(t["bg"], m.xyz_, e, m.opacity_, m.scaling_, m.rotation_, 1.0, e, t["viewmatrix"],
                                     t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W,
                                     m.features_dc_.view(-1, 3), m.features_rest_, degree, t["campos"], False, False,
                                     raw_params=True)


def test_mapping_loop_with_densification():
    scene = scenes.tiny(seed=51, P=6000, W=208, H=144, degree=3)
    t = to_dev(scene)
    cam, e = scene.cam, t["empty"]
    target = gs.RasterizeGaussiansCUDA(t["bg"], t["means3D"], e, t["opacity"], t["scales"], t["rotations"], 1.0, e,
                                       t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W,
                                       t["dc"], t["sh"], 3, t["campos"], False, False)[2].clone()
    # the map starts from a sub-sampled, recoloured version of the scene's points
    gen = torch.Generator(device="cuda").manual_seed(5)
    keep = torch.rand(scene.P, device="cuda", generator=gen) < 0.5
    pts = t["means3D"][keep].contiguous()
    cols = (t["dc"][keep].view(-1, 3) * 0.28209479177387814 + 0.5).clamp(0, 1)
    m = GaussianModel(sh_degree=3, device="cuda")
    m.createFromPcd(pts, cols, 1.0)
    m.lr_ = dict(xyz=1.6e-4, f_dc=0.0025 * 4, f_rest=0.000125 * 4, opacity=0.05, scaling=0.005, rotation=0.001)
    sizes, losses = [m.P], []
    for it in range(1, 61):
        f = _render(m, t, cam, 3)
        R, B, image, radii = f[0], f[1], f[2], f[3]
        stats, dL = l1_ssim_loss(image, target, None, 0.2)
        g = gs.RasterizeGaussiansBackwardCUDA(t["bg"], m.xyz_, radii, e, m.scaling_, m.rotation_, 1.0, e,
                                              t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy, dL,
                                              m.features_dc_.view(-1, 3), m.features_rest_, 3, t["campos"], f[4], R,
                                              f[5], f[6], B, f[7], False, raw_opacity=m.opacity_)
        losses.append(float(stats[0]))
        m.addDensificationStats(g[0], radii)
        # (xyz, f_dc, f_rest, opacity, scaling, rotation) <- (dL_dmeans3D, dL_ddc, dL_dsh, dL_dopacity, dL_dscales, dL_drot)
        if it % 10 == 0:
            c = m.densifyAndPrune(0.00005, 0.005, 4.0, 20 if it > 30 else 0)
            assert c["rows"] == m.P
            sizes.append(m.P)
        else:
            m.step([g[3], g[5], g[6], g[2], g[7], g[8]], row_mask=radii)
        if it == 25:
            extra = t["means3D"][~keep][:1500].contiguous()
            ecol = (t["dc"][~keep][:1500].view(-1, 3) * 0.28209479177387814 + 0.5).clamp(0, 1)
            added = m.increasePcd(extra, ecol, it)
            assert 0 < added <= 1500
            sizes.append(m.P)
        if it == 45:
            m.resetOpacity()
        assert m.xyz_.shape[0] == m.opacity_.shape[0] == m.exist_since_iter_.shape[0] == m.P
    torch.cuda.synchronize()
    assert all(np.isfinite(losses))
    for p in m.params():
        assert bool(torch.isfinite(p).all())
    assert len(set(sizes)) >= 4, sizes                     # the count changed at every surgery
    assert min(losses[-5:]) < 0.8 * losses[0], (losses[0], losses[-5:])
