"""Shared helpers of the test-suite: upload a seeded Scene, run the product (through the
C ABI) and the GPU oracle (reference CUDA build) on identical inputs."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import gsdf_slam_b200 as gs  # noqa: E402
from gsdf_slam_b200 import _lib  # noqa: E402
from gsdf_slam_b200 import scenes  # noqa: E402


def to_dev(scene, dev="cuda"):
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in scene.arrays().items()}
    t["empty"] = torch.empty(0, dtype=torch.float32, device=dev)
    return t


def run_product_forward(scene, t, colors_precomp=None, cov3D_precomp=None, sh=None, degree=None,
                        debug=False):
    cam = scene.cam
    e = t["empty"]
    sh = t["sh"] if sh is None else sh
    degree = scene.degree if degree is None else degree
    scales = e if cov3D_precomp is not None else t["scales"]
    rots = e if cov3D_precomp is not None else t["rotations"]
    return gs.RasterizeGaussiansCUDA(
        t["bg"], t["means3D"], e if colors_precomp is None else colors_precomp, t["opacity"],
        scales, rots, 1.0, e if cov3D_precomp is None else cov3D_precomp, t["viewmatrix"],
        t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W, t["dc"], sh, degree, t["campos"],
        False, debug)


def run_product_backward(scene, t, fwd, colors_precomp=None, cov3D_precomp=None, sh=None,
                         degree=None, dL=None):
    cam = scene.cam
    e = t["empty"]
    sh = t["sh"] if sh is None else sh
    degree = scene.degree if degree is None else degree
    scales = e if cov3D_precomp is not None else t["scales"]
    rots = e if cov3D_precomp is not None else t["rotations"]
    R, B, color, radii, geom, binning, img, sample = fwd
    out = gs.RasterizeGaussiansBackwardCUDA(
        t["bg"], t["means3D"], radii, e if colors_precomp is None else colors_precomp, scales, rots,
        1.0, e if cov3D_precomp is None else cov3D_precomp, t["viewmatrix"], t["projmatrix"],
        cam.tanfovx, cam.tanfovy, t["dL_dout"] if dL is None else dL, t["dc"], sh, degree,
        t["campos"], geom, R, binning, img, B, sample, False)
    names = ["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_ddc",
             "dL_dsh", "dL_dscales", "dL_drotations"]
    return dict(zip(names, out))


def product_exports(scene, t, fwd):
    """Decode the product's opaque buffers: sorted keys / ids, ranges, n_contrib, final_T ..."""
    import ctypes as C
    L = _lib.lib()
    R, B, color, radii, geom, binning, img, sample = fwd
    dev = color.device
    P, W, H = scene.P, scene.cam.W, scene.cam.H
    N = W * H
    T = ((W + 15) // 16) * ((H + 15) // 16)
    keys = torch.zeros(R, dtype=torch.int64, device=dev)
    ids = torch.zeros(R, dtype=torch.int32, device=dev)
    p = lambda x: C.c_void_p(x.data_ptr()) if x.numel() else None  # noqa: E731
    if R > 0:
        _lib.check(L.gsr_export_binning(p(geom), p(binning), p(img), P, R, W, H, p(keys), p(ids), None),
                   "export_binning")
    out = dict(keys=keys, ids=ids,
               final_T=torch.zeros(N, dtype=torch.float32, device=dev),
               n_contrib=torch.zeros(N, dtype=torch.int32, device=dev),
               ranges=torch.zeros((T, 2), dtype=torch.int32, device=dev),
               max_contrib=torch.zeros(T, dtype=torch.int32, device=dev),
               depths=torch.zeros(P, dtype=torch.float32, device=dev),
               means2D=torch.zeros((P, 2), dtype=torch.float32, device=dev),
               conic_opacity=torch.zeros((P, 4), dtype=torch.float32, device=dev),
               rgb=torch.zeros((P, 3), dtype=torch.float32, device=dev),
               tiles=torch.zeros(P, dtype=torch.int32, device=dev),
               clamped=torch.zeros(P, dtype=torch.uint8, device=dev))
    _lib.check(L.gsr_export_image(p(img), W, H, p(out["final_T"]), p(out["n_contrib"]),
                                  p(out["ranges"]), p(out["max_contrib"]), None), "export_image")
    _lib.check(L.gsr_export_geometry(p(geom), P, p(out["depths"]), p(out["means2D"]),
                                     p(out["conic_opacity"]), p(out["rgb"]), p(out["tiles"]),
                                     p(out["clamped"]), None), "export_geometry")
    torch.cuda.synchronize()
    return out


def run_reference(scene, t, colors_precomp=None, cov3D_precomp=None, sh=None, degree=None,
                  backward=True, dL=None):
    """Run the GPU oracle (reference CUDA) on the same tensors; returns dict of everything."""
    from oracle import ref_binding as rb
    cam = scene.cam
    e = t["empty"]
    sh = t["sh"] if sh is None else sh
    degree = scene.degree if degree is None else degree
    scales = e if cov3D_precomp is not None else t["scales"]
    rots = e if cov3D_precomp is not None else t["rotations"]
    cp = e if colors_precomp is None else colors_precomp
    cv = e if cov3D_precomp is None else cov3D_precomp
    rr = rb.RefRasterizer()
    R, B, color, radii = rr.forward(t["bg"], t["means3D"], cp, t["opacity"], scales, rots, 1.0, cv,
                                    t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy,
                                    cam.H, cam.W, t["dc"], sh, degree, t["campos"])
    torch.cuda.synchronize()
    dev = color.device
    out = dict(R=R, B=B, color=color, radii=radii)
    out.update(rr.geometry(dev))
    keys, vals = rr.binning(dev)
    out["keys"], out["ids"] = keys, vals
    out.update(rr.image(dev))
    if backward:
        g = rr.backward(t["bg"], t["means3D"], radii, cp, scales, rots, 1.0, cv, t["viewmatrix"],
                        t["projmatrix"], cam.tanfovx, cam.tanfovy, t["dL_dout"] if dL is None else dL,
                        t["dc"], sh, degree, t["campos"])
        torch.cuda.synchronize()
        out["grads"] = g
    out["_handle"] = rr
    return out


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().flatten()
    b = b.double().flatten()
    den = float(b.norm())
    num = float((a - b).norm())
    if den == 0.0:
        return 0.0 if num == 0.0 else float("inf")
    return num / den


def valid_prefix(keys: torch.Tensor):
    """Number of non-padding entries of a reference sorted key list (tile != 0xFFFFFFFF)."""
    tiles = (keys >> 32) & 0xFFFFFFFF
    return int((tiles != 0xFFFFFFFF).sum().item())
