"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libgsdf_ref.so — the
UNMODIFIED reference CUDA rasterizer compiled for sm_100a (oracle/build_ref.sh,
oracle/ref_capi.cu).  Used by tests/ (bit-exact GPU parity), by
tests/golden/make_golden.py and by ``bench.py --impl reference``.  Never imported
by the product package.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB_PATH = os.path.join(_HERE, "_ref", "libgsdf_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(REF_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_LIB_PATH)
        L.ref_create.restype = C.c_void_p
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_forward.restype = C.c_int
        L.ref_backward.restype = C.c_int
        for n in ("ref_get_geometry", "ref_get_binning", "ref_get_image", "ref_mark_visible", "ref_knn"):
            getattr(L, n).restype = C.c_int
        _lib = L
    return _lib


def _p(t):
    if t is None or t.numel() == 0:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


class RefRasterizer:
    """One reference rasterizer state (its four scratch buffers live in the handle)."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.ref_create())
        self.P = self.R = self.B = self.W = self.H = 0

    def __del__(self):
        try:
            self.L.ref_destroy(self.h)
        except Exception:
            pass

    def forward(self, bg, means3D, colors_precomp, opacity, scales, rotations, scale_modifier,
                cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, H, W, dc, sh, degree,
                campos, prefiltered=False, debug=False, out_color=None, radii=None):
        P = means3D.size(0)
        dev = means3D.device
        M = sh.size(1) if (sh is not None and sh.dim() >= 2 and sh.size(0) != 0) else 0
        if out_color is None:
            out_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        if radii is None:
            radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        R, B = C.c_int(0), C.c_int(0)
        rc = self.L.ref_forward(
            self.h, C.c_int(P), C.c_int(degree), C.c_int(M), _p(bg), C.c_int(W), C.c_int(H),
            _p(means3D), _p(dc), _p(sh), _p(colors_precomp), _p(opacity), _p(scales),
            C.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp), _p(viewmatrix),
            _p(projmatrix), _p(campos), C.c_float(tanfovx), C.c_float(tanfovy),
            C.c_int(int(prefiltered)), _p(out_color), _p(radii), C.c_int(int(debug)),
            C.byref(R), C.byref(B))
        if rc != 0:
            raise RuntimeError("ref_forward failed")
        self.P, self.R, self.B, self.W, self.H = P, R.value, B.value, W, H
        return R.value, B.value, out_color, radii

    def backward(self, bg, means3D, radii, colors_precomp, scales, rotations, scale_modifier,
                 cov3D_precomp, viewmatrix, projmatrix, tanfovx, tanfovy, dL_dout, dc, sh, degree,
                 campos, debug=False):
        P = means3D.size(0)
        dev = means3D.device
        H, W = dL_dout.size(1), dL_dout.size(2)
        M = sh.size(1) if (sh is not None and sh.dim() >= 2 and sh.size(0) != 0) else 0
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
        # same zero-filled allocations as the reference shim (src/rasterize_points.cu:157-166)
        g = dict(dL_dmeans3D=z(P, 3), dL_dmeans2D=z(P, 3), dL_dcolors=z(P, 3), dL_dconic=z(P, 2, 2),
                 dL_dopacity=z(P, 1), dL_dcov3D=z(P, 6), dL_ddc=z(P, 1, 3), dL_dsh=z(P, M, 3),
                 dL_dscales=z(P, 3), dL_drotations=z(P, 4))
        rc = self.L.ref_backward(
            self.h, C.c_int(P), C.c_int(degree), C.c_int(M), C.c_int(self.R), C.c_int(self.B), _p(bg),
            C.c_int(W), C.c_int(H), _p(means3D), _p(dc), _p(sh), _p(colors_precomp), _p(scales),
            C.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp), _p(viewmatrix),
            _p(projmatrix), _p(campos), C.c_float(tanfovx), C.c_float(tanfovy), _p(radii),
            _p(dL_dout), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]),
            _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_ddc"]),
            _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]), C.c_int(int(debug)))
        if rc != 0:
            raise RuntimeError("ref_backward failed")
        return g

    # ---- decoders of the reference's opaque buffers ----
    def geometry(self, dev):
        P = self.P
        out = dict(depths=torch.zeros(P, dtype=torch.float32, device=dev),
                   means2D=torch.zeros((P, 2), dtype=torch.float32, device=dev),
                   conic_opacity=torch.zeros((P, 4), dtype=torch.float32, device=dev),
                   rgb=torch.zeros((P, 3), dtype=torch.float32, device=dev),
                   tiles_touched=torch.zeros(P, dtype=torch.int32, device=dev),
                   clamped=torch.zeros((P, 3), dtype=torch.uint8, device=dev),
                   cov3D=torch.zeros((P, 6), dtype=torch.float32, device=dev),
                   point_offsets=torch.zeros(P, dtype=torch.int32, device=dev))
        rc = self.L.ref_get_geometry(self.h, _p(out["depths"]), _p(out["means2D"]),
                                     _p(out["conic_opacity"]), _p(out["rgb"]),
                                     _p(out["tiles_touched"]), _p(out["clamped"]), _p(out["cov3D"]),
                                     _p(out["point_offsets"]))
        assert rc == 0
        return out

    def binning(self, dev):
        R = self.R
        keys = torch.zeros(R, dtype=torch.int64, device=dev)
        vals = torch.zeros(R, dtype=torch.int32, device=dev)
        rc = self.L.ref_get_binning(self.h, _p(keys), _p(vals), C.c_void_p(0), C.c_void_p(0))
        assert rc == 0
        return keys, vals

    def image(self, dev):
        N = self.W * self.H
        T = ((self.W + 15) // 16) * ((self.H + 15) // 16)
        out = dict(final_T=torch.zeros(N, dtype=torch.float32, device=dev),
                   n_contrib=torch.zeros(N, dtype=torch.int32, device=dev),
                   ranges=torch.zeros((T, 2), dtype=torch.int32, device=dev),
                   max_contrib=torch.zeros(T, dtype=torch.int32, device=dev),
                   bucket_offsets=torch.zeros(T, dtype=torch.int32, device=dev))
        rc = self.L.ref_get_image(self.h, _p(out["final_T"]), _p(out["n_contrib"]), _p(out["ranges"]),
                                  _p(out["max_contrib"]), _p(out["bucket_offsets"]))
        assert rc == 0
        return out


def ref_knn(points: torch.Tensor) -> torch.Tensor:
    P = points.size(0)
    out = torch.zeros(P, dtype=torch.float32, device=points.device)
    pts = points.contiguous().clone()
    assert lib().ref_knn(C.c_int(P), _p(pts), _p(out)) == 0
    torch.cuda.synchronize()
    return out


def ref_mark_visible(means3D, viewmatrix, projmatrix):
    P = means3D.size(0)
    out = torch.zeros(P, dtype=torch.uint8, device=means3D.device)
    assert lib().ref_mark_visible(C.c_int(P), _p(means3D), _p(viewmatrix), _p(projmatrix), _p(out)) == 0
    torch.cuda.synchronize()
    return out.bool()
