"""TEST INFRASTRUCTURE ONLY: numpy/ctypes binding of the CPU oracle (oracle/cpu_ref.cpp,
built by oracle/build_oracle.sh).  Imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg — never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_cpu.so")
_lib = None


def build() -> str:
    res = subprocess.run(["bash", os.path.join(_HERE, "build_oracle.sh")], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the CPU oracle failed:\n" + res.stderr[-3000:])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.oracle_create.restype = C.c_void_p
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_forward.restype = C.c_int
        L.oracle_backward.restype = C.c_int
        L.oracle_num_rendered.restype = C.c_longlong
        L.oracle_num_rendered.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    if a is None or a.size == 0:
        return C.c_void_p(0)
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


class CpuOracle:
    def __init__(self, threads: int | None = None):
        self.L = lib()
        if threads is not None:
            self.L.oracle_set_threads(int(threads))
        self.h = C.c_void_p(self.L.oracle_create())

    def __del__(self):
        try:
            self.L.oracle_destroy(self.h)
        except Exception:
            pass

    @property
    def threads(self) -> int:
        return int(self.L.oracle_num_threads())

    def forward(self, scene, colors_precomp=None, cov3D_precomp=None, sh=None, degree=None):
        cam = scene.cam
        P, W, H = scene.P, cam.W, cam.H
        sh = scene.sh if sh is None else sh
        degree = scene.degree if degree is None else degree
        M = sh.shape[1] if (sh is not None and sh.ndim >= 2 and sh.shape[0] != 0) else 0
        self._args = dict(
            P=P, D=int(degree), M=int(M), bg=_f(scene.bg), means3D=_f(scene.means3D), dc=_f(scene.dc),
            sh=_f(sh) if M > 0 else None, colors=_f(colors_precomp), opacity=_f(scene.opacity),
            scales=None if cov3D_precomp is not None else _f(scene.scales),
            rots=None if cov3D_precomp is not None else _f(scene.rotations), cov=_f(cov3D_precomp),
            view=_f(cam.viewmatrix), proj=_f(cam.projmatrix), campos=_f(cam.campos),
            tanx=float(cam.tanfovx), tany=float(cam.tanfovy), W=W, H=H)
        a = self._args
        out_color = np.zeros((3, H, W), np.float32)
        radii = np.zeros(P, np.int32)
        R = C.c_longlong(0)
        rc = self.L.oracle_forward(
            self.h, C.c_int(P), C.c_int(a["D"]), C.c_int(a["M"]), _p(a["bg"]), C.c_int(W), C.c_int(H),
            _p(a["means3D"]), _p(a["dc"]), _p(a["sh"]), _p(a["colors"]), _p(a["opacity"]), _p(a["scales"]),
            C.c_float(1.0), _p(a["rots"]), _p(a["cov"]), _p(a["view"]), _p(a["proj"]), _p(a["campos"]),
            C.c_float(a["tanx"]), C.c_float(a["tany"]), _p(out_color), _p(radii), C.byref(R))
        assert rc == 0
        Rn = int(R.value)
        T = ((W + 15) // 16) * ((H + 15) // 16)
        keys = np.zeros(Rn, np.uint64)
        ids = np.zeros(Rn, np.uint32)
        self.L.oracle_get_binning(self.h, _p(keys), _p(ids))
        final_T = np.zeros(H * W, np.float32)
        n_contrib = np.zeros(H * W, np.uint32)
        ranges = np.zeros((T, 2), np.uint32)
        self.L.oracle_get_image(self.h, _p(final_T), _p(n_contrib), _p(ranges))
        geo = dict(depths=np.zeros(P, np.float32), means2D=np.zeros((P, 2), np.float32),
                   conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32),
                   tiles=np.zeros(P, np.uint32), clamped=np.zeros((P, 3), np.uint8))
        self.L.oracle_get_geometry(self.h, _p(geo["depths"]), _p(geo["means2D"]), _p(geo["conic_opacity"]),
                                   _p(geo["rgb"]), _p(geo["tiles"]), _p(geo["clamped"]))
        out = dict(R=Rn, color=out_color, radii=radii, keys=keys, ids=ids, final_T=final_T,
                   n_contrib=n_contrib, ranges=ranges)
        out.update(geo)
        return out

    def backward(self, dL_dout):
        a = self._args
        P, M = a["P"], a["M"]
        g = dict(dL_dmeans2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 2, 2), np.float32),
                 dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolors=np.zeros((P, 3), np.float32),
                 dL_dmeans3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
                 dL_ddc=np.zeros((P, 1, 3), np.float32), dL_dsh=np.zeros((P, M, 3), np.float32),
                 dL_dscales=np.zeros((P, 3), np.float32), dL_drotations=np.zeros((P, 4), np.float32))
        dL = _f(dL_dout)
        rc = self.L.oracle_backward(
            self.h, C.c_int(P), C.c_int(a["D"]), C.c_int(M), _p(a["bg"]), _p(a["means3D"]), _p(a["dc"]),
            _p(a["sh"]), _p(a["colors"]), _p(a["scales"]), C.c_float(1.0), _p(a["rots"]), _p(a["cov"]),
            _p(a["view"]), _p(a["proj"]), _p(a["campos"]), C.c_float(a["tanx"]), C.c_float(a["tany"]),
            _p(dL), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]),
            _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]), _p(g["dL_ddc"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]),
            _p(g["dL_drotations"]))
        assert rc == 0
        return g


def adam(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-15):
    """In-place LibTorch-Adam step on float32 numpy arrays."""
    L = lib()
    L.oracle_adam(_p(p), _p(g), _p(m), _p(v), C.c_longlong(p.size), C.c_double(lr), C.c_double(beta1),
                  C.c_double(beta2), C.c_double(eps), C.c_longlong(step))


def knn(points):
    pts = np.ascontiguousarray(points, np.float32)
    out = np.zeros(pts.shape[0], np.float32)
    lib().oracle_knn(C.c_int(pts.shape[0]), _p(pts), _p(out))
    return out


def mark_visible(means3D, viewmatrix):
    m = np.ascontiguousarray(means3D, np.float32)
    v = np.ascontiguousarray(viewmatrix, np.float32)
    out = np.zeros(m.shape[0], np.uint8)
    lib().oracle_mark_visible(C.c_int(m.shape[0]), _p(m), _p(v), _p(out))
    return out.astype(bool)


def colorize(points, image, proj, Tr, pose):
    """CPU restatement of colorize_launcher (cuda_colorize/colorize_pointcloud.cu:57-107):
    returns (world points [k,3], colours [k,3]) of the points that project into the image."""
    L = lib()
    L.oracle_colorize.restype = C.c_longlong
    pts = _f(points)
    img = _f(image)
    H, W = img.shape[0], img.shape[1]
    n = pts.shape[0]
    po = np.empty((n, 3), np.float32)
    co = np.empty((n, 3), np.float32)
    k = L.oracle_colorize(C.c_longlong(n), _p(pts), _p(_f(pose).ravel()), _p(_f(proj).ravel()), _p(_f(Tr).ravel()),
                          _p(img), C.c_int(W), C.c_int(H), _p(po), _p(co))
    return po[:k].copy(), co[:k].copy()
