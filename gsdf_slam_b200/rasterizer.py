"""Host-side mirror of the reference's rasterizer interface, on top of the C ABI.

Same names, argument order and error behaviour as the reference's C++ layer:

* ``RasterizeGaussiansCUDA`` / ``RasterizeGaussiansBackwardCUDA`` / ``markVisible``
  — /root/reference/include/rasterize_points.h:18-39, 62-87, 112-115
  (implementation src/rasterize_points.cu:38-226)
* ``distCUDA2`` — submodules/simple-knn/spatial.h:14 (spatial.cu:15-26)
* ``GaussianRasterizationSettings`` / ``GaussianRasterizerFunction`` /
  ``rasterizeGaussians`` / ``GaussianRasterizer``
  — include/gaussian_rasterizer.h:26-132, src/gaussian_rasterizer.cpp:18-305

PyTorch is only the tensor container (device memory + current stream); all the
arithmetic happens in libgsr.so.  (The C++/LibTorch shim exporting the identical
C++ symbols for linking into the reference build lives in csrc/torch_shim/.)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib

NUM_CHANNELS = 3  # reference NUM_CHAFFELS (cuda_rasterizer/config.h:14)


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous()


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _Scratch:
    """A resizable device byte buffer handed to the C ABI through a callback — the
    analogue of the reference's resizeFunctional (src/rasterize_points.cu:28-36)."""

    def __init__(self, device):
        holder = [torch.empty(0, dtype=torch.uint8, device=device)]

        def _resize(_user, nbytes):
            # closes over `holder` only (no reference cycle through self: the buffers must
            # be freed by refcount as soon as autograd drops them, not by the cyclic GC)
            try:
                holder[0] = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
                return holder[0].data_ptr() if nbytes > 0 else 0
            except Exception:  # never let an exception cross the C ABI
                return 0

        self._holder = holder
        self.cb = _lib.RESIZE_FN(_resize)

    @property
    def t(self) -> torch.Tensor:
        return self._holder[0]


def RasterizeGaussiansCUDA(background, means3D, colors, opacity, scales, rotations, scale_modifier,
                           cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height,
                           image_width, dc, sh, degree, campos, prefiltered, debug,
                           radii_ready_event: Optional["torch.cuda.Event"] = None, raw_params: bool = False):
    """-> (rendered, num_buckets, out_color, radii, geomBuffer, binningBuffer, imgBuffer, sampleBuffer)

    ``radii_ready_event`` (extension): a ``torch.cuda.Event`` recorded on the current stream as
    soon as ``radii`` is final (after the first kernel of the forward pass), so that a side
    stream can start the zero-gradient phase of the visibility-split Adam step early.
    ``raw_params`` (extension): ``opacity`` / ``scales`` / ``rotations`` are the
    model's RAW parameters; sigmoid / exp / normalize (src/gaussian_model.cpp:48-71) are applied
    inside the preprocess kernel instead of by three LibTorch ops per view."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    L = _lib.lib()
    dev = means3D.device
    if dev.type != "cuda":
        raise RuntimeError("RasterizeGaussiansCUDA needs CUDA tensors (no CPU fallback)")
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    geom, binning, img, sample = (_Scratch(dev) for _ in range(4))
    M = int(sh.size(1)) if (sh is not None and sh.dim() >= 2 and sh.size(0) != 0) else 0

    keep = [_f32c(t) if t is not None else None for t in
            (background, means3D, dc, sh, colors, opacity, scales, rotations, cov3D_precomp,
             viewmatrix, projmatrix, campos)]
    (bg_c, means_c, dc_c, sh_c, col_c, op_c, sc_c, rot_c, cov_c, view_c, proj_c, cam_c) = keep
    rendered, buckets = C.c_int(0), C.c_int(0)
    a = _lib.ForwardArgs(
        P=P, D=int(degree), M=M, W=W, H=H,
        background=_ptr(bg_c), means3D=_ptr(means_c), dc=_ptr(dc_c), shs=_ptr(sh_c),
        colors_precomp=_ptr(col_c), opacities=_ptr(op_c), scales=_ptr(sc_c),
        scale_modifier=float(scale_modifier), rotations=_ptr(rot_c), cov3D_precomp=_ptr(cov_c),
        viewmatrix=_ptr(view_c), projmatrix=_ptr(proj_c), campos=_ptr(cam_c),
        tan_fovx=float(tan_fovx), tan_fovy=float(tan_fovy), prefiltered=int(bool(prefiltered)),
        debug=int(bool(debug)), out_color=_ptr(out_color), radii=_ptr(radii),
        geom_resize=geom.cb, geom_user=None, binning_resize=binning.cb, binning_user=None,
        image_resize=img.cb, image_user=None, sample_resize=sample.cb, sample_user=None,
        stream=_stream(dev), num_rendered=C.pointer(rendered), num_buckets=C.pointer(buckets),
        raw_params=int(bool(raw_params)), radii_ready_event=None)
    with torch.cuda.device(dev):
        if radii_ready_event is not None:
            if not radii_ready_event.cuda_event:      # torch creates the cudaEvent_t lazily
                radii_ready_event.record(torch.cuda.current_stream(dev))
            a.radii_ready_event = C.c_void_p(radii_ready_event.cuda_event)
        _lib.check(L.gsr_forward(C.byref(a)), "gsr_forward")
    return (rendered.value, buckets.value, out_color, radii, geom.t, binning.t, img.t, sample.t)


def RasterizeGaussiansBackwardCUDA(background, means3D, radii, colors, scales, rotations,
                                   scale_modifier, cov3D_precomp, viewmatrix, projmatrix, tan_fovx,
                                   tan_fovy, dL_dout_color, dc, sh, degree, campos, geomBuffer, R,
                                   binningBuffer, imageBuffer, B, sampleBuffer, debug,
                                   want_dcov3D: bool = True, out: Optional[dict] = None,
                                   rendered_rows_only: bool = False, raw_opacity: Optional[torch.Tensor] = None,
                                   accumulate: bool = False, accumulate_rows: Optional[torch.Tensor] = None,
                                   grad_rows: Optional[torch.Tensor] = None):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_ddc, dL_dsh,
    dL_dscales, dL_drotations)   (same order as src/rasterize_points.cu:204)

    ``out`` (extension): optional dict name -> preallocated contiguous tensor (e.g. views into
    the flat gradient arena that is all-reduced across GPUs) to receive a gradient in place.
    ``rendered_rows_only`` (extension): write only the rows of the rendered Gaussians and leave
    the others untouched (no dense zero fill) — for consumers that go through the ``radii`` mask
    (``FusedAdam.step_rows(..., ROWS_ALL)``).
    ``raw_opacity`` (extension, row f-2): the forward ran with ``raw_params``; ``scales`` /
    ``rotations`` are raw too and dL_dopacity / dL_dscales / dL_drotations come back w.r.t. the RAW
    parameters (chain rule through sigmoid / exp / normalize inside the backward preprocess).
    ``accumulate`` / ``accumulate_rows`` (extension, keyframe batches): add the parameter gradients to the rows'
    current values — to every rendered row, or only to the rows marked (> 0) in the int32 ``accumulate_rows``
    (the rows an earlier keyframe of the batch has written); see gsr_backward_args.
    ``grad_rows`` (extension): int32 [P] in/out row mask kept by the callee — rows with a value > 0 are added to
    (with ``accumulate``), and every rendered row that received a gradient is set to 1; rendered rows that no pixel
    blended are neither flagged nor (with ``rendered_rows_only``) written.  Zero it at the start of a batch and hand it
    to ``FusedAdam.step_rows`` as the row mask."""
    L = _lib.lib()
    dev = means3D.device
    P = int(means3D.size(0))
    for name, rows in (("accumulate_rows", accumulate_rows), ("grad_rows", grad_rows)):
        if rows is not None and (rows.dtype != torch.int32 or not rows.is_contiguous() or rows.numel() != P):
            raise RuntimeError(f"{name} must be a contiguous int32 tensor with one entry per Gaussian")
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    M = int(sh.size(1)) if (sh is not None and sh.dim() >= 2 and sh.size(0) != 0) else 0
    opts = dict(dtype=torch.float32, device=dev)
    # the callee writes every row of every output: no zero-fill needed here
    out = out or {}

    def alloc(name, shape):
        t = out.get(name)
        if t is not None:
            assert t.is_contiguous() and t.numel() == int(torch.Size(shape).numel()), name
            return t
        return (torch.empty if P > 0 else torch.zeros)(shape, **opts)

    dL_dmeans3D = alloc("dL_dmeans3D", (P, 3))
    dL_dmeans2D = alloc("dL_dmeans2D", (P, 3))
    dL_dcolors = alloc("dL_dcolors", (P, NUM_CHANNELS))
    dL_dopacity = alloc("dL_dopacity", (P, 1))
    dL_dcov3D = alloc("dL_dcov3D", (P, 6)) if want_dcov3D else None
    dL_ddc = alloc("dL_ddc", (P, 1, 3))
    dL_dsh = alloc("dL_dsh", (P, M, 3))
    dL_dscales = alloc("dL_dscales", (P, 3))
    dL_drotations = alloc("dL_drotations", (P, 4))
    if P != 0:
        keep = [_f32c(t) if t is not None else None for t in
                (background, means3D, dc, sh, colors, scales, rotations, cov3D_precomp, viewmatrix,
                 projmatrix, campos, dL_dout_color)]
        (bg_c, means_c, dc_c, sh_c, col_c, sc_c, rot_c, cov_c, view_c, proj_c, cam_c, dpix_c) = keep
        radii_c = radii.contiguous()
        raw_op_c = _f32c(raw_opacity) if raw_opacity is not None else None
        a = _lib.BackwardArgs(
            P=P, D=int(degree), M=M, R=int(R), B=int(B), W=W, H=H,
            background=_ptr(bg_c), means3D=_ptr(means_c), dc=_ptr(dc_c), shs=_ptr(sh_c),
            colors_precomp=_ptr(col_c), opacities=_ptr(raw_op_c), scales=_ptr(sc_c),
            scale_modifier=float(scale_modifier), rotations=_ptr(rot_c), cov3D_precomp=_ptr(cov_c),
            viewmatrix=_ptr(view_c), projmatrix=_ptr(proj_c), campos=_ptr(cam_c),
            tan_fovx=float(tan_fovx), tan_fovy=float(tan_fovy), radii=_ptr(radii_c),
            geom_buffer=_ptr(geomBuffer), binning_buffer=_ptr(binningBuffer),
            image_buffer=_ptr(imageBuffer), sample_buffer=_ptr(sampleBuffer), dL_dpix=_ptr(dpix_c),
            dL_dmean2D=_ptr(dL_dmeans2D), dL_dconic=None, dL_dopacity=_ptr(dL_dopacity),
            dL_dcolor=_ptr(dL_dcolors), dL_dmean3D=_ptr(dL_dmeans3D), dL_dcov3D=_ptr(dL_dcov3D),
            dL_ddc=_ptr(dL_ddc), dL_dsh=_ptr(dL_dsh), dL_dscale=_ptr(dL_dscales),
            dL_drot=_ptr(dL_drotations), debug=int(bool(debug)), stream=_stream(dev),
            raw_params=int(raw_opacity is not None), rendered_rows_only=int(bool(rendered_rows_only)),
            accumulate=int(bool(accumulate)),
            accumulate_rows=_ptr(accumulate_rows) if accumulate_rows is not None else None,
            grad_rows=_ptr(grad_rows) if grad_rows is not None else None)
        with torch.cuda.device(dev):
            _lib.check(L.gsr_backward(C.byref(a)), "gsr_backward")
    if dL_dcov3D is None:
        dL_dcov3D = torch.zeros((0,), **opts)
    return (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_ddc, dL_dsh,
            dL_dscales, dL_drotations)


def markVisible(means3D, viewmatrix, projmatrix):
    """bool[P]: view-space z > 0.2 (src/rasterize_points.cu:207-226)."""
    L = _lib.lib()
    P = int(means3D.size(0))
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        m, v, p = means3D.contiguous(), viewmatrix.contiguous(), projmatrix.contiguous()
        with torch.cuda.device(means3D.device):
            _lib.check(L.gsr_mark_visible(P, _ptr(m), _ptr(v), _ptr(p), _ptr(present),
                                          _stream(means3D.device)), "gsr_mark_visible")
    return present


def distCUDA2(points):
    """float[P]: mean squared distance to the 3 nearest neighbours (spatial.cu:15-26)."""
    L = _lib.lib()
    P = int(points.size(0))
    means = torch.zeros((P,), dtype=torch.float32, device=points.device)
    if P != 0:
        pts = points.contiguous()
        with torch.cuda.device(points.device):
            _lib.check(L.gsr_knn_mean_dist2(P, _ptr(pts), _ptr(means), _stream(points.device)),
                       "gsr_knn_mean_dist2")
    return means


@dataclass
class GaussianRasterizationSettings:
    """include/gaussian_rasterizer.h:26-63 (trailing underscores dropped)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    debug: bool = False


class GaussianRasterizerFunction(torch.autograd.Function):
    """src/gaussian_rasterizer.cpp:28-204: same input order, saved state and gradient order."""

    @staticmethod
    def forward(ctx, means3D, means2D, dc, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        (num_rendered, num_buckets, color, radii, geomBuffer, binningBuffer, imgBuffer,
         sampleBuffer) = RasterizeGaussiansCUDA(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
            rs.image_width, dc, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.num_buckets = num_buckets
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, dc,
                              sh, geomBuffer, binningBuffer, imgBuffer, sampleBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, dc, sh, geomBuffer,
         binningBuffer, imgBuffer, sampleBuffer) = ctx.saved_tensors
        (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_ddc, dL_dsh, dL_dscales,
         dL_drotations) = RasterizeGaussiansBackwardCUDA(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, dc,
            sh, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer,
            ctx.num_buckets, sampleBuffer, rs.debug)

        def g(t, like):
            return t if (like is not None and like.numel() != 0) else None

        return (dL_dmeans3D, dL_dmeans2D, g(dL_ddc, dc), g(dL_dsh, sh), g(dL_dcolors, colors_precomp),
                dL_dopacity, g(dL_dscales, scales), g(dL_drotations, rotations), g(dL_dcov3D, cov3Ds_precomp), None)


def rasterizeGaussians(means3D, means2D, dc, sh, colors_precomp, opacities, scales, rotations,
                       cov3Ds_precomp, raster_settings):
    return GaussianRasterizerFunction.apply(means3D, means2D, dc, sh, colors_precomp, opacities,
                                            scales, rotations, cov3Ds_precomp, raster_settings)


class GaussianRasterizer(torch.nn.Module):
    """include/gaussian_rasterizer.h:108-132, src/gaussian_rasterizer.cpp:18-26, 260-305."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisibleGaussians(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return markVisible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, dc=None, shs=None, colors_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None) -> Tuple[torch.Tensor, torch.Tensor]:
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise RuntimeError("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise RuntimeError("Please provide exactly one of either scale/rotation pair or "
                               "precomputed 3D covariance!")
        dev = means3D.device
        empty = lambda: torch.empty(0, dtype=torch.float32, device=dev)  # noqa: E731
        shs = empty() if shs is None else shs
        colors_precomp = empty() if colors_precomp is None else colors_precomp
        scales = empty() if scales is None else scales
        rotations = empty() if rotations is None else rotations
        cov3D_precomp = empty() if cov3D_precomp is None else cov3D_precomp
        dc = empty() if dc is None else dc
        color, radii = rasterizeGaussians(means3D, means2D, dc, shs, colors_precomp, opacities,
                                          scales, rotations, cov3D_precomp, self.raster_settings)
        return color, radii
