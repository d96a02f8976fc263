"""Host-side mirror of the reference's frame-ingest entry point.

``colorize_launcher(points, colors, proj, Tr, pose, image)`` has the reference's name and argument
order (/root/reference/cuda_colorize/colorize_pointcloud.h:4-11, implementation
colorize_pointcloud.cu:57-107, called from submodules/data_loader/data_loader.cpp:189).  The reference
re-binds its ``points`` / ``colors`` arguments to the compacted survivors; Python has no reference
parameters, so the pair is returned.  One call into libgsr.so (csrc/ingest.cu); no eager fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .rasterizer import _stream


def colorize_launcher(points: torch.Tensor, colors: torch.Tensor, proj: torch.Tensor, Tr: torch.Tensor,
                      pose: torch.Tensor, image: torch.Tensor):
    """points [N,3] float (LiDAR frame); colors [N,3] float (overwritten in the reference, only its
    device matters here); proj [12] row-major 3x4; Tr, pose [16] row-major 4x4; image [H,W,3] float.
    Returns (points_world [K,3], colors [K,3]) of the K points that project into the image with
    camera z > 0.2, in input order."""
    L = _lib.lib()
    dev = points.device if points.is_cuda else torch.device("cuda")
    pts = points.to(device=dev, dtype=torch.float32).contiguous()
    if pts.dim() != 2 or pts.size(1) != 3:
        raise RuntimeError("colorize_launcher: points must have dimensions (num_points, 3)")
    img = image.to(device=dev, dtype=torch.float32).contiguous()
    if img.dim() != 3 or img.size(2) != 3:
        raise RuntimeError("colorize_launcher: image must have dimensions (H, W, 3)")
    mats = [m.to(device=dev, dtype=torch.float32).contiguous().reshape(-1) for m in (pose, proj, Tr)]
    if mats[0].numel() != 16 or mats[1].numel() != 12 or mats[2].numel() != 16:
        raise RuntimeError("colorize_launcher: pose/Tr must have 16 elements and proj 12")
    n = int(pts.size(0))
    H, W = int(img.size(0)), int(img.size(1))
    out_p = torch.empty((n, 3), dtype=torch.float32, device=dev)
    out_c = torch.empty((n, 3), dtype=torch.float32, device=dev)
    count = C.c_int64(0)
    if n:
        with torch.cuda.device(dev):
            ws = torch.empty(L.gsr_colorize_workspace_bytes(n), dtype=torch.uint8, device=dev)
            _lib.check(L.gsr_colorize_points(C.c_void_p(pts.data_ptr()), n, C.c_void_p(mats[0].data_ptr()),
                                             C.c_void_p(mats[1].data_ptr()), C.c_void_p(mats[2].data_ptr()),
                                             C.c_void_p(img.data_ptr()), W, H, C.c_void_p(out_p.data_ptr()),
                                             C.c_void_p(out_c.data_ptr()), C.c_void_p(ws.data_ptr()),
                                             C.byref(count), _stream(dev)), "gsr_colorize_points")
    k = int(count.value)
    return out_p[:k], out_c[:k]


def points_from_records(records, stride_floats: int = 4, device="cuda") -> torch.Tensor:
    """The [n,3] point tensor of a LiDAR frame from its on-disk records (KITTI .bin: x, y, z, intensity -> stride 4;
    anything buffer-like / numpy / a tensor).  Replaces the reference's per-element fill of a CUDA tensor
    (submodules/data_loader/data_loader.cpp:172-179: 3 n device operations) with ONE pinned host-to-device copy and
    one repack kernel (gsr_pack_xyz)."""
    import numpy as np
    L = _lib.lib()
    dev = torch.device(device)
    if isinstance(records, torch.Tensor):
        flat = records.reshape(-1).to(torch.float32)
    else:
        flat = torch.from_numpy(np.ascontiguousarray(np.frombuffer(records, dtype=np.float32)
                                                     if not isinstance(records, np.ndarray) else records,
                                                     dtype=np.float32).reshape(-1).copy())
    if flat.numel() % stride_floats != 0:
        raise RuntimeError("points_from_records: the buffer is not a whole number of records")
    n = flat.numel() // stride_floats
    if not flat.is_cuda:
        flat = flat.pin_memory().to(dev, non_blocking=True)
    out = torch.empty((n, 3), dtype=torch.float32, device=dev)
    if n:
        with torch.cuda.device(dev):
            _lib.check(L.gsr_pack_xyz(C.c_void_p(flat.data_ptr()), n, int(stride_floats), C.c_void_p(out.data_ptr()),
                                      _stream(dev)), "gsr_pack_xyz")
    return out
