"""ctypes binding of the C entry for the keyframe-batch mapping step (include/gsr_mapper.h, csrc/mapper.cu).

The C++ side owns everything a C++ gaussian_mapper.cpp would need for the single- or multi-GPU step: peer-mapped
arenas (CUDA IPC), the cross-GPU barrier, the per-keyframe calls and the fused peer-memory exchange.  This wrapper
exists so that the tests (and Python users) can drive exactly that code path; the only thing it adds is the
all-gather of the GSR_MAPPER_HANDLE_BYTES blobs through torch.distributed at start-up.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib
from .mapping import DEFAULT_LR, Keyframe

SEG = {"means3D": 0, "dc": 1, "opacity": 2, "scales": 3, "rotations": 4, "sh": 5}     # GSR_SEG_*
STAT = {"max_radii2D": 0, "xyz_gradient_accum": 1, "denom": 2}                      # GSR_STAT_*


class _DevArray:
    """Zero-copy torch view of device memory owned by the C mapper (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _view(ptr: int, n: int, device) -> torch.Tensor:
    with torch.cuda.device(device):
        return torch.as_tensor(_DevArray(ptr, n), device=device)


class CMapper:
    def __init__(self, params: Dict[str, torch.Tensor], degree: int, bg: torch.Tensor, rank: int = 0, world: int = 1,
                 process_group=None, lr: Optional[Dict[str, float]] = None, eps: float = 1e-15, raw_params: bool = False,
                 with_stats: bool = True, lambda_dssim: float = 0.2):
        L = _lib.lib()
        self.L = L
        dev = params["means3D"].device
        self.device = dev
        self.P = int(params["means3D"].size(0))
        self.M = int(params["sh"].size(1)) if params["sh"].dim() >= 2 else 0
        self.bg = bg.contiguous()
        cfg = _lib.MapperConfig()
        cfg.P, cfg.M, cfg.sh_degree, cfg.rank, cfg.world = self.P, self.M, int(degree), int(rank), int(world)
        lr = dict(DEFAULT_LR, **(lr or {}))
        for name, k in SEG.items():
            cfg.lr[k] = float(lr[name])
        cfg.beta1, cfg.beta2, cfg.eps = 0.9, 0.999, float(eps)
        cfg.raw_params, cfg.with_stats, cfg.lambda_dssim = int(raw_params), int(with_stats), float(lambda_dssim)
        blob = C.create_string_buffer(_lib.GSR_MAPPER_HANDLE_BYTES)
        with torch.cuda.device(dev):
            self.h = L.gsr_mapper_create(C.byref(cfg), blob)
        if not self.h:
            raise RuntimeError("gsr_mapper_create failed")
        if world > 1:
            import torch.distributed as dist
            blobs = [None] * world
            dist.all_gather_object(blobs, bytes(blob.raw), group=process_group)
            allb = C.create_string_buffer(b"".join(blobs), world * _lib.GSR_MAPPER_HANDLE_BYTES)
            with torch.cuda.device(dev):
                _lib.check(L.gsr_mapper_connect(C.c_void_p(self.h), allb), "gsr_mapper_connect")
            dist.barrier(group=process_group)
        widths = {"means3D": 3, "dc": 3, "opacity": 1, "scales": 3, "rotations": 4, "sh": 3 * self.M}
        self.param_ptr = {}
        for name, k in SEG.items():
            n = self.P * widths[name]
            if n == 0:
                continue
            ptr = L.gsr_mapper_param(C.c_void_p(self.h), k)
            self.param_ptr[name] = (ptr, n, tuple(params[name].shape))
            # the replica lives in the mapper's (peer-mapped) arena: views of it are ordinary tensors
            _view(ptr, n, dev).view(params[name].shape).copy_(params[name])
        self.with_stats = bool(with_stats)
        torch.cuda.synchronize(dev)

    def param_views(self) -> Dict[str, torch.Tensor]:
        """LIVE zero-copy views of this rank's parameter replica (valid until close())."""
        return {name: _view(ptr, n, self.device).view(shape) for name, (ptr, n, shape) in self.param_ptr.items()}

    def params(self) -> Dict[str, torch.Tensor]:
        """Copies of this rank's parameter replica."""
        torch.cuda.synchronize(self.device)
        return {k: v.clone() for k, v in self.param_views().items()}

    def stats(self) -> Dict[str, torch.Tensor]:
        if not self.with_stats:
            return {}
        torch.cuda.synchronize(self.device)
        return {name: _view(self.L.gsr_mapper_stat(C.c_void_p(self.h), k), self.P, self.device).clone()
                for name, k in STAT.items()}

    def set_lr(self, name: str, lr: float) -> None:
        _lib.check(self.L.gsr_mapper_set_lr(C.c_void_p(self.h), SEG[name], float(lr)), "gsr_mapper_set_lr")

    def step(self, keyframes: Sequence[Keyframe], grad_scale: float = 1.0, out_colors: Optional[Sequence[torch.Tensor]] = None,
             loss_out: Optional[torch.Tensor] = None) -> None:
        n = len(keyframes)
        arr = (_lib.MapperKeyframe * max(n, 1))()
        keep = []
        for i, kf in enumerate(keyframes):
            v, p, c = kf.viewmatrix.contiguous(), kf.projmatrix.contiguous(), kf.campos.contiguous()
            keep += [v, p, c]
            a = arr[i]
            a.viewmatrix, a.projmatrix, a.campos = v.data_ptr(), p.data_ptr(), c.data_ptr()
            a.tanfovx, a.tanfovy, a.W, a.H = float(kf.tanfovx), float(kf.tanfovy), int(kf.W), int(kf.H)
            a.dL_dpix = kf.dL_dout.data_ptr() if kf.dL_dout is not None else None
            a.gt_image = kf.gt_image.data_ptr() if kf.gt_image is not None else None
            a.mask = kf.mask.data_ptr() if kf.mask is not None else None
            a.out_color = out_colors[i].data_ptr() if out_colors is not None else None
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(self.L.gsr_mapper_step(C.c_void_p(self.h), arr, n, C.c_void_p(self.bg.data_ptr()), float(grad_scale),
                                              C.c_void_p(loss_out.data_ptr()) if loss_out is not None else None, stream),
                       "gsr_mapper_step")

    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.gsr_mapper_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
