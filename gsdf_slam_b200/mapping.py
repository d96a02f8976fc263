"""Keyframe-batch data-parallel mapping step.

The reference trains on ONE keyframe per iteration on one GPU
(/root/reference/src/gaussian_mapper.cpp:335-468: render -> loss -> backward -> Adam).  Here a
batch of keyframes is split disjointly across ranks (one process per GPU); every rank renders
and back-propagates its keyframes over a REPLICATED Gaussian set, the per-rank gradients are
summed with ONE NCCL all-reduce over a flat gradient arena, and every rank applies the
identical fused Adam step (gsr_adam_step) to its replica — no parameter broadcast is needed.

Gradient arena layout (fp32, each segment 16-byte aligned and itself a valid contiguous
gradient tensor): [xyz 3P | f_dc 3P | f_rest 3MP | opacity P | scaling 3P | rotation 4P].

The loss (L1 + SSIM) and the activation functions are outside this path (SURVEY.md §8f-1,
f-2): the step consumes dL/d(image) per keyframe and works on activated parameters.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from .optim import FusedAdam
from .rasterizer import RasterizeGaussiansBackwardCUDA, RasterizeGaussiansCUDA

SEGMENTS = ("means3D", "dc", "sh", "opacity", "scales", "rotations")
GRAD_NAME = {"means3D": "dL_dmeans3D", "dc": "dL_ddc", "sh": "dL_dsh", "opacity": "dL_dopacity",
             "scales": "dL_dscales", "rotations": "dL_drotations"}
# the reference's per-group learning rates (cfg/gaussian_mapper/*/*.yaml, gaussian_model.cpp:325-345)
DEFAULT_LR = {"means3D": 1.6e-4, "dc": 0.0025, "sh": 0.0025 / 20.0, "opacity": 0.05, "scales": 0.005,
              "rotations": 0.001}


def partition_keyframes(num_keyframes: int, world_size: int, rank: int) -> List[int]:
    """Rank r owns keyframes {r, r+n, r+2n, ...}."""
    return list(range(rank, num_keyframes, world_size))


def arena_layout(P: int, M: int) -> Dict[str, tuple]:
    """name -> (offset, numel, shape); offsets are multiples of 4 floats (16 bytes)."""
    shapes = {"means3D": (P, 3), "dc": (P, 1, 3), "sh": (P, M, 3), "opacity": (P, 1), "scales": (P, 3),
              "rotations": (P, 4)}
    out, off = {}, 0
    for name in SEGMENTS:
        n = 1
        for d in shapes[name]:
            n *= d
        out[name] = (off, n, shapes[name])
        off += (n + 3) // 4 * 4
    out["_total"] = (0, off, (off,))
    return out


@dataclass
class Keyframe:
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    campos: torch.Tensor
    tanfovx: float
    tanfovy: float
    H: int
    W: int
    dL_dout: torch.Tensor  # [3,H,W] gradient of the (out-of-scope) loss w.r.t. the rendered image


class KeyframeBatchMapper:
    """Replicated Gaussians + flat gradient arena + fused Adam; one instance per rank."""

    def __init__(self, params: Dict[str, torch.Tensor], degree: int, bg: torch.Tensor,
                 lr: Optional[Dict[str, float]] = None, eps: float = 1e-15,
                 process_group=None, freeze_render_inputs: bool = False):
        self.params = {k: params[k].contiguous() for k in SEGMENTS}
        self.degree = int(degree)
        self.bg = bg
        self.pg = process_group
        dev = self.params["means3D"].device
        self.P = int(self.params["means3D"].size(0))
        self.M = int(self.params["sh"].size(1)) if self.params["sh"].dim() >= 2 else 0
        self.layout = arena_layout(self.P, self.M)
        total = self.layout["_total"][1]
        self.arena = torch.zeros(total, dtype=torch.float32, device=dev)
        self._scratch_arena: Optional[torch.Tensor] = None
        self.grad_views = self._views(self.arena)
        # Adam updates `self.opt_params`; with freeze_render_inputs (benchmarks) these are shadow
        # copies, so every step renders the same scene while doing the full optimiser work.
        self.opt_params = ({k: v.clone() for k, v in self.params.items()} if freeze_render_inputs
                           else self.params)
        lr = dict(DEFAULT_LR, **(lr or {}))
        self.optimizer = FusedAdam([{"params": [self.opt_params[k]], "lr": lr[k]} for k in SEGMENTS
                                    if self.opt_params[k].numel() > 0], lr=0.0, eps=eps)
        self.last_radii: Optional[torch.Tensor] = None
        self.last_color: Optional[torch.Tensor] = None

    def _views(self, arena: torch.Tensor) -> Dict[str, torch.Tensor]:
        v = {}
        for name in SEGMENTS:
            off, n, shape = self.layout[name]
            v[name] = arena[off:off + n].view(shape)
        return v

    @property
    def world_size(self) -> int:
        import torch.distributed as dist
        return dist.get_world_size(self.pg) if (dist.is_available() and dist.is_initialized()) else 1

    def render_backward(self, kf: Keyframe, grad_views: Dict[str, torch.Tensor]):
        """fwd + bwd of one keyframe; parameter gradients are written into grad_views."""
        p = self.params
        e = torch.empty(0, dtype=torch.float32, device=p["means3D"].device)
        fwd = RasterizeGaussiansCUDA(self.bg, p["means3D"], e, p["opacity"], p["scales"], p["rotations"],
                                     1.0, e, kf.viewmatrix, kf.projmatrix, kf.tanfovx, kf.tanfovy, kf.H,
                                     kf.W, p["dc"], p["sh"], self.degree, kf.campos, False, False)
        R, B, color, radii, geom, binning, img, sample = fwd
        out = {GRAD_NAME[k]: grad_views[k] for k in SEGMENTS if grad_views[k].numel() > 0}
        g = RasterizeGaussiansBackwardCUDA(self.bg, p["means3D"], radii, e, p["scales"], p["rotations"], 1.0,
                                           e, kf.viewmatrix, kf.projmatrix, kf.tanfovx, kf.tanfovy,
                                           kf.dL_dout, p["dc"], p["sh"], self.degree, kf.campos, geom, R,
                                           binning, img, B, sample, False, want_dcov3D=False, out=out)
        self.last_radii, self.last_color = radii, color
        return color, radii, g

    def accumulate(self, keyframes: Sequence[Keyframe]) -> None:
        """Sum of the local keyframes' gradients -> arena (== K sequential backward() calls
        without zero_grad, the semantics pinned in SURVEY.md §8e)."""
        if len(keyframes) == 0:
            self.arena.zero_()
            return
        self.render_backward(keyframes[0], self.grad_views)      # first keyframe writes in place
        for kf in keyframes[1:]:
            if self._scratch_arena is None:
                self._scratch_arena = torch.zeros_like(self.arena)  # padding slots stay 0
            self.render_backward(kf, self._views(self._scratch_arena))
            self.arena.add_(self._scratch_arena)

    def allreduce(self) -> None:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1:
            dist.all_reduce(self.arena, op=dist.ReduceOp.SUM, group=self.pg)

    def adam_step(self, grad_scale: float = 1.0) -> None:
        grads = {self.opt_params[k]: self.grad_views[k] for k in SEGMENTS
                 if self.opt_params[k].numel() > 0}
        self.optimizer.step(grads=grads, grad_scale=grad_scale)

    def step(self, keyframes: Sequence[Keyframe], grad_scale: float = 1.0) -> None:
        """One mapping iteration over this rank's keyframes."""
        self.accumulate(keyframes)
        self.allreduce()
        self.adam_step(grad_scale)
