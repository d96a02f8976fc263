"""Fused multi-tensor Adam behind the interface the reference uses.

The reference steps ``torch::optim::Adam`` over six parameter groups with
per-group learning rates, eps = 1e-15, betas (0.9, 0.999), no weight decay
(/root/reference/src/gaussian_model.cpp:319-352; ``optimizer_->step()`` at
src/gaussian_mapper.cpp:465).  ``FusedAdam`` keeps that interface
(``param_groups``, ``state``, ``step()``, ``zero_grad(set_to_none)``) and performs
the whole step with ONE kernel launch through ``gsr_adam_step``.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List

import torch

from . import _lib


class FusedAdam:
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        groups = list(params)
        if len(groups) == 0:
            raise ValueError("optimizer got an empty parameter list")
        if not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups: List[dict] = []
        for g in groups:
            g = dict(g)
            ps = g["params"]
            g["params"] = [ps] if isinstance(ps, torch.Tensor) else list(ps)
            g.setdefault("lr", lr)
            g.setdefault("betas", tuple(betas))
            g.setdefault("eps", eps)
            self.param_groups.append(g)
        self.state = {}

    def _state(self, p: torch.Tensor) -> dict:
        st = self.state.get(p)
        if st is None:
            st = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
            self.state[p] = st
        return st

    def _collect(self, grads, bump_step: bool, need_grad: bool = True):
        """{(beta1, beta2, eps, device): [AdamGroup, ...]} over every parameter with a gradient."""
        by_cfg = {}
        for g in self.param_groups:
            for p in g["params"]:
                grad = (grads.get(p) if grads is not None else p.grad)
                if grad is None and need_grad:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()
                        and (grad is None or grad.is_contiguous())):
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA tensors")
                st = self._state(p)
                if bump_step:
                    st["step"] += 1
                key = (g["betas"][0], g["betas"][1], g["eps"], p.device)
                by_cfg.setdefault(key, []).append(
                    _lib.AdamGroup(param=p.data_ptr(), grad=grad.data_ptr() if grad is not None else None,
                                   exp_avg=st["exp_avg"].data_ptr(),
                                   exp_avg_sq=st["exp_avg_sq"].data_ptr(), numel=p.numel(),
                                   lr=float(g["lr"]), step=int(st["step"])))
        return by_cfg

    @torch.no_grad()
    def step(self, grads=None, grad_scale: float = 1.0) -> None:
        """One Adam step.  ``grads`` optionally maps parameter -> gradient tensor
        (e.g. views into the all-reduced flat arena); default is ``p.grad``."""
        L = _lib.lib()
        for (b1, b2, eps, dev), items in self._collect(grads, bump_step=True).items():
            for i in range(0, len(items), 16):
                chunk = items[i:i + 16]
                arr = (_lib.AdamGroup * len(chunk))(*chunk)
                with torch.cuda.device(dev):
                    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                    _lib.check(L.gsr_adam_step(arr, len(chunk), float(b1), float(b2), float(eps),
                                               float(grad_scale), stream), "gsr_adam_step")

    ROWS_ZERO_GRAD = 0   # gsr_capi.h: GSR_ADAM_ROWS_ZERO_GRAD
    ROWS_WITH_GRAD = 1   # gsr_capi.h: GSR_ADAM_ROWS_WITH_GRAD
    ROWS_ALL = 2         # gsr_capi.h: GSR_ADAM_ROWS_ALL

    @torch.no_grad()
    def step_rows(self, row_mask: torch.Tensor, phase: int, grads=None, grad_scale: float = 1.0,
                  new_step: bool = False) -> None:
        """One phase of the visibility-split step (``gsr_adam_step_rows``) on the current stream.

        Every parameter must be a ``[rows, ...]`` tensor with ``rows == row_mask.numel()``;
        ``row_mask`` is int32 (> 0: the row received a gradient, e.g. the rasterizer's ``radii``).
        ``phase=ROWS_ZERO_GRAD`` updates the other rows with a gradient of exactly zero and does
        not read ``grads`` (so it can run before / while the backward pass executes);
        ``phase=ROWS_WITH_GRAD`` updates the marked rows; ``phase=ROWS_ALL`` updates every row in
        one pass, reading ``grads`` only for the marked rows (the other rows of ``grads`` may be
        uninitialised).  Pass ``new_step=True`` to exactly one call of an iteration (the first):
        it advances the step counters."""
        L = _lib.lib()
        if row_mask.dtype != torch.int32 or not row_mask.is_contiguous():
            raise RuntimeError("row_mask must be a contiguous int32 tensor")
        rows = int(row_mask.numel())
        need_grad = phase != self.ROWS_ZERO_GRAD
        for (b1, b2, eps, dev), items in self._collect(grads, bump_step=new_step, need_grad=need_grad).items():
            for i in range(0, len(items), 16):
                chunk = items[i:i + 16]
                arr = (_lib.AdamGroup * len(chunk))(*chunk)
                with torch.cuda.device(dev):
                    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                    _lib.check(L.gsr_adam_step_rows(arr, len(chunk), float(b1), float(b2), float(eps),
                                                    float(grad_scale), rows, C.c_void_p(row_mask.data_ptr()),
                                                    int(phase), stream), "gsr_adam_step_rows")

    def zero_grad(self, set_to_none: bool = True) -> None:
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()
