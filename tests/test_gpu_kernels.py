"""GPU unit tests of the stand-alone kernels: onesweep radix sort, fused Adam, kNN."""
import ctypes as C

import numpy as np
import pytest
import torch

from util import _lib, rel_l2, scenes, to_dev

pytestmark = pytest.mark.gpu


def _sort(items: torch.Tensor, begin: int, end: int) -> torch.Tensor:
    L = _lib.lib()
    n = items.numel()
    out = torch.empty_like(items)
    alt = torch.empty_like(items)
    tmp = torch.empty(int(L.gsr_sort_temp_bytes(n)), dtype=torch.uint8, device=items.device)
    p = lambda x: C.c_void_p(x.data_ptr()) if x.numel() else None  # noqa: E731
    _lib.check(L.gsr_sort_u64(p(items), p(out), p(alt), n, begin, end, p(tmp), None), "gsr_sort_u64")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 100_003, 3_000_001])
@pytest.mark.parametrize("bits", [(32, 64), (32, 44), (0, 64), (32, 36), (8, 24)])
def test_onesweep_sort_is_a_stable_sort_on_the_selected_bits(n, bits):
    begin, end = bits
    g = torch.Generator(device="cuda").manual_seed(n + begin * 7 + end)
    hi = torch.randint(0, 2 ** 31 - 1, (n,), device="cuda", generator=g, dtype=torch.int64)
    lo = torch.randint(0, 2 ** 31 - 1, (n,), device="cuda", generator=g, dtype=torch.int64)
    # few distinct digits in the selected range -> many ties -> stability is exercised
    items = ((hi % 5003) << 32) | (lo % 977) | ((hi >> 20) << 52)
    got = _sort(items.clone(), begin, end)
    width = end - begin
    mask = (1 << width) - 1 if width < 64 else -1
    if width == 64:
        # compare as unsigned: flip the sign bit for the reference ordering
        key = items ^ (-(1 << 63))
    else:
        key = (items >> begin) & mask
    order = torch.sort(key, stable=True).indices
    expect = items[order]
    assert torch.equal(got, expect)


def test_sort_rejects_bad_arguments():
    L = _lib.lib()
    x = torch.zeros(8, dtype=torch.int64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    assert L.gsr_sort_u64(p(x), p(x), p(x), 8, 0, 64, p(x), None) != 0   # aliasing
    assert L.gsr_sort_u64(p(x), p(x), None, 8, 0, 70, p(x), None) != 0   # bad bit range


def _libtorch_adam_reference(p, g, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-15):
    """torch::optim::Adam::step, op by op (torch/csrc/api/src/optim/adam.cpp), on the GPU."""
    bias_correction1 = 1 - beta1 ** step
    bias_correction2 = 1 - beta2 ** step
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    denom = (v.sqrt() / np.sqrt(bias_correction2)).add_(eps)
    step_size = lr / bias_correction1
    p.addcdiv_(m, denom, value=-step_size)


def test_fused_adam_matches_libtorch_adam_op_sequence():
    from gsdf_slam_b200 import FusedAdam
    gen = torch.Generator(device="cuda").manual_seed(5)
    shapes = [(50_001, 3), (50_001, 1, 3), (50_001, 15, 3), (50_001, 1), (50_001, 3), (50_001, 4)]
    lrs = [1.6e-4, 0.0025, 0.000125, 0.05, 0.005, 0.001]
    params = [torch.randn(s, device="cuda", generator=gen) for s in shapes]
    ref_p = [p.clone() for p in params]
    ref_m = [torch.zeros_like(p) for p in params]
    ref_v = [torch.zeros_like(p) for p in params]
    opt = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(params, lrs)], lr=0.0, eps=1e-15)
    max_ulp = 0
    for step in range(1, 6):
        grads = {}
        for i, p in enumerate(params):
            gten = torch.randn(p.shape, device="cuda", generator=gen) * (10.0 ** (-(i % 4)))
            gten[::7] = 0.0  # invisible Gaussians: zero gradient still decays m, v and moves p
            grads[p] = gten
            _libtorch_adam_reference(ref_p[i], gten, ref_m[i], ref_v[i], lrs[i], step)
        opt.step(grads=grads)
        torch.cuda.synchronize()
        for i, p in enumerate(params):
            ulp = (p.view(torch.int32).long() - ref_p[i].view(torch.int32).long()).abs().max().item()
            max_ulp = max(max_ulp, ulp)
            st = opt.state[p]
            assert torch.equal(st["exp_avg"], ref_m[i])
            assert torch.equal(st["exp_avg_sq"], ref_v[i])
    assert max_ulp <= 2, f"parameters differ from LibTorch Adam by {max_ulp} ulp"


def _brute_knn(pts: torch.Tensor) -> torch.Tensor:
    out = torch.empty(pts.size(0), device=pts.device)
    for s in range(0, pts.size(0), 2048):
        q = pts[s:s + 2048]
        d = ((q[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        d[torch.arange(q.size(0)), torch.arange(s, s + q.size(0))] = float("inf")
        best = d.topk(3, largest=False).values
        out[s:s + 2048] = best.sum(-1) / 3.0
    return out


@pytest.mark.parametrize("P", [4, 1000, 1025, 40_000])
def test_knn_mean_dist2(P):
    import gsdf_slam_b200 as gs
    gen = torch.Generator(device="cuda").manual_seed(P)
    pts = torch.randn((P, 3), device="cuda", generator=gen) * torch.tensor([4.0, 1.0, 3.0], device="cuda")
    mine = gs.distCUDA2(pts)
    torch.cuda.synchronize()
    brute = _brute_knn(pts)
    assert rel_l2(mine, brute) < 1e-5
    # permutation invariance: the result is a per-point quantity
    perm = torch.randperm(P, device="cuda", generator=gen)
    mine_p = gs.distCUDA2(pts[perm].contiguous())
    assert torch.equal(mine_p, mine[perm])
    from oracle import ref_binding as rb
    if rb.available():
        ref = rb.ref_knn(pts)
        assert torch.equal(mine, ref), f"max diff {(mine - ref).abs().max().item()}"


def test_knn_tiny_inputs():
    import gsdf_slam_b200 as gs
    assert gs.distCUDA2(torch.zeros((0, 3), device="cuda")).numel() == 0
    # fewer than 4 points: the reference yields FLT_MAX-based sums
    out = gs.distCUDA2(torch.tensor([[0.0, 0, 0], [1.0, 0, 0]], device="cuda"))
    assert torch.isinf(out).all() or (out > 1e37).all()
