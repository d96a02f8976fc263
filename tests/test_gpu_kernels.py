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


@pytest.mark.parametrize("rows,vis_frac", [(50_001, 0.15), (4099, 0.5), (130, 0.0), (130, 1.0)])
def test_visibility_split_adam_is_bit_identical_to_the_fused_step(rows, vis_frac):
    """gsr_adam_step_rows: phase ZERO_GRAD (rows with radii == 0, gradient not read) + phase
    WITH_GRAD (rendered rows) == one gsr_adam_step over everything, bit for bit."""
    from gsdf_slam_b200 import FusedAdam
    gen = torch.Generator(device="cuda").manual_seed(11)
    shapes = [(rows, 3), (rows, 1, 3), (rows, 15, 3), (rows, 1), (rows, 3), (rows, 4)]
    lrs = [1.6e-4, 0.0025, 0.000125, 0.05, 0.005, 0.001]
    pa = [torch.randn(s, device="cuda", generator=gen) for s in shapes]
    pb = [p.clone() for p in pa]
    oa = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15)
    ob = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pb, lrs)], lr=0.0, eps=1e-15)
    side = torch.cuda.Stream()
    for step in range(1, 5):
        radii = (torch.rand(rows, device="cuda", generator=gen) < vis_frac).to(torch.int32) * (step + 2)
        vis = radii > 0
        ga, gb = {}, {}
        for a, b in zip(pa, pb):
            g = torch.randn(a.shape, device="cuda", generator=gen)
            g[~vis] = 0.0                       # what gsr_backward writes for non-rendered rows
            ga[a], gb[b] = g, g.clone()
        # poison phase 0's view of the gradient: it must not be read there
        poison = {b: torch.full_like(g, float("nan")) for b, g in gb.items()}
        oa.step(grads=ga)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            ob.step_rows(radii, ob.ROWS_ZERO_GRAD, grads=poison, new_step=True)
        ob.step_rows(radii, ob.ROWS_WITH_GRAD, grads=gb)   # concurrently on the main stream
        torch.cuda.synchronize()
        for a, b in zip(pa, pb):
            assert torch.equal(a.view(torch.int32), b.view(torch.int32))
            assert torch.equal(oa.state[a]["exp_avg"].view(torch.int32), ob.state[b]["exp_avg"].view(torch.int32))
            assert torch.equal(oa.state[a]["exp_avg_sq"].view(torch.int32),
                               ob.state[b]["exp_avg_sq"].view(torch.int32))
            assert oa.state[a]["step"] == ob.state[b]["step"] == step


def test_masked_single_pass_adam_ignores_the_masked_out_gradient_rows():
    """GSR_ADAM_ROWS_ALL == gsr_adam_step on a zero-filled gradient, bit for bit, even when the
    rows with radii == 0 hold NaN (gsr_backward with rendered_rows_only leaves them untouched)."""
    from gsdf_slam_b200 import FusedAdam
    gen = torch.Generator(device="cuda").manual_seed(12)
    rows = 33_333
    shapes = [(rows, 3), (rows, 1, 3), (rows, 15, 3), (rows, 1), (rows, 3), (rows, 4)]
    lrs = [1.6e-4, 0.0025, 0.000125, 0.05, 0.005, 0.001]
    pa = [torch.randn(s, device="cuda", generator=gen) for s in shapes]
    pb = [p.clone() for p in pa]
    oa = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15)
    ob = FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pb, lrs)], lr=0.0, eps=1e-15)
    for step in range(1, 4):
        radii = (torch.rand(rows, device="cuda", generator=gen) < 0.2).to(torch.int32) * 7
        vis = radii > 0
        ga, gb = {}, {}
        for a, b in zip(pa, pb):
            g = torch.randn(a.shape, device="cuda", generator=gen)
            g[~vis] = 0.0
            ga[a] = g
            gp = g.clone()
            gp[~vis] = float("nan")
            gb[b] = gp
        oa.step(grads=ga)
        ob.step_rows(radii, ob.ROWS_ALL, grads=gb, new_step=True)
        torch.cuda.synchronize()
        for a, b in zip(pa, pb):
            assert torch.equal(a.view(torch.int32), b.view(torch.int32))
            assert torch.equal(oa.state[a]["exp_avg"].view(torch.int32), ob.state[b]["exp_avg"].view(torch.int32))
            assert torch.equal(oa.state[a]["exp_avg_sq"].view(torch.int32),
                               ob.state[b]["exp_avg_sq"].view(torch.int32))


def test_adam_step_rows_rejects_bad_arguments():
    L = _lib.lib()
    x = torch.zeros(12, device="cuda")
    mask = torch.zeros(5, dtype=torch.int32, device="cuda")
    grp = (_lib.AdamGroup * 1)(_lib.AdamGroup(param=x.data_ptr(), grad=x.data_ptr(), exp_avg=x.data_ptr(),
                                              exp_avg_sq=x.data_ptr(), numel=12, lr=1e-3, step=1))
    assert L.gsr_adam_step_rows(grp, 1, 0.9, 0.999, 1e-8, 1.0, 5, C.c_void_p(mask.data_ptr()), 0, None) != 0  # 12 % 5
    assert L.gsr_adam_step_rows(grp, 1, 0.9, 0.999, 1e-8, 1.0, 4, None, 0, None) != 0                          # no mask
    assert L.gsr_adam_step_rows(grp, 1, 0.9, 0.999, 1e-8, 1.0, 4, C.c_void_p(mask.data_ptr()), 3, None) != 0  # phase


def test_overlapped_mapping_step_equals_sequential_step():
    """KeyframeBatchMapper.step with the Adam update of the non-rendered rows overlapped on a
    side stream == accumulate() + one fused Adam step, bit for bit (1 and 2 local keyframes)."""
    from gsdf_slam_b200.mapping import SEGMENTS, Keyframe, KeyframeBatchMapper
    scene = scenes.tiny(seed=9, P=6000, W=176, H=120, degree=3)
    t = to_dev(scene)
    cam = scene.cam
    kfs = [Keyframe(t["viewmatrix"], t["projmatrix"], t["campos"], cam.tanfovx, cam.tanfovy, cam.H, cam.W,
                    t["dL_dout"] * s) for s in (1.0, -0.5)]
    params_init = dict(means3D=t["means3D"], dc=t["dc"], sh=t["sh"], opacity=t["opacity"], scales=t["scales"],
                       rotations=t["rotations"])
    for K in (1, 2):
        res = []
        for overlap in (False, True):
            params = dict(means3D=t["means3D"].clone(), dc=t["dc"].clone(), sh=t["sh"].clone(),
                          opacity=t["opacity"].clone(), scales=t["scales"].clone(), rotations=t["rotations"].clone())
            # reference arm: dense zero-filled gradients + one fused Adam launch
            m = KeyframeBatchMapper(params, scene.degree, t["bg"], freeze_render_inputs=True, overlap_adam=overlap,
                                    masked_single_view=False)
            for _ in range(3):
                m.step(kfs[:K])
            torch.cuda.synchronize()
            res.append(m)
        a, b = res
        hidden = a.last_radii == 0   # same camera for every keyframe: never rendered
        assert 0 < int(hidden.sum()) < scene.P
        for k in SEGMENTS:
            pa_, pb_ = a.opt_params[k], b.opt_params[k]
            sa, sb = a.optimizer.state[pa_], b.optimizer.state[pb_]
            assert sa["step"] == sb["step"] == 3
            # rows without gradient: exact; rendered rows: the blend backward sums with atomics, so
            # two runs agree to the gradient tolerance (1e-4 rel-L2), not bit for bit
            assert torch.equal(pa_[hidden].view(torch.int32), pb_[hidden].view(torch.int32)), k
            assert torch.equal(sa["exp_avg_sq"][hidden], sb["exp_avg_sq"][hidden]), k
            assert rel_l2(sa["exp_avg"], sb["exp_avg"]) <= 1e-4, k
            assert rel_l2(pa_ - params_init[k], pb_ - params_init[k]) <= 1e-3, k


def test_masked_single_view_step_equals_dense_step():
    """KeyframeBatchMapper.step on one keyframe: backward with rendered_rows_only + one masked Adam
    pass == zero-filled backward + fused Adam (rows never rendered: bit for bit)."""
    from gsdf_slam_b200.mapping import SEGMENTS, Keyframe, KeyframeBatchMapper
    scene = scenes.tiny(seed=10, P=7000, W=176, H=120, degree=3)
    t = to_dev(scene)
    cam = scene.cam
    kf = Keyframe(t["viewmatrix"], t["projmatrix"], t["campos"], cam.tanfovx, cam.tanfovy, cam.H, cam.W,
                  t["dL_dout"])
    init = dict(means3D=t["means3D"], dc=t["dc"], sh=t["sh"], opacity=t["opacity"], scales=t["scales"],
                rotations=t["rotations"])
    res = []
    for masked in (False, True):
        m = KeyframeBatchMapper({k: v.clone() for k, v in init.items()}, scene.degree, t["bg"],
                                freeze_render_inputs=True, masked_single_view=masked)
        if masked:
            m.arena.fill_(float("nan"))   # the non-rendered rows must never be read
        for _ in range(3):
            m.step([kf])
        torch.cuda.synchronize()
        res.append(m)
    a, b = res
    hidden = a.last_radii == 0
    assert 0 < int(hidden.sum()) < scene.P
    for k in SEGMENTS:
        pa_, pb_ = a.opt_params[k], b.opt_params[k]
        sa, sb = a.optimizer.state[pa_], b.optimizer.state[pb_]
        assert sa["step"] == sb["step"] == 3
        assert torch.isfinite(pb_).all(), k
        assert torch.equal(pa_[hidden].view(torch.int32), pb_[hidden].view(torch.int32)), k
        assert torch.equal(sa["exp_avg_sq"][hidden], sb["exp_avg_sq"][hidden]), k
        assert rel_l2(sa["exp_avg"], sb["exp_avg"]) <= 1e-4, k
        assert rel_l2(pa_ - init[k], pb_ - init[k]) <= 1e-3, k


def test_keyframe_batch_accumulation_equals_sum_of_single_view_gradients():
    """BASELINE configs[4] semantics on one rank: a batch of K keyframes accumulates the SUM of the
    K single-view gradients (== K backward() calls without zero_grad), then one Adam step."""
    from gsdf_slam_b200.mapping import SEGMENTS, Keyframe, KeyframeBatchMapper
    scene = scenes.room_scene(20_000, 320, 184, 160.0, 160.0, seed=5, name="room20k")
    cams = scenes.ring_cameras(4, 320, 184, 160.0, 160.0, radius=0.3)
    t = to_dev(scene)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    kfs = [Keyframe(tt(c.viewmatrix), tt(c.projmatrix), tt(c.campos), c.tanfovx, c.tanfovy, c.H, c.W, t["dL_dout"])
           for c in cams]
    params = dict(means3D=t["means3D"], dc=t["dc"], sh=t["sh"], opacity=t["opacity"], scales=t["scales"],
                  rotations=t["rotations"])
    m = KeyframeBatchMapper({k: v.clone() for k, v in params.items()}, scene.degree, t["bg"],
                            freeze_render_inputs=True)
    m.accumulate(kfs)
    torch.cuda.synchronize()
    total = {k: m.grad_views[k].clone() for k in SEGMENTS}
    ref = {k: torch.zeros_like(v) for k, v in total.items()}
    for kf in kfs:
        m.accumulate([kf])
        for k in SEGMENTS:
            ref[k] += m.grad_views[k]
    torch.cuda.synchronize()
    for k in SEGMENTS:
        assert rel_l2(total[k], ref[k]) <= 1e-5, k
    visible_anywhere = sum((m.render(kf)[3] > 0).long() for kf in kfs) > 0
    assert 0.2 < float(visible_anywhere.float().mean()) <= 1.0


def test_peer_exchange_matches_nccl_path_on_two_gpus():
    """gsr_adam_step_peer (fused reduction + Adam + broadcast over peer memory) against the NCCL
    all-reduce + fused Adam path; needs two GPUs (skipped on the single-GPU box)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "check_peer.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "peer exchange OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_two_devices_in_one_process_and_thread():
    """One host thread driving two GPUs (round-1 advisor finding): the read-back event, the pinned word and the
    capacity hints live in a per-device gsr_context, kernel attributes are set per device — the second device must
    render the same image and gradients as the first, including a scene with long per-tile lists (the kernel that
    needs the 80 KB dynamic shared-memory attribute)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from util import run_product_backward, run_product_forward
    for sc in (scenes.tiny(seed=3, P=6000), scenes.cfg4(1_200_000)):      # cfg4 at 1.2 M: per-tile lists beyond 4092
        outs = []
        for dev in ("cuda:0", "cuda:1", "cuda:0"):
            with torch.cuda.device(dev):
                t = to_dev(sc, dev)
                fwd = run_product_forward(sc, t)
                g = run_product_backward(sc, t, fwd)
                torch.cuda.synchronize(dev)
                outs.append((fwd[0], fwd[2].cpu(), fwd[3].cpu(), g["dL_dmeans3D"].cpu()))
        for o in outs[1:]:
            assert o[0] == outs[0][0]
            assert torch.equal(o[2], outs[0][2])                                   # radii
            assert float((o[1] - outs[0][1]).abs().max()) <= 1e-6                  # image (atomics-free: exact)
            assert rel_l2(o[3], outs[0][3]) <= 1e-5


@pytest.mark.parametrize("world", [2, 3, 8])
def test_peer_exchange_with_all_ranks_emulated_on_one_gpu(world):
    """The whole multi-rank exchange on ONE device (the driver's -m gpu box has one GPU): every "rank" owns its own
    gradient / parameter arena and flag buffer in ordinary device memory, the kernels run once per rank — rank masks
    from the flag bytes (rendered | gradient valid), then gsr_adam_step_peer on that rank's interleaved shard with
    shard-local moments, reading all ranks' gradient arenas and storing into all ranks' parameter arenas (unicast).
    Checked against the definition: Adam (LibTorch op sequence, moments bit-exact) on the rank-ordered sum of the
    gradients of the ranks whose gradient bit is set; the statistics segments sum / sum / max the deltas of the ranks
    whose RENDERED bit is set; all replicas end bit-identical; rows nobody flags are never read (NaN poison)."""
    import ctypes as C
    from gsdf_slam_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(100 + world)
    P = 70_001                                   # not a multiple of 4: segment padding and partial float4 groups
    widths = [3, 3, 45, 1, 3, 4]
    lrs = [1.6e-4, 0.0025, 0.0025 / 20, 0.05, 0.005, 0.001]
    starts, off = [], 0
    for w in widths:
        starts.append(off)
        off += (w * P + 3) // 4 * 4
    stride = (P + 3) // 4 * 4
    stat_off, arena_len = off, off + 3 * stride
    chunk = 65536
    rendered = torch.rand((world, P), device=dev, generator=gen) < 0.3          # rank k rendered the row
    has_grad = rendered & (torch.rand((world, P), device=dev, generator=gen) < 0.6)   # ... and got a gradient
    flags = [torch.zeros(P + 16, dtype=torch.uint8, device=dev) for _ in range(world)]
    for k in range(world):
        flags[k][:P] = rendered[k].to(torch.uint8) + 2 * has_grad[k].to(torch.uint8)
    grads = [torch.full((arena_len,), float("nan"), device=dev) for _ in range(world)]
    p0 = torch.zeros(arena_len, device=dev)
    ref_g = []
    for s, w in zip(starts, widths):
        p0[s:s + w * P] = torch.randn(w * P, device=dev, generator=gen)
        tot = torch.zeros((P, w), device=dev)
        for k in range(world):                                        # rank order, like the kernel
            gk = torch.randn((P, w), device=dev, generator=gen) * 1e-2
            grads[k][s:s + w * P].view(P, w)[has_grad[k]] = gk[has_grad[k]]
            tot = torch.where(has_grad[k][:, None], tot + gk, tot)
        ref_g.append(tot)
    # statistics: accumulated values in the parameter arena, this step's deltas in the gradient arenas
    stat0 = torch.rand((3, P), device=dev, generator=gen)
    ref_stat = stat0.clone()
    sums = torch.zeros((2, P), device=dev)
    mx = torch.zeros(P, device=dev)
    for j in range(3):
        p0[stat_off + j * stride: stat_off + j * stride + P] = stat0[j]
    for k in range(world):
        d = torch.rand((3, P), device=dev, generator=gen) * 5.0
        for j in range(3):
            grads[k][stat_off + j * stride: stat_off + j * stride + P][rendered[k]] = d[j][rendered[k]]
        sums = torch.where(rendered[k][None, :], sums + d[:2], sums)
        mx = torch.where(rendered[k], torch.maximum(mx, d[2]), mx)
    anyr = rendered.any(0)
    ref_stat[0] = torch.where(anyr, stat0[0] + sums[0], stat0[0])
    ref_stat[1] = torch.where(anyr, stat0[1] + sums[1], stat0[1])
    ref_stat[2] = torch.where(anyr, torch.maximum(stat0[2], mx), stat0[2])
    params = [p0.clone() for _ in range(world)]
    n_chunks = (arena_len + chunk - 1) // chunk
    moments = []
    for r in range(world):
        owned = (n_chunks - r + world - 1) // world if n_chunks > r else 0
        moments.append((torch.zeros(max(owned, 1) * chunk, device=dev), torch.zeros(max(owned, 1) * chunk, device=dev)))
    masks = torch.zeros(P + 4, dtype=torch.int32, device=dev)
    fptrs = (C.c_void_p * world)(*[f.data_ptr() for f in flags])
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.gsr_peer_rank_mask_u8(fptrs, world, P, C.c_void_p(masks.data_ptr()), stream), "rank_mask")
    want_mask = torch.zeros(P, dtype=torch.int32, device=dev)
    for k in range(world):
        want_mask |= has_grad[k].int() << k
        want_mask |= rendered[k].int() << (8 + k)
    assert torch.equal(masks[:P], want_mask)
    for step in (1, 2):
        for r in range(world):
            a = _lib.PeerAdamArgs()
            a.n_peers, a.rank = world, r
            for k in range(world):
                a.param_peers[k], a.grad_peers[k] = params[k].data_ptr(), grads[k].data_ptr()
            for i, (s0, w, lr) in enumerate(zip(starts, widths, lrs)):
                a.segments[i].start, a.segments[i].numel, a.segments[i].width, a.segments[i].lr = s0, w * P, w, lr
            for j, rule in enumerate((_lib.GSR_PEER_RULE_SUM, _lib.GSR_PEER_RULE_SUM, _lib.GSR_PEER_RULE_MAX)):
                sg = a.segments[6 + j]
                sg.start, sg.numel, sg.width, sg.rule = stat_off + j * stride, P, 1, rule
            a.n_segments = 9
            a.shard_lo, a.shard_hi, a.shard_chunk = 0, arena_len, chunk
            a.exp_avg, a.exp_avg_sq = moments[r][0].data_ptr(), moments[r][1].data_ptr()
            a.union_mask = masks.data_ptr()
            a.beta1, a.beta2, a.eps, a.grad_scale, a.step = 0.9, 0.999, 1e-15, 1.0, step
            a.stream = torch.cuda.current_stream().cuda_stream
            _lib.check(L.gsr_adam_step_peer(C.byref(a)), "gsr_adam_step_peer")
        if step == 1:   # the statistics after ONE exchange
            torch.cuda.synchronize()
            for j in range(3):
                got = params[0][stat_off + j * stride: stat_off + j * stride + P]
                assert torch.equal(got, ref_stat[j]), j
    torch.cuda.synchronize()
    for k in range(1, world):
        assert torch.equal(params[0].view(torch.int32), params[k].view(torch.int32))     # replicas bit-identical
    assert bool(torch.isfinite(params[0]).all())
    for s, w, lr, g in zip(starts, widths, lrs, ref_g):
        p = p0[s:s + w * P].view(P, w).clone()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        for step in (1, 2):
            _libtorch_adam_reference(p, g, m, v, lr, step)
        got = params[0][s:s + w * P].view(P, w)
        assert float((got - p).abs().max()) <= 2e-7 * max(1.0, float(p.abs().max())), w     # <= 2 ulp, like gsr_adam_step


def _brute_knn(pts: torch.Tensor) -> torch.Tensor:
    out = torch.empty(pts.size(0), device=pts.device)
    for s in range(0, pts.size(0), 2048):
        q = pts[s:s + 2048]
        d = ((q[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        d[torch.arange(q.size(0)), torch.arange(s, s + q.size(0))] = float("inf")
        best = d.topk(3, largest=False).values
        out[s:s + 2048] = best.sum(-1) / 3.0
    return out


@pytest.mark.parametrize("P", [4, 33, 1000, 1025, 40_000, 300_000])
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


@pytest.mark.parametrize("kind", ["plane_with_duplicates", "two_far_clusters", "line", "grid_ties"])
def test_knn_structured_clouds_bit_equal_to_reference(kind):
    """Clouds that stress the AABB-tree pruning: degenerate boxes, empty space, exact ties."""
    import gsdf_slam_b200 as gs
    from oracle import ref_binding as rb
    gen = torch.Generator(device="cuda").manual_seed(17)
    P = 50_000
    if kind == "plane_with_duplicates":
        pts = torch.rand((P, 3), device="cuda", generator=gen) * torch.tensor([8.0, 6.0, 0.0], device="cuda")
        pts[P // 2:] = pts[: P - P // 2]                     # every point exists twice
    elif kind == "two_far_clusters":
        pts = torch.randn((P, 3), device="cuda", generator=gen) * 0.05
        pts[P // 3:] += torch.tensor([100.0, -40.0, 7.0], device="cuda")
    elif kind == "line":
        pts = torch.zeros((P, 3), device="cuda")
        pts[:, 0] = torch.rand(P, device="cuda", generator=gen) * 50.0
    else:
        g = torch.arange(40, device="cuda", dtype=torch.float32) * 0.25
        pts = torch.stack(torch.meshgrid(g, g, g[:30], indexing="ij"), -1).reshape(-1, 3).contiguous()
    mine = gs.distCUDA2(pts)
    assert rel_l2(mine, _brute_knn(pts)) < 1e-5
    if rb.available():
        assert torch.equal(mine, rb.ref_knn(pts))


def test_knn_tiny_inputs():
    import gsdf_slam_b200 as gs
    assert gs.distCUDA2(torch.zeros((0, 3), device="cuda")).numel() == 0
    # fewer than 4 points: the reference yields FLT_MAX-based sums
    out = gs.distCUDA2(torch.tensor([[0.0, 0, 0], [1.0, 0, 0]], device="cuda"))
    assert torch.isinf(out).all() or (out > 1e37).all()
