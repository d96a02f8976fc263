"""CPU tests of the keyframe-batch data-parallel logic (gsdf_slam_b200/mapping.py) with
world_size 2 over gloo: keyframe partition, flat gradient arena, all-reduce semantics
(n-rank result == 1-rank sequential accumulation) and replica consistency after Adam.
The render/backward and the fused Adam need a GPU, so they are replaced here by a
deterministic fake render and by the CPU oracle's Adam (test infrastructure)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gsdf_slam_b200.mapping import (DEFAULT_LR, SEGMENTS, KeyframeBatchMapper, arena_layout,
                                    partition_keyframes)
from oracle import cpu_binding as ob


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_partition_is_disjoint_and_complete():
    for n in (1, 2, 3, 4, 8):
        for K in (1, 7, 8, 16):
            parts = [partition_keyframes(K, n, r) for r in range(n)]
            flat = sorted(k for p in parts for k in p)
            assert flat == list(range(K))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_arena_layout_is_aligned_and_59_floats_per_gaussian():
    P, M = 1001, 15
    lay = arena_layout(P, M)
    total = lay["_total"][1]
    assert total >= 59 * P and total - 59 * P < 4 * len(SEGMENTS)
    end = 0
    from gsdf_slam_b200.mapping import ARENA_ORDER
    for name in ARENA_ORDER:
        off, n, shape = lay[name]
        assert off % 4 == 0 and off >= end
        assert n == int(np.prod(shape))
        end = off + n
    assert arena_layout(10, 0)["sh"][1] == 0


class _FakeMapper(KeyframeBatchMapper):
    """render_backward replaced by a deterministic function of (keyframe id, parameter); the statistics kernel
    (gsr_peer_local_stats) by the same arithmetic in torch."""

    def render_backward(self, kf, grad_views, fwd=None, on_forward=None, rendered_rows_only=False,
                        accumulate=False, accumulate_rows=None, grad_rows=None):
        for i, name in enumerate(SEGMENTS):
            g = grad_views[name]
            if g.numel():
                v = torch.sin(self.params[name] * (kf + 1) + i)
                if accumulate:       # gsr_backward_args.accumulate: later keyframes add to the arena
                    g.add_(v)
                else:
                    g.copy_(v)
        # fake per-view outputs: keyframe kf "renders" the rows r with (r + kf) % 3 != 0, radius depends on both
        rows = torch.arange(self.P)
        radii = torch.where((rows + kf) % 3 != 0, (rows * 7 + kf * 13) % 29 + 1, torch.zeros_like(rows)).to(torch.int32)
        dmean2D = torch.stack([torch.cos(rows * 0.1 + kf), torch.sin(rows * 0.2 - kf), torch.zeros(self.P)], 1).float()
        return None, radii, (dmean2D,)

    def _local_stats(self, radii, dL_dmeans2D, stats_local, flags, grad_rows=None):
        vis = radii > 0
        flags[: self.P][vis] = 1
        if stats_local is not None:
            P, S = self.P, self._stats_stride
            norm = torch.sqrt(dL_dmeans2D[:, 0] ** 2 + dL_dmeans2D[:, 1] ** 2)
            stats_local[:P][vis] += norm[vis]
            stats_local[S:S + P][vis] += 1.0
            stats_local[2 * S:2 * S + P][vis] = torch.maximum(stats_local[2 * S:2 * S + P][vis], radii[vis].float())

    def _add_stats(self, radii, dL_dmeans2D):       # single rank: the reference's per-view sequence
        if self.stats is None:
            return
        vis = radii > 0
        st = self.stats
        st["max_radii2D"][vis] = torch.maximum(st["max_radii2D"][vis], radii[vis].float())
        st["xyz_gradient_accum"][vis] += torch.sqrt(dL_dmeans2D[:, 0] ** 2 + dL_dmeans2D[:, 1] ** 2)[vis]
        st["denom"][vis] += 1.0

    def _adam_ranges(self, ranges, steps, grad_scale):
        for name, lo, hi in ranges:
            self._adam_range(name, lo, hi, steps[name], grad_scale)

    def _adam_range(self, name, lo, hi, step, grad_scale):
        p = self.opt_params[name]
        st = self.optimizer._state(p)
        pn = p.numpy().reshape(-1)[lo:hi]
        gn = (self.grad_views[name] * grad_scale).numpy().reshape(-1)[lo:hi].copy()
        ob.adam(pn, gn, st["exp_avg"].numpy().reshape(-1)[lo:hi], st["exp_avg_sq"].numpy().reshape(-1)[lo:hi],
                float(self._group(name)["lr"]), step)

    def adam_step(self, grad_scale: float = 1.0):
        for name in SEGMENTS:
            p = self.opt_params[name]
            if p.numel() == 0:
                continue
            st = self.optimizer._state(p)
            st["step"] += 1
            self._adam_range(name, 0, p.numel(), st["step"], grad_scale)


def _make_params(P=257, M=15, seed=0):
    g = torch.Generator().manual_seed(seed)
    return dict(means3D=torch.randn(P, 3, generator=g), dc=torch.randn(P, 1, 3, generator=g),
                sh=torch.randn(P, M, 3, generator=g), opacity=torch.rand(P, 1, generator=g),
                scales=torch.rand(P, 3, generator=g), rotations=torch.randn(P, 4, generator=g))


def _make_stats(P=257):
    return {"max_radii2D": torch.zeros(P), "xyz_gradient_accum": torch.zeros(P), "denom": torch.zeros(P)}


def _worker(rank, world, port, K, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _FakeMapper(_make_params(), 3, torch.zeros(3), stats=_make_stats())
        for it in range(2):
            if it == 1:
                m.set_lr("means3D", 3.2e-4)      # a learning-rate schedule takes effect on every path
            m.step(partition_keyframes(K, world, rank))
        torch.save({"arena": m.arena.clone(), "params": {k: v.clone() for k, v in m.params.items()},
                    "stats": {k: v.clone() for k, v in m.stats.items()}},
                   os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_step_equals_single_rank_sequential_accumulation(tmp_path):
    K = 5
    port = _free_port()
    mp.spawn(_worker, args=(2, port, K, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    # replicas stay bit-identical without any parameter broadcast
    assert torch.equal(r0["arena"], r1["arena"])
    for k in SEGMENTS:
        assert torch.equal(r0["params"][k], r1["params"][k])
    # the densification statistics of all K views reach every rank (sum / sum / max), identically
    for k in ("max_radii2D", "xyz_gradient_accum", "denom"):
        assert torch.equal(r0["stats"][k], r1["stats"][k]), k
    # == one rank accumulating the same K keyframes sequentially (fp32 reduction-order tolerance)
    single = _FakeMapper(_make_params(), 3, torch.zeros(3), stats=_make_stats())
    for it in range(2):
        if it == 1:
            single.set_lr("means3D", 3.2e-4)
        single.step(list(range(K)))
    assert torch.equal(single.stats["denom"], r0["stats"]["denom"])
    assert torch.equal(single.stats["max_radii2D"], r0["stats"]["max_radii2D"])
    assert torch.allclose(single.stats["xyz_gradient_accum"], r0["stats"]["xyz_gradient_accum"], rtol=1e-6, atol=1e-6)
    assert float(single.stats["denom"].max()) > K          # two steps of K views, each row seen by ~2/3 of them
    assert float(single._group("means3D")["lr"]) == 3.2e-4
    rel = float((single.arena - r0["arena"]).norm() / single.arena.norm())
    assert rel <= 1e-5
    for k in SEGMENTS:
        assert torch.allclose(single.params[k], r0["params"][k], rtol=1e-5, atol=1e-6)
