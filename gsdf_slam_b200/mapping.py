"""Keyframe-batch data-parallel mapping step.

The reference trains on ONE keyframe per iteration on one GPU
(/root/reference/src/gaussian_mapper.cpp:335-468: render -> loss -> backward -> densification
statistics -> Adam).  Here a batch of K keyframes is split disjointly across ranks (one process per
GPU); every rank renders and back-propagates its keyframes over a REPLICATED Gaussian set, then the
ranks exchange, and the semantics pinned by SURVEY.md §8e hold: batch gradient = SUM over the K
keyframes (K sequential backward() calls without zero_grad), ONE Adam step per batch, and the
per-view densification statistics (max_radii2D, xyz_gradient_accum, denom;
src/gaussian_mapper.cpp:443-447, src/gaussian_model.cpp:667-681) of all K views.

Exchange, N > 1 ranks on one NVLink node (default): ONE fused kernel over peer memory
(gsr_adam_step_peer): every rank owns an interleaved 1/N shard of the flat arena, sums the gradient
arenas of the ranks that hold a valid gradient row (per-row rank bitmasks from one flag byte per rank and row:
rendered | gradient valid; gsr_peer_rank_mask_u8), runs Adam on shard-local moments and stores the new
parameters into all N replicas (one multicast store through the NVSwitch when the symmetric allocation has a
multicast address, else N unicast stores).  The densification statistics ride through the same kernel as three
one-float segments in the tail of the arenas (sum / sum / max over the ranks that rendered the row).
Fallback (no symmetric memory, or GSR_PEER_EXCHANGE=0): NCCL all-reduce of the arena (chunked, pipelined
with the fused Adam) and all-reduce sum / max of the statistics; every rank then applies the identical
full-size update.  Single rank: the K backward passes write / add only the rows that received a gradient
(gsr_backward_args.grad_rows, a row mask the kernel keeps) and ONE masked Adam pass reads the gradient only for
those rows.

Gradient arena layout (fp32, each segment 16-byte aligned and itself a valid contiguous gradient
tensor): [xyz 3P | f_dc 3P | opacity P | scaling 3P | rotation 4P | f_rest 3MP].

A keyframe either carries dL/d(image) (the caller owns the loss) or a ground-truth image, in which
case the fused L1 + SSIM loss of loss.py (row f-1) runs between the two passes.  With ``raw_params`` the
model's RAW opacity / scaling / rotation are the inputs and the activations run inside the preprocess
kernels (row f-2).
"""
from __future__ import annotations

import ctypes as C
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
STATS = ("max_radii2D", "xyz_gradient_accum", "denom")


def partition_keyframes(num_keyframes: int, world_size: int, rank: int) -> List[int]:
    """Rank r owns keyframes {r, r+n, r+2n, ...}."""
    return list(range(rank, num_keyframes, world_size))


# order of the segments inside the arena: the five small ones first (one 14*P-float chunk),
# the 3*M*P-float SH block last so that it can be cut into a few large chunks
ARENA_ORDER = ("means3D", "dc", "opacity", "scales", "rotations", "sh")


def arena_layout(P: int, M: int) -> Dict[str, tuple]:
    """name -> (offset, numel, shape); offsets are multiples of 4 floats (16 bytes)."""
    shapes = {"means3D": (P, 3), "dc": (P, 1, 3), "sh": (P, M, 3), "opacity": (P, 1), "scales": (P, 3),
              "rotations": (P, 4)}
    out, off = {}, 0
    for name in ARENA_ORDER:
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
    dL_dout: Optional[torch.Tensor] = None   # [3,H,W] dL/d(rendered image), if the caller owns the loss
    gt_image: Optional[torch.Tensor] = None  # [3,H,W] ground truth: dL_dout then comes from the fused
    mask: Optional[torch.Tensor] = None      # [H,W]   L1 + SSIM loss (loss.py)


class KeyframeBatchMapper:
    """Replicated Gaussians + flat gradient arena + fused Adam; one instance per rank.

    ``stats``: optional dict {"max_radii2D", "xyz_gradient_accum", "denom"} of float32 tensors with P elements
    (the replicated densification statistics, e.g. GaussianModel's); updated with the statistics of all K
    keyframes of every step, identically on every rank.
    ``single_rank``: ignore the process group (sequential reference of the multi-rank step in tests / bench).
    """

    def __init__(self, params: Dict[str, torch.Tensor], degree: int, bg: torch.Tensor,
                 lr: Optional[Dict[str, float]] = None, eps: float = 1e-15,
                 process_group=None, freeze_render_inputs: bool = False, overlap_adam: bool = False,
                 masked_single_view: bool = True, peer_exchange: Optional[bool] = None,
                 lambda_dssim: float = 0.2, raw_params: bool = False,
                 stats: Optional[Dict[str, torch.Tensor]] = None, single_rank: bool = False,
                 multicast: Optional[bool] = None):
        self.params = {k: params[k].contiguous() for k in SEGMENTS}
        self.degree = int(degree)
        self.bg = bg
        self.pg = process_group
        self.single_rank = bool(single_rank)
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
        self._stats_stride = (max(self.P, 1) + 3) // 4 * 4     # floats between the three per-rank delta arrays
        self.stats = None
        self._stats_user = stats     # with the peer exchange the statistics move into the arena: both dicts get the views
        if stats is not None:
            self.stats = {}
            for k in STATS:
                t = stats[k]
                if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != self.P:
                    raise RuntimeError(f"stats[{k!r}] must be a contiguous float32 tensor with P elements")
                self.stats[k] = t
        lr = dict(DEFAULT_LR, **(lr or {}))
        self.last_radii: Optional[torch.Tensor] = None
        self.last_color: Optional[torch.Tensor] = None
        self.lambda_dssim = float(lambda_dssim)
        # raw_params: `params` holds the model's RAW opacity / scaling / rotation;
        # the activations and their backward run inside the preprocess kernels, so the gradients in
        # the arena are those of the leaf parameters and Adam trains the model itself
        self.raw_params = bool(raw_params)
        self.last_loss: Optional[torch.Tensor] = None   # device [3] {loss, L1, SSIM} of the last keyframe
        # N > 1 ranks on one NVLink node: gradient reduction + Adam + parameter broadcast fused in
        # one kernel over peer memory (gsr_adam_step_peer) instead of NCCL all-reduce + full Adam
        self.peer = None
        self.traffic_counters = None   # set by measure_exchange_traffic only
        self._grad_rows: Optional[torch.Tensor] = None
        if peer_exchange is None:
            import os
            peer_exchange = os.environ.get("GSR_PEER_EXCHANGE", "1") != "0"
        if multicast is None:
            import os
            multicast = os.environ.get("GSR_PEER_MULTICAST", "1") != "0"
        if peer_exchange and dev.type == "cuda" and self.world_size > 1:
            self._setup_peer_exchange(eps, bool(multicast))
        # the optimiser is built over the FINAL parameter tensors (the symmetric views on the peer path):
        # learning rates / betas / eps are read from its param_groups on every step by every path
        self.optimizer = FusedAdam([{"params": [self.opt_params[k]], "lr": lr[k], "name": k} for k in SEGMENTS
                                    if self.opt_params[k].numel() > 0], lr=0.0, eps=eps)
        # statistics deltas of this rank's keyframes for the NCCL fallback (the peer path keeps them in
        # symmetric memory)
        self._stats_local: Optional[torch.Tensor] = None
        # visibility-split Adam (gsr_adam_step_rows) overlapped with the rasterizer on a side stream.  OFF by
        # default: measured on B200 (cfg2) the streaming sweep slows the co-running blend kernels by as much
        # as it hides
        self.overlap_adam = bool(overlap_adam) and dev.type == "cuda"
        # single rank: the backward passes write only the rendered rows (no dense zero fill) and ONE masked
        # Adam pass reads the gradient only for those rows (gsr_adam_step_rows, GSR_ADAM_ROWS_ALL) — 276 +
        # 59*4 bytes per non-rendered Gaussian less HBM traffic per iteration than zero-fill + dense Adam,
        # bit-identical result
        self.masked_single_view = bool(masked_single_view) and dev.type == "cuda"
        self.serialize_side_stream = False   # profiling aid: run the side-stream work on the main stream
        self.trace = None                    # profiling aid: list collecting (label, timing event) pairs
        self._side: Optional[torch.cuda.Stream] = None
        self._ev_radii: Optional[torch.cuda.Event] = None
        self._ev_side: Optional[torch.cuda.Event] = None

    # ---- small helpers ---------------------------------------------------------------------------
    def _views(self, arena: torch.Tensor) -> Dict[str, torch.Tensor]:
        v = {}
        for name in SEGMENTS:
            off, n, shape = self.layout[name]
            v[name] = arena[off:off + n].view(shape)
        return v

    @property
    def world_size(self) -> int:
        if self.single_rank:
            return 1
        import torch.distributed as dist
        return dist.get_world_size(self.pg) if (dist.is_available() and dist.is_initialized()) else 1

    def _group(self, name: str) -> dict:
        return next(g for g in self.optimizer.param_groups if g.get("name") == name)

    def set_lr(self, name: str, lr: float) -> None:
        """The reference's updateLearningRate (src/gaussian_model.cpp:354-396): takes effect on every path."""
        self._group(name)["lr"] = float(lr)

    def _stream_ptr(self):
        return C.c_void_p(torch.cuda.current_stream(self.arena.device).cuda_stream)

    def _mark(self, label: str, stream) -> None:
        if self.trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            self.trace.append((label, ev))

    # ---- peer-memory exchange (N > 1) ------------------------------------------------------------
    def _setup_peer_exchange(self, eps: float, multicast: bool) -> None:
        """Moves the gradient arena and the optimised parameters into symmetric (peer-mapped)
        memory and shards the Adam state.  Falls back to the NCCL path (with a warning) when
        symmetric memory cannot be set up on this system."""
        import sys
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.pg if self.pg is not None else dist.group.WORLD
            dev = self.arena.device
            total = self.layout["_total"][1]
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            if world > 8:
                raise RuntimeError("more than GSR_MAX_PEERS ranks")
            # with densification statistics the arenas grow by three one-float segments [accum | denom | max_radii]:
            # gradient arena = this rank's per-step deltas, parameter arena = the replicated statistics, so the
            # fused kernel reduces and broadcasts them with everything else (GSR_PEER_RULE_SUM / _MAX)
            stat_off = (total + 3) // 4 * 4
            arena_len = stat_off + 3 * self._stats_stride if self.stats is not None else total
            g_arena = symm.empty(arena_len, dtype=torch.float32, device=dev)
            p_arena = symm.empty(arena_len, dtype=torch.float32, device=dev)
            flags = symm.empty(max(self.P, 1) + 16, dtype=torch.uint8, device=dev)   # per row: rendered | gradient valid
            g_arena.zero_(); p_arena.zero_(); flags.zero_()
            hg = symm.rendezvous(g_arena, group)
            hp = symm.rendezvous(p_arena, group)
            hf = symm.rendezvous(flags, group)
            st_local = g_arena[stat_off:stat_off + 3 * self._stats_stride] if self.stats is not None else None
        except Exception as exc:  # noqa: BLE001
            print(f"[gsdf_slam_b200] peer-memory exchange unavailable ({exc!r}); using the NCCL all-reduce path",
                  file=sys.stderr)
            return
        # parameters the optimiser updates now live in the symmetric parameter arena
        p_views = self._views(p_arena)
        frozen = self.opt_params is not self.params
        for k in SEGMENTS:
            p_views[k].copy_(self.opt_params[k])
        self.opt_params = p_views
        if not frozen:
            self.params = p_views
        self.arena = g_arena
        self.grad_views = self._views(g_arena)
        if self.stats is not None:
            # the caller's statistics move into the parameter arena; the dict it handed over now holds the views
            order = ("xyz_gradient_accum", "denom", "max_radii2D")
            for j, name in enumerate(order):
                view = p_arena[stat_off + j * self._stats_stride: stat_off + j * self._stats_stride + self.P]
                view.copy_(self.stats[name])
                self.stats[name] = view
                self._stats_user[name] = view
            total = arena_len
        # interleaved sharding: rank r owns every world-th 64 Ki-element chunk of the arena, so every rank
        # gets the same mix of narrow (xyz, opacity, ...) and wide (sh) rows — a contiguous 1/N range gave
        # the ranks holding the narrow segments 30 % more parameter stores (measured at N = 8)
        chunk = 65536
        n_chunks = (total + chunk - 1) // chunk
        owned = (n_chunks - rank + world - 1) // world if n_chunks > rank else 0
        n_mom = max(owned * chunk, 4)
        mc_ptr = 0
        if multicast:
            try:
                mc_ptr = int(getattr(hp, "multicast_ptr", 0) or 0)
            except Exception:  # noqa: BLE001
                mc_ptr = 0
        self.peer = dict(world=world, rank=rank, hg=hg, hp=hp, hf=hf, g_arena=g_arena, p_arena=p_arena,
                         flags=flags, stats_local=st_local, stat_off=stat_off, lo=0, hi=total, step=0, eps=float(eps),
                         chunk=chunk,
                         mc_ptr=mc_ptr,
                         exp_avg=torch.zeros(n_mom, dtype=torch.float32, device=dev),
                         exp_avg_sq=torch.zeros(n_mom, dtype=torch.float32, device=dev),
                         union=torch.zeros(max(self.P, 1) + 4, dtype=torch.int32, device=dev))
        torch.cuda.synchronize(dev)
        dist.barrier(group)

    def measure_exchange_traffic(self, step_fn) -> Optional[dict]:
        """Runs `step_fn` (one optimiser step) with the counting variant of the peer kernel and returns this rank's
        NVLink traffic of that step in bytes: gradient rows pulled from other ranks and parameter groups pushed."""
        if self.peer is None:
            return None
        self.traffic_counters = torch.zeros(2, dtype=torch.int64, device=self.arena.device)
        try:
            step_fn()
            torch.cuda.synchronize(self.arena.device)
            loads, changed = (int(x) for x in self.traffic_counters.tolist())
        finally:
            self.traffic_counters = None
        n = self.peer["world"]
        return {"grad_in_bytes": 16 * loads, "param_changed_bytes": 16 * changed,
                "param_out_bytes": 16 * changed * (1 if self.peer["mc_ptr"] else n - 1)}

    def _grad_rows_mask(self) -> torch.Tensor:
        """The batch's row mask (int32 [P], zeroed): gsr_backward marks the rows that received a gradient."""
        if self._grad_rows is None:
            self._grad_rows = torch.zeros(self.P, dtype=torch.int32, device=self.arena.device)
        else:
            self._grad_rows.zero_()
        return self._grad_rows

    def exchange_bytes(self) -> Optional[dict]:
        """Model of the NVLink traffic of one peer exchange, per rank (bench.py prints it next to the measured
        time): parameter broadcast = this rank's shard to N-1 peers (one multicast store if available)."""
        if self.peer is None:
            return None
        pe = self.peer
        total = self.layout["_total"][1]
        shard = 4 * total / pe["world"]
        return {"param_shard_bytes": int(shard), "param_out_bytes_unicast": int(shard * (pe["world"] - 1)),
                "param_in_bytes": int(shard * (pe["world"] - 1)), "multicast": bool(pe["mc_ptr"])}

    def _local_stats(self, radii: torch.Tensor, dL_dmeans2D: torch.Tensor, stats_local: Optional[torch.Tensor],
                     flags: torch.Tensor, grad_rows: Optional[torch.Tensor] = None) -> None:
        from . import _lib
        r = radii if radii.dtype == torch.int32 else radii.to(torch.int32)
        _lib.check(_lib.lib().gsr_peer_local_stats(self.P, C.c_void_p(r.data_ptr()), C.c_void_p(dL_dmeans2D.data_ptr()),
                                                   C.c_void_p(stats_local.data_ptr()) if stats_local is not None else None,
                                                   self._stats_stride, C.c_void_p(flags.data_ptr()),
                                                   C.c_void_p(grad_rows.data_ptr()) if grad_rows is not None else None,
                                                   self._stream_ptr()),
                   "gsr_peer_local_stats")

    def _step_peer(self, keyframes: Sequence[Keyframe], grad_scale: float, on_forward=None) -> None:
        from . import _lib
        pe = self.peer
        dev = self.arena.device
        L = _lib.lib()
        main = torch.cuda.current_stream(dev)
        self._mark("step_begin", main)
        # this rank's "rendered" flags (and statistics deltas) of the step; the previous step's second barrier
        # guarantees that no peer still reads them
        pe["flags"].zero_()
        if pe["stats_local"] is not None:
            pe["stats_local"].zero_()
        rows = self._grad_rows_mask()
        for i, kf in enumerate(keyframes):
            # rows without a valid gradient on this rank are never read by the exchange (per-row rank bitmask), so
            # no backward pass needs the dense zero fill: the first keyframe writes the rows that received a
            # gradient, the others add to the rows already written and overwrite the rest; rendered rows that no
            # pixel blended are not touched at all (gsr_backward_args.grad_rows, kept by the kernel)
            _, radii, g = self.render_backward(kf, self.grad_views, on_forward=on_forward,
                                               rendered_rows_only=True, accumulate=(i > 0), grad_rows=rows)
            self._local_stats(radii, g[0], pe["stats_local"], pe["flags"], rows)
        stream = self._stream_ptr()
        self._mark("render_end", main)
        pe["hg"].barrier(channel=0)          # every rank's gradients, flags and statistics are written
        self._mark("barrier1_end", main)
        fptrs = (C.c_void_p * pe["world"])(*[int(x) for x in pe["hf"].buffer_ptrs])
        _lib.check(L.gsr_peer_rank_mask_u8(fptrs, pe["world"], self.P, C.c_void_p(pe["union"].data_ptr()), stream),
                   "gsr_peer_rank_mask_u8")
        self._mark("mask_end", main)
        pe["step"] += 1
        a = _lib.PeerAdamArgs()
        a.n_peers, a.rank = pe["world"], pe["rank"]
        for k in range(pe["world"]):
            a.param_peers[k] = int(pe["hp"].buffer_ptrs[k])
            a.grad_peers[k] = int(pe["hg"].buffer_ptrs[k])
        segs = [n for n in ARENA_ORDER if self.layout[n][1] > 0]
        a.n_segments = len(segs)
        betas = eps = None
        for i, n in enumerate(segs):
            off, numel, _shape = self.layout[n]
            grp = self._group(n)
            if betas is None:
                betas, eps = grp["betas"], grp["eps"]
            elif (betas, eps) != (grp["betas"], grp["eps"]):
                raise RuntimeError("gsr_adam_step_peer needs the same betas / eps in every parameter group")
            a.segments[i].start, a.segments[i].numel = off, numel
            a.segments[i].width = numel // self.P
            a.segments[i].lr = float(grp["lr"])
        if pe["stats_local"] is not None:   # accum, denom: summed over the ranks; max_radii: their maximum
            for j, rule in enumerate((_lib.GSR_PEER_RULE_SUM, _lib.GSR_PEER_RULE_SUM, _lib.GSR_PEER_RULE_MAX)):
                sg = a.segments[len(segs) + j]
                sg.start, sg.numel = pe["stat_off"] + j * self._stats_stride, self.P
                sg.width, sg.rule, sg.lr = 1, rule, 0.0
            a.n_segments = len(segs) + 3
        a.shard_lo, a.shard_hi, a.shard_chunk = pe["lo"], pe["hi"], pe["chunk"]
        a.exp_avg, a.exp_avg_sq = pe["exp_avg"].data_ptr(), pe["exp_avg_sq"].data_ptr()
        a.union_mask = pe["union"].data_ptr()
        a.beta1, a.beta2, a.eps, a.grad_scale = float(betas[0]), float(betas[1]), float(eps), float(grad_scale)
        a.step = pe["step"]
        a.stream = stream
        a.param_multicast = pe["mc_ptr"]
        if self.traffic_counters is not None:   # measurement only (bench.py): counted variant of the kernel
            a.traffic_counters = self.traffic_counters.data_ptr()
        _lib.check(L.gsr_adam_step_peer(C.byref(a)), "gsr_adam_step_peer")
        self._mark("peer_adam_end", main)
        pe["hp"].barrier(channel=0)          # every rank's parameter stores have landed
        self._mark("step_end", main)

    # ---- rendering -------------------------------------------------------------------------------
    def render(self, kf: Keyframe, radii_ready_event=None):
        """Forward pass of one keyframe -> the reference's 8-tuple (rasterize_points.h:18-39)."""
        p = self.params
        e = torch.empty(0, dtype=torch.float32, device=p["means3D"].device)
        return RasterizeGaussiansCUDA(self.bg, p["means3D"], e, p["opacity"], p["scales"], p["rotations"],
                                      1.0, e, kf.viewmatrix, kf.projmatrix, kf.tanfovx, kf.tanfovy, kf.H,
                                      kf.W, p["dc"], p["sh"], self.degree, kf.campos, False, False,
                                      radii_ready_event=radii_ready_event, raw_params=self.raw_params)

    def render_backward(self, kf: Keyframe, grad_views: Dict[str, torch.Tensor], fwd=None, on_forward=None,
                        rendered_rows_only: bool = False, accumulate: bool = False,
                        accumulate_rows: Optional[torch.Tensor] = None, grad_rows: Optional[torch.Tensor] = None):
        """fwd + bwd of one keyframe; parameter gradients are written into grad_views.
        ``on_forward(kf, fwd)`` runs between the two passes (e.g. to start reading the image back
        and to make the stream wait for the uploaded dL/d(image))."""
        p = self.params
        e = torch.empty(0, dtype=torch.float32, device=p["means3D"].device)
        if fwd is None:
            fwd = self.render(kf)
            if on_forward is not None:
                on_forward(kf, fwd)
        R, B, color, radii, geom, binning, img, sample = fwd
        dL_dout = kf.dL_dout
        if dL_dout is None:
            # photometric loss of the reference's mapping iteration (src/gaussian_mapper.cpp:414-431),
            # forward + backward in two kernels
            from .loss import l1_ssim_loss
            if kf.gt_image is None:
                raise RuntimeError("Keyframe needs either dL_dout or gt_image")
            self.last_loss, dL_dout = l1_ssim_loss(color, kf.gt_image, kf.mask, self.lambda_dssim)
        out = {GRAD_NAME[k]: grad_views[k] for k in SEGMENTS if grad_views[k].numel() > 0}
        g = RasterizeGaussiansBackwardCUDA(self.bg, p["means3D"], radii, e, p["scales"], p["rotations"], 1.0,
                                           e, kf.viewmatrix, kf.projmatrix, kf.tanfovx, kf.tanfovy,
                                           dL_dout, p["dc"], p["sh"], self.degree, kf.campos, geom, R,
                                           binning, img, B, sample, False, want_dcov3D=False, out=out,
                                           rendered_rows_only=rendered_rows_only, accumulate=accumulate,
                                           accumulate_rows=accumulate_rows, grad_rows=grad_rows,
                                           raw_opacity=p["opacity"] if self.raw_params else None)
        self.last_radii, self.last_color = radii, color
        return color, radii, g

    def _add_stats(self, radii: torch.Tensor, dL_dmeans2D: torch.Tensor) -> None:
        """One view's densification statistics straight into the (single-rank) state: the reference's sequence
        (src/gaussian_mapper.cpp:443-447 + GaussianModel::addDensificationStats)."""
        if self.stats is None or self.P == 0:
            return
        from . import _lib
        st = self.stats
        r = radii if radii.dtype == torch.int32 else radii.to(torch.int32)
        _lib.check(_lib.lib().gsr_add_densification_stats(self.P, C.c_void_p(r.data_ptr()),
                                                          C.c_void_p(dL_dmeans2D.data_ptr()),
                                                          C.c_void_p(st["max_radii2D"].data_ptr()),
                                                          C.c_void_p(st["xyz_gradient_accum"].data_ptr()),
                                                          C.c_void_p(st["denom"].data_ptr()), self._stream_ptr()),
                   "gsr_add_densification_stats")

    def accumulate(self, keyframes: Sequence[Keyframe], on_forward=None) -> None:
        """Sum of the local keyframes' gradients -> arena, DENSE (rows nobody rendered hold zeros): what an
        all-reduce needs.  == K sequential backward() calls without zero_grad."""
        multi = self.world_size > 1
        if self.stats is not None and multi:
            if self._stats_local is None:
                self._stats_local = torch.zeros(3 * self._stats_stride, dtype=torch.float32, device=self.arena.device)
                self._flags_local = torch.zeros(max(self.P, 1), dtype=torch.uint8, device=self.arena.device)
            self._stats_local.zero_()
        if len(keyframes) == 0:
            self.arena.zero_()
            return
        for i, kf in enumerate(keyframes):
            # the first keyframe writes in place (dense: zero fill + rendered rows); later keyframes add their
            # rendered rows straight into the arena (gsr_backward accumulate): no scratch arena, no dense add
            _, radii, g = self.render_backward(kf, self.grad_views, on_forward=on_forward, accumulate=(i > 0))
            if self.stats is not None:
                if multi:
                    self._local_stats(radii, g[0], self._stats_local, self._flags_local)
                else:
                    self._add_stats(radii, g[0])

    def allreduce(self) -> None:
        import torch.distributed as dist
        if self.world_size > 1:
            dist.all_reduce(self.arena, op=dist.ReduceOp.SUM, group=self.pg)

    def _allreduce_stats(self) -> None:
        """NCCL fallback of gsr_peer_stats_exchange: sum {accum, denom}, max {max_radii} over the ranks."""
        import torch.distributed as dist
        if self.stats is None or self.world_size <= 1 or self._stats_local is None:
            return
        P, S = self.P, self._stats_stride
        sl = self._stats_local
        dist.all_reduce(sl[: 2 * S], op=dist.ReduceOp.SUM, group=self.pg)
        dist.all_reduce(sl[2 * S: 3 * S], op=dist.ReduceOp.MAX, group=self.pg)
        st = self.stats
        st["xyz_gradient_accum"].view(-1).add_(sl[:P])
        st["denom"].view(-1).add_(sl[S: S + P])
        torch.maximum(st["max_radii2D"].view(-1), sl[2 * S: 2 * S + P], out=st["max_radii2D"].view(-1))

    def adam_step(self, grad_scale: float = 1.0) -> None:
        grads = {self.opt_params[k]: self.grad_views[k] for k in SEGMENTS
                 if self.opt_params[k].numel() > 0}
        self.optimizer.step(grads=grads, grad_scale=grad_scale)

    # ---- pipelined exchange: all-reduce chunk k+1 while Adam consumes chunk k -------------
    def _chunks(self, target_elems: int = 8 * 1024 * 1024):
        """[(arena_lo, arena_hi, [(segment, lo, hi), ...])]: contiguous arena ranges of roughly
        32 MB — the five small segments together, the SH block in a few pieces."""
        small = [n for n in ARENA_ORDER if n != "sh" and self.layout[n][1] > 0]
        out = []
        if small:
            lo = self.layout[small[0]][0]
            hi = self.layout[small[-1]][0] + self.layout[small[-1]][1]
            out.append((lo, hi, [(n, 0, self.layout[n][1]) for n in small]))
        off, n, _ = self.layout["sh"]
        if n > 0:
            k = max(1, (n + target_elems - 1) // target_elems)
            step = ((n + k - 1) // k + 3) // 4 * 4
            for lo in range(0, n, step):
                hi = min(n, lo + step)
                out.append((off + lo, off + hi, [("sh", lo, hi)]))
        return out

    def _adam_ranges(self, ranges, steps, grad_scale: float) -> None:
        """One fused Adam launch over sub-ranges [(segment, lo, hi)] of the parameter tensors."""
        from . import _lib
        items = []
        betas = eps = None
        for name, lo, hi in ranges:
            p = self.opt_params[name].view(-1)
            st = self.optimizer._state(self.opt_params[name])
            g = self.grad_views[name].view(-1)
            grp = self._group(name)
            betas, eps = grp["betas"], grp["eps"]
            items.append(_lib.AdamGroup(param=p[lo:hi].data_ptr(), grad=g[lo:hi].data_ptr(),
                                        exp_avg=st["exp_avg"].view(-1)[lo:hi].data_ptr(),
                                        exp_avg_sq=st["exp_avg_sq"].view(-1)[lo:hi].data_ptr(),
                                        numel=hi - lo, lr=float(grp["lr"]), step=int(steps[name])))
        arr = (_lib.AdamGroup * len(items))(*items)
        _lib.check(_lib.lib().gsr_adam_step(arr, len(items), float(betas[0]), float(betas[1]), float(eps),
                                            float(grad_scale), self._stream_ptr()),
                   "gsr_adam_step")

    def exchange_and_update(self, grad_scale: float = 1.0) -> None:
        """Gradient exchange + optimiser.  With more than one rank the arena is all-reduced in
        ~32 MB chunks on NCCL's stream and the fused Adam consumes each chunk as soon as it has
        landed, so most of the Adam pass (HBM-bound) hides behind the NVLink transfer."""
        import torch.distributed as dist
        if self.world_size <= 1:
            self.adam_step(grad_scale)
            return
        chunks = self._chunks()
        works = []
        for lo, hi, _ in chunks:
            works.append(dist.all_reduce(self.arena[lo:hi], op=dist.ReduceOp.SUM, group=self.pg,
                                         async_op=True))
        steps = {}
        for name in SEGMENTS:
            if self.opt_params[name].numel():
                st = self.optimizer._state(self.opt_params[name])
                st["step"] += 1
                steps[name] = st["step"]
        for (_, _, ranges), w in zip(chunks, works):
            w.wait()  # the compute stream waits for this chunk only
            self._adam_ranges(ranges, steps, grad_scale)
        self._allreduce_stats()

    # ---- single-rank step with the visibility-split Adam overlapped with the rasterizer -----
    def _streams(self):
        dev = self.arena.device
        if self._side is None:
            import os
            # the side kernel is a small persistent grid: give it priority so that its few CTAs are
            # placed at once and then stream under the rasterizer kernels
            prio = -1 if os.environ.get("GSR_SIDE_PRIO", "high") == "high" else 0
            self._side = torch.cuda.Stream(device=dev, priority=prio)
            self._ev_radii = torch.cuda.Event()
            self._ev_side = torch.cuda.Event()
        return torch.cuda.current_stream(dev), self._side

    def _step_overlapped(self, keyframes: Sequence[Keyframe], grad_scale: float, on_forward=None) -> None:
        """Rows that no local keyframe renders have a zero gradient: their Adam update only needs
        `radii`, so it is enqueued on a side stream right after the (last) forward pass and runs
        under the blend / backward kernels; the rendered rows follow the backward pass on the
        main stream.  Bit-identical to accumulate() + adam_step() (tests/test_gpu_kernels.py)."""
        main, side = self._streams()
        if self.serialize_side_stream:
            side = main
        opt = self.optimizer
        grads = {self.opt_params[k]: self.grad_views[k] for k in SEGMENTS if self.opt_params[k].numel() > 0}
        K = len(keyframes)
        mask = None
        self._mark("step_begin", main)
        for i, kf in enumerate(keyframes):
            last = i == K - 1
            fwd = self.render(kf, radii_ready_event=self._ev_radii if (last and K == 1) else None)
            if on_forward is not None:
                on_forward(kf, fwd)
            radii = fwd[3]
            mask = radii if mask is None else torch.maximum(mask, radii)
            if last:
                if K > 1:
                    self._ev_radii.record(main)
                side.wait_event(self._ev_radii)
                mask.record_stream(side)
                self._mark("forward_end", main)
                with torch.cuda.stream(side):
                    self._mark("side_adam_begin", side)
                    opt.step_rows(mask, opt.ROWS_ZERO_GRAD, grads=None, grad_scale=grad_scale, new_step=True)
                    self._mark("side_adam_end", side)
                    self._ev_side.record(side)
            if i == 0:
                _, _, g = self.render_backward(kf, self.grad_views, fwd=fwd)
            else:
                if self._scratch_arena is None:
                    self._scratch_arena = torch.zeros_like(self.arena)
                _, _, g = self.render_backward(kf, self._views(self._scratch_arena), fwd=fwd)
                self.arena.add_(self._scratch_arena)
            self._add_stats(radii, g[0])
        self._mark("backward_end", main)
        opt.step_rows(mask, opt.ROWS_WITH_GRAD, grads=grads, grad_scale=grad_scale, new_step=False)
        self._mark("main_adam_end", main)
        main.wait_event(self._ev_side)  # later work on this stream sees the complete update

    def _step_masked(self, keyframes: Sequence[Keyframe], grad_scale: float, on_forward=None) -> None:
        """Single rank, K >= 1 keyframes: no backward pass zero-fills; the first writes the rows that received a
        gradient, the others add to the rows already written and overwrite the rest (gsr_backward_args.grad_rows);
        ONE masked Adam pass reads the gradient only for those rows.  Same result as zero fill + dense sum + dense
        Adam (rows without gradient contribute exact zeros there)."""
        mask = self._grad_rows_mask()
        for i, kf in enumerate(keyframes):
            _, radii, g = self.render_backward(kf, self.grad_views, on_forward=on_forward, rendered_rows_only=True,
                                               accumulate=(i > 0), grad_rows=mask)
            self._add_stats(radii, g[0])
        grads = {self.opt_params[k]: self.grad_views[k] for k in SEGMENTS if self.opt_params[k].numel() > 0}
        self.optimizer.step_rows(mask, self.optimizer.ROWS_ALL, grads=grads, grad_scale=grad_scale, new_step=True)

    def step(self, keyframes: Sequence[Keyframe], grad_scale: float = 1.0, on_forward=None) -> None:
        """One mapping iteration over this rank's keyframes."""
        if self.peer is not None and self.P > 0:
            self._step_peer(keyframes, grad_scale, on_forward)
            return
        if self.overlap_adam and self.world_size == 1 and len(keyframes) > 0 and self.P > 0:
            self._step_overlapped(keyframes, grad_scale, on_forward)
            return
        if self.masked_single_view and self.world_size == 1 and len(keyframes) >= 1 and self.P > 0:
            self._step_masked(keyframes, grad_scale, on_forward)
            return
        self.accumulate(keyframes, on_forward)
        self.exchange_and_update(grad_scale)
