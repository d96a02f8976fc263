#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native rasterizer hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config cfg2]

Headline workload (BASELINE.json configs[1], "cfg2"): Replica-style keyframe, 500 k SH-3 Gaussians at
1200x680, seeded synthetic scene (gsdf_slam_b200/scenes.py).  One STEP = one mapping iteration of the hot
path on every rank: RasterizeGaussiansCUDA (forward) + RasterizeGaussiansBackwardCUDA (backward) of that
rank's keyframe over the replicated Gaussian set, the gradient exchange when N > 1 (fused peer-memory
kernel) and the fused Adam step.  Weak scaling: one keyframe per rank per step, value = keyframe
iterations / s over all ranks.  dL/d(image) is a seeded tensor (--with-loss puts the L1+SSIM loss inside).

`value`  : inputs resident in HBM, K steps timed with CUDA events between barriers.
`e2e`    : the same step through the public API with HOST buffers: every step copies the keyframe inputs
           (camera block, dL/d(image)) from pinned host memory and reads the rendered image back into pinned
           host memory; uploads and read-backs run on two copy streams, double-buffered — the image of step k is
           waited for before step k+2 is enqueued, so one rank's PCIe stall is not broadcast through the barriers.
Further blocks of the same JSON line:
`rasterizer_only` (N = 1): the metric's literal quantity — RasterizeGaussiansCUDA + RasterizeGaussiansBackwardCUDA
           through the C++ drop-in shim (dense gradients, the reference's allocation pattern), it/s.
`mapping_cfg5` : BASELINE.json configs[4]: 3 M Gaussians, K = 8 keyframes per optimiser step split over the N
           ranks (STRONG scaling of a fixed batch): fwd + bwd of the local keyframes, densification statistics,
           exchange, Adam; steps/s, per-rank breakdown, NVLink bytes of the exchange.
`exchange_parity` (N > 1, untimed): the same steps three ways — fused peer exchange, NCCL all-reduce path, one
           rank accumulating all keyframes sequentially — and the differences between them.
--impl reference : the UNMODIFIED reference CUDA rasterizer compiled for sm_100a (oracle/_ref, see
           oracle/build_ref.sh) + LibTorch's Adam op sequence, same step(s), rank 0 only.  (The reference has no
           CPU implementation; its CPU restatement — the oracle port — is what `cpu_baseline` times.)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rasterizer fwd+bwd it/s @1200x680, 500k Gaussians; mapping it/s at 1/2/4/8 GPU"
UNIT = "keyframe it/s (fwd+bwd+Adam)"
CFG5_KEYFRAMES = 8


def env_int(name, default):
    return int(os.environ.get(name, default))


# --------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            inside = (t0 - 0.05) <= ts <= (t1 + 0.15)
            try:
                if inside:
                    sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            if inside:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"), f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: use every sample we have
            sm = [float(l.split(",")[1]) for _, l in self.lines if len(l.split(",")) >= 8] or [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


NVLINK_PEER_GBS = 770.0   # measured peer-copy bandwidth per direction per GPU on this pool (B200_PROFILING.md)


def make_scene(cfg: str, rank: int):
    from gsdf_slam_b200 import scenes
    if cfg == "cfg2":
        scene = scenes.cfg2()
        # every rank renders its own keyframe: cameras on a small ring, looking outward
        cams = scenes.ring_cameras(8, scene.cam.W, scene.cam.H, 600.0, 600.0, radius=0.3)
        cam = scene.cam if rank == 0 else cams[rank % 8]
    elif cfg == "cfg1":
        scene = scenes.cfg1(sh_coeffs=15)
        cam = scene.cam
    elif cfg == "cfg4":
        scene = scenes.cfg4()
        cam = scene.cam
    elif cfg == "small":  # CI-sized
        scene = scenes.room_scene(60_000, 640, 360, 320.0, 320.0, seed=2, name="small_room")
        cam = scene.cam
    else:
        raise SystemExit(f"unknown --config {cfg}")
    return scene, cam


def make_cfg5(small: bool):
    from gsdf_slam_b200 import scenes
    if small:   # CI-sized stand-in (--config small / --small-cfg5)
        s = scenes.room_scene(120_000, 640, 360, 320.0, 320.0, seed=5, box=(16.0, 3.0, 12.0), name="cfg5_small")
        s.extra_cams = scenes.ring_cameras(CFG5_KEYFRAMES, 640, 360, 320.0, 320.0, 1.0)
        return s
    return scenes.cfg5()


def host_keyframe(scene, cam, with_loss=False):
    """Pinned host copies of one step's inputs.  The image-sized input ("dL_dout") is the seeded
    dL/d(image), or — with --with-loss — a seeded ground-truth image for the L1+SSIM loss."""
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    big = scene.dL_dout
    if with_loss:
        big = np.random.default_rng(77).uniform(0.0, 1.0, scene.dL_dout.shape).astype(np.float32)
    return dict(viewmatrix=pin(cam.viewmatrix), projmatrix=pin(cam.projmatrix), campos=pin(cam.campos),
                bg=pin(scene.bg), dL_dout=pin(big))


CAMERA_KEYS = ("viewmatrix", "projmatrix", "campos", "bg")


def camera_block(hk, dev):
    """The per-step camera inputs (view 16 + proj 16 + campos 3 + background 3 floats) as ONE pinned host
    block and ONE device block with named views: a single 152-byte copy per step instead of four."""
    flat = torch.cat([hk[k].reshape(-1) for k in CAMERA_KEYS]).pin_memory()
    d = torch.empty_like(flat, device=dev)
    views, off = {}, 0
    for k in CAMERA_KEYS:
        n = hk[k].numel()
        views[k] = d[off:off + n].view(hk[k].shape)
        off += n
    return flat, d, views


def dev_params(scene, dev):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    return dict(means3D=t(scene.means3D), dc=t(scene.dc), sh=t(scene.sh), opacity=t(scene.opacity),
                scales=t(scene.scales), rotations=t(scene.rotations))


def workload_counts(scene, cam, params, dev):
    """P, V, R, E, pair counts ... of this rank's keyframe (computed once, outside the timed region)."""
    import gsdf_slam_b200 as gs
    from gsdf_slam_b200 import _lib
    e = torch.empty(0, device=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    bg = t(scene.bg)
    fwd = gs.RasterizeGaussiansCUDA(bg, params["means3D"], e, params["opacity"], params["scales"],
                                    params["rotations"], 1.0, e, t(cam.viewmatrix), t(cam.projmatrix),
                                    cam.tanfovx, cam.tanfovy, cam.H, cam.W, params["dc"], params["sh"],
                                    scene.degree, t(cam.campos), False, False)
    L = _lib.lib()
    N = cam.W * cam.H
    T = ((cam.W + 15) // 16) * ((cam.H + 15) // 16)
    n_contrib = torch.zeros(N, dtype=torch.int32, device=dev)
    ranges = torch.zeros((T, 2), dtype=torch.int32, device=dev)
    maxc = torch.zeros(T, dtype=torch.int32, device=dev)
    _lib.check(L.gsr_export_image(C.c_void_p(fwd[6].data_ptr()), cam.W, cam.H, None,
                                  C.c_void_p(n_contrib.data_ptr()), C.c_void_p(ranges.data_ptr()),
                                  C.c_void_p(maxc.data_ptr()), None), "export_image")
    pc = torch.zeros(3, dtype=torch.int64, device=dev)
    scratch = torch.empty((3, cam.H, cam.W), dtype=torch.float32, device=dev)
    if fwd[0] > 0:
        _lib.check(L.gsr_export_pair_counts(C.c_void_p(fwd[5].data_ptr()), C.c_void_p(fwd[6].data_ptr()), int(fwd[0]),
                                            cam.W, cam.H, C.c_void_p(bg.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                            C.c_void_p(pc.data_ptr()), None), "export_pair_counts")
    torch.cuda.synchronize()
    lens = (ranges[:, 1] - ranges[:, 0]).long()
    pcs = [int(x) for x in pc.tolist()]
    return dict(P=scene.P, V=int((fwd[3] > 0).sum()), R=int(fwd[0]), T=T, N=N, M=int(params["sh"].size(1)),
                E=int(n_contrib.long().sum()), max_tile_len=int(lens.max()),
                sum_max_contrib=int(maxc.long().sum()),
                pairs_visited=pcs[0], pairs_gated=pcs[1], pairs_blended=pcs[2])


# Algorithmic flops of the blend kernels, from COUNTED pairs (gsr_export_pair_counts), using the reference's own
# per-pair arithmetic (forward.cu:422-448, backward.cu:597-637; an FMA counts 2):
#   every visited pair  : d = xy - pix (2), power (5 mul + 2 fma = 9)                                  = 11
#   every gated pair    : expf (8 + MUFU), alpha = min(.99, o*G) (2), test_T (2)                        = 12
#   every blended pair  : fwd: 3 mul + 3 fma + T (10);  bwd: the gradient terms of backward.cu:609-637  = 10 / 49
FLOP_VISIT, FLOP_GATE, FLOP_BLEND_FWD, FLOP_BLEND_BWD = 11, 12, 10, 49


def stage_model(counts):
    """Algorithmic bytes / flops per launch of each stage."""
    P, V, R, N, T, M = (counts[k] for k in ("P", "V", "R", "N", "T", "M"))
    smc = counts["sum_max_contrib"]
    b = {
        # stage 1 (cull): xyz + scale + rotation rows of every Gaussian, radii / tiles_touched zeros of the culled ones
        "precull": 40 * P + 8 * P,
        # stage 2 (the reference's arithmetic on the candidates, ~ the rendered ones): candidate id + input rows +
        # SH row in, per-Gaussian outputs + list entry + 48-byte record out
        "preprocess": (4 + 56 + 12 * M) * V + 99 * V,
        "tile_meta": 4 * T + 16 * T,
        # emit: list entry + rect inputs per visible Gaussian, one slot atomic (RMW) + one item per duplicate
        "emit": (8 + 4 + 8 + 16 + 4) * V + (8 + 8) * R,
        # per-tile sort + pack: items in/out, 48-byte record gathered and written per duplicate
        "tile_sort_pack": (16 + 48 + 48) * R,
        # the blend kernels stop at the deepest contributor of a tile: 48 bytes per record up to there (each of
        # the two half tiles reads them: the second read hits L2, HBM sees one)
        "blend_fwd": 48 * min(R, smc) + 32 * N + 4 * T,
        "blend_bwd": 48 * min(R, smc) + 32 * N + 72 * min(R, smc),
        # moment rows of every rendered Gaussian (selection) + the full row traffic as if every rendered Gaussian
        # had a gradient (an upper bound: hidden splats are skipped)
        "preprocess_bwd": 56 * V + (87 + 12 * M) * V + 48 * V + (52 + 12 * M) * V + 4 * V,
        "adam": 28 * (14 + 3 * M) * P,
        # visibility-split Adam: rows without gradient move 24 B/element (grad not read)
        "adam_rows_zero_grad": 24 * (14 + 3 * M) * (P - V),
        "adam_rows_with_grad": 28 * (14 + 3 * M) * V,
        "adam_rows_all": (24 * P + 4 * V) * (14 + 3 * M),
    }
    vis, gat, ble = counts["pairs_visited"], counts["pairs_gated"], counts["pairs_blended"]
    f = {"blend_fwd": FLOP_VISIT * vis + FLOP_GATE * gat + FLOP_BLEND_FWD * ble,
         "blend_bwd": FLOP_VISIT * vis + FLOP_GATE * gat + FLOP_BLEND_BWD * ble}
    return b, f


def read_profile(L):
    buf = C.create_string_buffer(8192)
    L.gsr_profile_read(buf, 8192)
    prof = {}
    for line in buf.value.decode().splitlines():
        name, cnt, tot = line.split()
        prof[name] = (float(tot), int(cnt))
    return prof


def checksum(t: torch.Tensor) -> torch.Tensor:
    """Order-independent 2x64-bit digest of a tensor's bit pattern (replica comparison across ranks)."""
    v = t.contiguous().view(torch.int32).long()
    return torch.stack([v.sum(), (v * v % 1000003).sum()])


# --------------------------------------------------------------------------------------
def timed_loop(fn, steps, warmup, world, dev, L):
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        fn()
    barrier()
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
    launches0 = L.gsr_launch_count()
    t0 = time.time()
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    timed_loop.host_ms_per_step = (time.time() - t0) * 1e3 / max(1, steps)   # host time to ENQUEUE a step
    barrier()
    t1 = time.time()
    ms = a.elapsed_time(b)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    return ms, L.gsr_launch_count() - launches0, (t0, t1)


def rank_breakdown_of(mapper, step_fn, world, dev, n_tr=10):
    """Where the step time goes on every rank (render = fwd + bwd of this rank's views; wait = first barrier,
    i.e. waiting for the slowest rank; exchange = rank mask / statistics + fused peer kernel + second barrier)."""
    import torch.distributed as dist
    acc = torch.zeros(6, device=dev, dtype=torch.float64)
    for _ in range(n_tr):
        mapper.trace = []
        step_fn()
        torch.cuda.synchronize()
        ev = dict(mapper.trace)
        acc[0] += ev["step_begin"].elapsed_time(ev["render_end"])
        acc[1] += ev["render_end"].elapsed_time(ev["barrier1_end"])
        acc[2] += ev["barrier1_end"].elapsed_time(ev["step_end"])
        acc[3] += ev["barrier1_end"].elapsed_time(ev["mask_end"])
        acc[4] += ev["mask_end"].elapsed_time(ev["peer_adam_end"])
        acc[5] += ev["peer_adam_end"].elapsed_time(ev["step_end"])
    mapper.trace = None
    acc /= n_tr
    allr = [torch.zeros_like(acc) for _ in range(world)]
    dist.all_gather(allr, acc)
    return [{"rank": r, "render_ms": round(float(x[0]), 4), "wait_ms": round(float(x[1]), 4),
             "exchange_ms": round(float(x[2]), 4), "mask_stats_ms": round(float(x[3]), 4),
             "peer_adam_ms": round(float(x[4]), 4), "barrier2_ms": round(float(x[5]), 4)}
            for r, x in enumerate(allr)]


def exchange_summary(mapper, breakdown, step_fn, world, dev):
    """MEASURED NVLink traffic of the fused exchange (one extra, untimed step through the counting variant of the
    peer kernel) next to its measured time, and the remaining loss named."""
    import torch.distributed as dist
    if mapper.peer is None or not breakdown:
        return None
    tr = mapper.measure_exchange_traffic(step_fn)
    mine = torch.tensor([tr["grad_in_bytes"], tr["param_changed_bytes"]], device=dev, dtype=torch.float64)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    grad_in = np.array([float(x[0]) for x in allr])
    changed = np.array([float(x[1]) for x in allr])
    param_in = changed.sum() - changed                       # every other rank's changed groups land here
    bytes_in = grad_in + param_in
    adam_ms = np.array([b["peer_adam_ms"] for b in breakdown])
    wait_ms = float(np.mean([b["wait_ms"] for b in breakdown]))
    bar_ms = float(np.mean([b["barrier2_ms"] for b in breakdown]))
    mc = bool(mapper.peer["mc_ptr"])
    out = {"arena_bytes": int(4 * mapper.layout["_total"][1]), "multicast": mc,
           "grad_in_bytes_per_rank": [int(x) for x in grad_in],
           "param_changed_bytes_per_rank": [int(x) for x in changed],
           "bytes_in_per_rank_mean": int(bytes_in.mean()),
           "bytes_out_per_rank_mean": int((changed * (1 if mc else world - 1)).mean() + grad_in.mean()),
           "achieved_in_gbs": round(float((bytes_in / np.maximum(adam_ms, 1e-6)).mean()) / 1e6, 1),
           "nvlink_peak_gbs": NVLINK_PEER_GBS,
           "frac_of_nvlink_peak": round(float((bytes_in / np.maximum(adam_ms, 1e-6)).mean()) / 1e6 / NVLINK_PEER_GBS, 3),
           "note": ("counted in the kernel: 16-byte gradient groups loaded from other ranks (responses arrive on this "
                    "rank's ingress; the same amount leaves it on average as served loads) and 16-byte parameter groups "
                    "whose value changed; unchanged groups (zero moments and zero gradient) are not sent; bytes_in / "
                    "peer_adam_ms per rank, averaged")}
    losses = {"peer kernel": float(adam_ms.mean()), "barrier wait (slowest view)": wait_ms, "second barrier": bar_ms}
    out["largest_remaining_loss"] = max(losses, key=losses.get)
    return out


# --------------------------------------------------------------------------------------
def rasterizer_only_product(scene, cam, params, dev, steps, warmup):
    """RasterizeGaussiansCUDA + RasterizeGaussiansBackwardCUDA through the C++ drop-in shim (dense gradients, the
    reference's allocation pattern) — the metric's literal quantity."""
    shim = os.path.join(ROOT, "gsdf_slam_b200", "libgsr_torch_shim.so")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    e = torch.empty(0, dtype=torch.float32, device=dev)
    bg, view, proj, campos, dL = t(scene.bg), t(cam.viewmatrix), t(cam.projmatrix), t(cam.campos), t(scene.dL_dout)
    api = "C++ shim (libgsr_torch_shim.so: RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA)"
    try:
        torch.ops.load_library(shim)
        ops = torch.ops.gsr_shim

        def once():
            r = ops.rasterize_forward(bg, params["means3D"], e, params["opacity"], params["scales"], params["rotations"],
                                      1.0, e, view, proj, cam.tanfovx, cam.tanfovy, cam.H, cam.W, params["dc"],
                                      params["sh"], scene.degree, campos, False, False)
            ops.rasterize_backward(bg, params["means3D"], r[3], e, params["scales"], params["rotations"], 1.0, e, view,
                                   proj, cam.tanfovx, cam.tanfovy, dL, params["dc"], params["sh"], scene.degree, campos,
                                   r[4], r[0], r[5], r[6], r[1], r[7], False)
        once()
    except Exception as exc:  # noqa: BLE001
        import gsdf_slam_b200 as gs
        api = f"ctypes mirror (shim unavailable: {exc!r})"

        def once():
            r = gs.RasterizeGaussiansCUDA(bg, params["means3D"], e, params["opacity"], params["scales"],
                                          params["rotations"], 1.0, e, view, proj, cam.tanfovx, cam.tanfovy, cam.H,
                                          cam.W, params["dc"], params["sh"], scene.degree, campos, False, False)
            gs.RasterizeGaussiansBackwardCUDA(bg, params["means3D"], r[3], e, params["scales"], params["rotations"],
                                              1.0, e, view, proj, cam.tanfovx, cam.tanfovy, dL, params["dc"],
                                              params["sh"], scene.degree, campos, r[4], r[0], r[5], r[6], r[1], r[7],
                                              False)
    for _ in range(warmup):
        once()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        once()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return {"value": round(1000.0 / ms, 2), "unit": "rasterizer fwd+bwd it/s", "ms_per_iteration": round(ms, 4),
            "api": api, "gradients": "dense [P, ...] tensors (zero rows for Gaussians that were not rendered)"}


# --------------------------------------------------------------------------------------
def cfg5_keyframes(scene5, dev, dL):
    from gsdf_slam_b200.mapping import Keyframe
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    return [Keyframe(t(c.viewmatrix), t(c.projmatrix), t(c.campos), c.tanfovx, c.tanfovy, c.H, c.W, dL)
            for c in scene5.extra_cams]


def zero_stats(P, dev):
    return {k: torch.zeros(P, dtype=torch.float32, device=dev)
            for k in ("max_radii2D", "xyz_gradient_accum", "denom")}


def run_cfg5(args, rank, world, dev, L, scene5, params5, kfs):
    """BASELINE.json configs[4]: K = 8 keyframes per optimiser step over 3 M Gaussians, split over the ranks."""
    from gsdf_slam_b200.mapping import KeyframeBatchMapper, partition_keyframes
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    local = [kfs[i] for i in partition_keyframes(len(kfs), world, rank)]
    mapper = KeyframeBatchMapper(params5, scene5.degree, t(scene5.bg), freeze_render_inputs=True,
                                 stats=zero_stats(scene5.P, dev))
    steps, warmup = max(5, min(args.steps, 12)), 3

    def step():
        mapper.step(local)

    ms, launches, _ = timed_loop(step, steps, warmup, world, dev, L)
    ms_step = ms / steps
    out = {"workload": f"{scene5.name}: {scene5.P} SH-{scene5.degree} Gaussians, {len(kfs)} keyframes "
                       f"({scene5.extra_cams[0].W}x{scene5.extra_cams[0].H}) per optimiser step split over {world} "
                       "rank(s): fwd + bwd of the local keyframes (rendered rows only, later keyframes add), "
                       "densification statistics of every view, exchange, Adam",
           "scaling": "strong (fixed batch of 8 keyframes)", "n_gpus": world, "keyframes_per_step": len(kfs),
           "keyframes_per_rank": len(local), "steps": steps, "ms_per_step": round(ms_step, 4),
           "value": round(1000.0 / ms_step, 2), "unit": "mapping it/s (optimiser steps of 8 keyframes)",
           "keyframes_per_s": round(1000.0 * len(kfs) / ms_step, 1), "gpu_launches": int(launches),
           "exchange": ("none (single rank): masked Adam over the rows some keyframe rendered" if world == 1 else
                        "fused peer-memory kernel (gsr_adam_step_peer: gradients, Adam, parameters, densification statistics)"
                        if mapper.peer is not None else "NCCL all-reduce fallback")}
    n1 = os.path.join(ROOT, "profiles", "mapping_cfg5_n1.json")
    if world > 1 and os.path.exists(n1) and not (args.small_cfg5 or args.config == "small"):
        try:
            ref = json.load(open(n1))
            out["n1_ms_per_step_recorded"] = ref["ms_per_step"]
            out["speedup_vs_recorded_n1"] = round(ref["ms_per_step"] / ms_step, 3)
            out["n1_record"] = ("profiles/mapping_cfg5_n1.json (an earlier N=1 run of this block; the driver's own "
                                "N=1 line is authoritative)")
        except Exception:  # noqa: BLE001
            pass
    # per-stage CUDA-event times of this rank's part of the step (untimed extra steps, rank 0's numbers)
    L.gsr_profile_enable(1)
    read_profile(L)
    mapper.serialize_side_stream = True
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    mapper.serialize_side_stream = False
    out["stages_ms_per_step_rank0"] = {k: round(tot / 3.0, 4) for k, (tot, cnt) in sorted(
        read_profile(L).items(), key=lambda kv: -kv[1][0])}
    L.gsr_profile_enable(0)
    if world > 1 and mapper.peer is not None:
        bd = rank_breakdown_of(mapper, step, world, dev, n_tr=5)
        out["rank_breakdown"] = bd
        out["nvlink"] = exchange_summary(mapper, bd, step, world, dev)
    del mapper
    torch.cuda.empty_cache()
    return out


def run_exchange_parity(rank, world, dev, scene5, params5, kfs):
    """Untimed: the same training steps (render inputs NOT frozen) three ways — fused peer exchange, NCCL all-reduce
    path, one rank accumulating all keyframes sequentially — twice:
      * ONE step with eps = 1 ("linear"): the Adam update is then proportional to the reduced gradient (denominator
        ~ 1), so the differences measure the EXCHANGE itself (bar: 1e-5, fp32 reduction order + the float atomics of
        the backward blend, whose run-to-run `noise_floor` is printed too);
      * TWO steps with eps = 1e-15 (the reference's): Adam's first steps are ~ lr * sign(g), so the handful of
        elements whose summed gradient is ~0 flip with the order of the float atomics — in ANY implementation — and
        the second step renders slightly different parameters; the run-to-run difference of the single-rank path
        (`noise_floor`) is printed next to the cross-path numbers."""
    import torch.distributed as dist
    from gsdf_slam_b200.mapping import SEGMENTS, KeyframeBatchMapper, partition_keyframes
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    bg = t(scene5.bg)
    local = [kfs[i] for i in partition_keyframes(len(kfs), world, rank)]

    def run(kw, batch, eps, n_steps, poison=False):
        st = zero_stats(scene5.P, dev)
        m = KeyframeBatchMapper({k: v.clone() for k, v in params5.items()}, scene5.degree, bg, stats=st, eps=eps, **kw)
        if kw.get("peer_exchange") and m.peer is None:
            del m
            return None
        if poison:
            m.arena.fill_(float("nan"))      # rows nobody renders must never be read
        for _ in range(n_steps):
            m.step(batch)
        torch.cuda.synchronize()
        out = ({k: m.opt_params[k].clone() for k in SEGMENTS}, {k: v.clone() for k, v in st.items()})
        del m
        torch.cuda.empty_cache()
        return out

    def upd_rel(a, b):
        worst = 0.0
        for k in SEGMENTS:
            d0 = (b[k] - params5[k]).double()
            worst = max(worst, float((a[k].double() - b[k].double()).norm() / max(float(d0.norm()), 1e-30)))
        return float(f"{worst:.3e}")

    out = {"workload": f"{scene5.name}, {len(kfs)} keyframes per step over {world} ranks, "
                       "training steps with live parameters"}
    for tag, eps, n_steps in (("linear_eps1", 1.0, 1), ("reference_eps1e-15", 1e-15, 2)):
        peer = run(dict(peer_exchange=True), local, eps, n_steps, poison=True)
        nccl = run(dict(peer_exchange=False), local, eps, n_steps)
        seq = run(dict(single_rank=True), kfs, eps, n_steps) if rank == 0 else None
        seq2 = run(dict(single_rank=True), kfs, eps, n_steps) if rank == 0 else None
        dist.barrier()
        blk = {"steps": n_steps}
        if peer is None:
            out["peer_path"] = "unavailable (no symmetric memory)"
            return out
        pa, sa = peer
        pb, sb = nccl
        blk["peer_vs_nccl_rel_l2"] = upd_rel(pa, pb)
        digest = torch.cat([checksum(pa[k]) for k in SEGMENTS] + [checksum(sa[k]) for k in sa])
        allg = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(allg, digest)
        blk["replicas_bit_identical"] = bool(all(torch.equal(allg[0], x) for x in allg))
        blk["finite"] = all(bool(torch.isfinite(pa[k]).all()) for k in SEGMENTS)
        if rank == 0:
            pc, sc = seq
            blk["vs_sequential_rel_l2"] = upd_rel(pa, pc)
            blk["nccl_vs_sequential_rel_l2"] = upd_rel(pb, pc)
            if seq2 is not None:
                blk["noise_floor_sequential_vs_itself_rel_l2"] = upd_rel(seq2[0], pc)
            blk["stats_denom_equal"] = bool(torch.equal(sa["denom"], sc["denom"]))
            blk["stats_max_radii_equal"] = bool(torch.equal(sa["max_radii2D"], sc["max_radii2D"]))
            den = float(sc["xyz_gradient_accum"].double().norm())
            num = float((sa["xyz_gradient_accum"].double() - sc["xyz_gradient_accum"].double()).norm())
            blk["stats_accum_rel_l2"] = float(f"{num / max(den, 1e-30):.3e}")
            blk["nccl_stats_denom_equal"] = bool(torch.equal(sb["denom"], sc["denom"]))
            blk["rows_with_statistics"] = int((sc["denom"] > 0).sum())
        out[tag] = blk
        del peer, nccl, seq, seq2
        torch.cuda.empty_cache()
    lin = out.get("linear_eps1", {})
    out["peer_vs_nccl_rel_l2"] = lin.get("peer_vs_nccl_rel_l2")
    out["vs_sequential_rel_l2"] = lin.get("vs_sequential_rel_l2")
    out["replicas_bit_identical"] = bool(lin.get("replicas_bit_identical")) and \
        bool(out.get("reference_eps1e-15", {}).get("replicas_bit_identical"))
    out["bar"] = ("linear_eps1: rel-L2 of the parameter update <= 1e-5 (fp32 reduction order); replicas bit-identical; "
                  "denom / max_radii exact; reference_eps1e-15: at its own noise floor")
    return out


# --------------------------------------------------------------------------------------
def run_product(args, rank, world, dev):
    import torch.distributed as dist
    from gsdf_slam_b200 import _lib
    from gsdf_slam_b200.mapping import Keyframe, KeyframeBatchMapper

    L = _lib.lib()
    scene, cam = make_scene(args.config, rank)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    params = dev_params(scene, dev)
    mapper = KeyframeBatchMapper(params, scene.degree, t(scene.bg), freeze_render_inputs=True)
    hk = host_keyframe(scene, cam, args.with_loss)
    big_dev = hk["dL_dout"].to(dev)

    def make_kf(view, proj, campos, big):
        if args.with_loss:
            return Keyframe(view, proj, campos, cam.tanfovx, cam.tanfovy, cam.H, cam.W, gt_image=big)
        return Keyframe(view, proj, campos, cam.tanfovx, cam.tanfovy, cam.H, cam.W, big)

    kf_dev = make_kf(t(cam.viewmatrix), t(cam.projmatrix), t(cam.campos), big_dev)
    counts = workload_counts(scene, cam, params, dev)
    copy_stream = torch.cuda.Stream(device=dev)    # device -> host (rendered image)
    h2d_stream = torch.cuda.Stream(device=dev)     # host -> device: its own stream, PCIe is full duplex and the
                                                   # upload of step k + 1 must not queue behind the read-back of step k

    def step_resident():
        mapper.step([kf_dev])

    # device staging buffers for the e2e path + pinned result buffers (both double-buffered)
    stage = {k: torch.empty_like(v, device=dev) for k, v in hk.items()}
    cam_host, cam_dev, cam_views = camera_block(hk, dev)
    stage.update(cam_views)
    img_host = [torch.empty((3, cam.H, cam.W), dtype=torch.float32).pin_memory() for _ in range(2)]
    h2d_bytes = sum(v.numel() * 4 for v in hk.values())
    d2h_bytes = img_host[0].numel() * 4
    ev_h2d = torch.cuda.Event()
    ev_fwd = torch.cuda.Event()
    ev_step = [torch.cuda.Event(), torch.cuda.Event()]
    ev_img = [torch.cuda.Event(), torch.cuda.Event()]
    for ev in ev_step + ev_img:
        ev.record(torch.cuda.current_stream(dev))
    dl_stage = [stage["dL_dout"], torch.empty_like(stage["dL_dout"])]
    counter = [0]
    wait_s = [0.0]      # host time spent waiting for the image of step k-2 (the rest of the loop is enqueue work)

    def step_e2e():
        main = torch.cuda.current_stream(dev)
        slot = counter[0] & 1
        counter[0] += 1
        # the result of step k-2 (the rendered image in img_host[slot]) must have reached the host before its
        # buffer is reused; the host never blocks on the step it has just enqueued
        tw = time.time()
        ev_img[slot].synchronize()
        wait_s[0] += time.time() - tw
        # camera + background first (tiny), the image-sized gradient on the copy stream so it
        # overlaps the forward (a real mapper uploads the ground-truth image the same way)
        cam_dev.copy_(cam_host, non_blocking=True)
        # the image-sized upload is double-buffered: it may start while the previous step's
        # backward pass still reads the other buffer (two steps back is finished: ev_step[slot])
        h2d_stream.wait_event(ev_step[slot])
        with torch.cuda.stream(h2d_stream):
            dl_stage[slot].copy_(hk["dL_dout"], non_blocking=True)
            ev_h2d.record(h2d_stream)
        mapper.bg = stage["bg"]
        kf = make_kf(stage["viewmatrix"], stage["projmatrix"], stage["campos"], dl_stage[slot])

        def after_forward(_kf, fwd):
            # read the image back while the backward runs; the backward needs the uploaded dL_dout
            color = fwd[2]
            ev_fwd.record(main)
            copy_stream.wait_event(ev_fwd)
            with torch.cuda.stream(copy_stream):
                img_host[slot].copy_(color, non_blocking=True)
                ev_img[slot].record(copy_stream)
            color.record_stream(copy_stream)
            main.wait_event(ev_h2d)

        mapper.step([kf], on_forward=after_forward)
        ev_step[slot].record(main)

    sampler = ClockSampler(dev.index if dev.index is not None else 0)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ms, launches, (t0, t1) = timed_loop(step_resident, args.steps, args.warmup, world, dev, L)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    for _ in range(3):
        step_e2e()
    wait_s[0] = 0.0
    e2e_ms, _, _ = timed_loop(step_e2e, args.steps, max(0, args.warmup // 2), world, dev, L)
    e2e_host_ms = timed_loop.host_ms_per_step
    e2e_wait_ms = wait_s[0] * 1e3 / max(1, args.steps + max(0, args.warmup // 2))
    for ev in ev_img:
        ev.synchronize()

    # per-stage durations with CUDA events on the launching stream (separate, untimed loop)
    L.gsr_profile_enable(1)
    mapper.serialize_side_stream = True   # per-stage times are taken without stream overlap
    for _ in range(max(5, min(args.steps, 20))):
        step_resident()
    torch.cuda.synchronize()
    mapper.serialize_side_stream = False
    prof = {k: tot / max(1, cnt) for k, (tot, cnt) in read_profile(L).items()}
    L.gsr_profile_enable(0)

    rank_breakdown = None
    if world > 1 and mapper.peer is not None:
        rank_breakdown = rank_breakdown_of(mapper, step_resident, world, dev)
    nvlink = exchange_summary(mapper, rank_breakdown, step_resident, world, dev) if rank_breakdown else None
    if world > 1:
        lt = torch.tensor([launches], device=dev, dtype=torch.float64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    peer_active = mapper.peer is not None

    # ---- the metric's literal quantity through the C++ drop-in shim (N = 1) ----
    rast_only = None
    if world == 1 and not args.no_shim:
        rast_only = rasterizer_only_product(scene, cam, params, dev, max(10, min(args.steps, 30)), 5)
    del mapper
    torch.cuda.empty_cache()

    # ---- configs[4]: 3 M Gaussians, 8 keyframes per step, strong scaling; N > 1: exchange parity ----
    cfg5_block = parity_block = None
    if not args.no_cfg5:
        scene5 = make_cfg5(args.small_cfg5 or args.config == "small")
        params5 = dev_params(scene5, dev)
        kfs = cfg5_keyframes(scene5, dev, t(scene5.dL_dout))
        cfg5_block = run_cfg5(args, rank, world, dev, L, scene5, params5, kfs)
        if world > 1 and not args.no_parity:
            parity_block = run_exchange_parity(rank, world, dev, scene5, params5, kfs)
        del params5, kfs
        torch.cuda.empty_cache()
    if rank != 0:
        return None

    hbm_peak, peak_src = load_peaks()
    bytes_model, flops_model = stage_model(counts)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    fp32_peak_tflops = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    traffic = {}
    if os.path.exists(traffic_file):
        traffic = json.load(open(traffic_file)).get(args.config, {})
        if not isinstance(traffic, dict):
            traffic = {}
    stages = {}
    for name, ms_stage in sorted(prof.items(), key=lambda kv: -kv[1]):
        ent = {"ms": round(ms_stage, 5)}
        if name in bytes_model:
            gbs = bytes_model[name] / (ms_stage * 1e-3) / 1e9
            ent.update(algorithmic_bytes=int(bytes_model[name]), achieved_gbs=round(gbs, 1),
                       frac_of_hbm_peak=round(gbs / hbm_peak, 4))
        if name in flops_model:
            tf = flops_model[name] / (ms_stage * 1e-3) / 1e12
            ent.update(algorithmic_flops=int(flops_model[name]), achieved_tflops=round(tf, 3),
                       frac_of_fp32_peak=round(tf / fp32_peak_tflops, 4))
        if name in traffic:
            ent["dram_bytes_ncu"] = traffic[name]
        stages[name] = ent
    modelled = {k: v for k, v in prof.items() if k in bytes_model or k in flops_model}
    dom = max(modelled, key=modelled.get) if modelled else None
    roofline = None
    if dom is not None:
        if dom in flops_model:
            ach = flops_model[dom] / (prof[dom] * 1e-3) / 1e12
            roofline = {"kernel": dom, "bound": "fp32", "achieved": round(ach, 3),
                        "peak": round(fp32_peak_tflops, 2), "unit": "TFLOP/s",
                        "frac": round(ach / fp32_peak_tflops, 4), "traffic": traffic.get(dom),
                        "peak_source": f"148 SMs x 128 FP32 lanes x 2 x {sm_mhz:.0f} MHz (clock sampled "
                                       "during the run); no tensor-core work exists on this path",
                        "flops_model": f"{FLOP_VISIT}*visited + {FLOP_GATE}*gated + "
                                       f"{FLOP_BLEND_BWD if dom == 'blend_bwd' else FLOP_BLEND_FWD}*blended pairs, pairs "
                                       "COUNTED by an instrumented replay (counts.pairs_*), the reference's per-pair arithmetic",
                        "hbm_achieved_gbs": stages[dom].get("achieved_gbs"),
                        "share_of_step": round(prof[dom] / sum(prof.values()), 3)}
        else:
            ach = bytes_model[dom] / (prof[dom] * 1e-3) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak,
                        "unit": "GB/s", "frac": round(ach / hbm_peak, 4), "traffic": traffic.get(dom),
                        "peak_source": peak_src}

    # largest HBM-bound stage, against the measured copy bandwidth
    hbm_stages = {k: v for k, v in stages.items() if k not in flops_model and "achieved_gbs" in v}
    roofline_hbm = None
    if hbm_stages:
        k = max(hbm_stages, key=lambda n: hbm_stages[n]["algorithmic_bytes"])
        roofline_hbm = {"kernel": k, "bound": "hbm", "achieved": hbm_stages[k]["achieved_gbs"],
                        "peak": hbm_peak, "unit": "GB/s", "frac": hbm_stages[k]["frac_of_hbm_peak"],
                        "traffic": traffic.get(k), "peak_source": peak_src}

    ms_step = ms / args.steps
    e2e_step = e2e_ms / args.steps
    line = {
        "metric": METRIC, "value": round(world * 1000.0 / ms_step, 2), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.config}: {scene.name}, one keyframe per GPU per step, "
                               + ("fwd + fused L1+SSIM loss + bwd + exchange + Adam; activated inputs (the drop-in API)"
                                  if args.with_loss else
                                  "fwd+bwd+(exchange)+fused Adam on activated inputs (the drop-in API); --with-loss adds the fused L1+SSIM loss"),
                   "P": counts["P"], "sh_degree": scene.degree, "image": [cam.W, cam.H],
                   "keyframes_per_step": world, "parallelism": f"dp{world} (keyframe batch)",
                   "exchange": ("none (single rank)" if world == 1 else
                                "fused peer-memory kernel: gradient reduction + Adam on a 1/N shard + parameter "
                                "broadcast over NVLink (gsr_adam_step_peer)" if peer_active else
                                "NCCL all-reduce of the gradient arena, pipelined with fused Adam"),
                   "l2_policy": "working set (params+grads+Adam state+scratch ~0.7 GB) exceeds the 126 MB L2",
                   "adam": ("fused, on a raw-parameter shadow copy (render inputs frozen); single rank: one "
                            "masked pass (gsr_adam_step_rows ALL) that reads the gradient only for the "
                            "rendered rows, the backward pass skips the dense zero fill")},
        "counts": counts,
        "e2e": {"value": round(world * 1000.0 / e2e_step, 2), "unit": UNIT, "ms_per_step": round(e2e_step, 4),
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "host_loop_ms_per_step": round(e2e_host_ms, 4), "host_wait_ms_per_step": round(e2e_wait_ms, 4),
                "pipelining": "uploads and read-backs double-buffered on two copy streams; the image of step k is awaited before step k+2"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_hbm": roofline_hbm,
        "stages": stages,
    }
    if rast_only is not None:
        line["rasterizer_only"] = rast_only
    if rank_breakdown is not None:
        line["rank_breakdown"] = rank_breakdown
        line["nvlink"] = nvlink
    if cfg5_block is not None:
        line["mapping_cfg5"] = cfg5_block
    if parity_block is not None:
        line["exchange_parity"] = parity_block
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.config)
    return line


def cpu_baseline(cfg: str):
    """The oracle port (scalar C++ restatement, OpenMP) on the host cores: bounded sample."""
    from oracle import cpu_binding as ob
    scene, cam = make_scene(cfg, 0)
    orc = ob.CpuOracle()
    n, t_tot = 0, 0.0
    while n < 2 or (t_tot < 8.0 and n < 6):
        t0 = time.perf_counter()
        orc.forward(scene)
        orc.backward(scene.dL_dout)
        t_tot += time.perf_counter() - t0
        n += 1
    return {"value": round(n / t_tot, 4), "unit": "keyframe it/s (fwd+bwd, no Adam)", "cores": orc.threads,
            "kind": "port", "sample": f"{n} full fwd+bwd iterations of {cfg} ({scene.P} Gaussians, "
                                      f"{cam.W}x{cam.H}), OpenMP over tiles/Gaussians"}


# --------------------------------------------------------------------------------------
class LibtorchAdam:
    """torch::optim::Adam::step (single-tensor path) restated with the same ATen ops: the reference's optimizer."""

    def __init__(self, p, names):
        from gsdf_slam_b200.mapping import DEFAULT_LR
        self.lr = DEFAULT_LR
        self.names = names
        self.shadow = {k: p[k].clone() for k in names}
        self.m = {k: torch.zeros_like(p[k]) for k in names}
        self.v = {k: torch.zeros_like(p[k]) for k in names}
        self.step_count = 0

    def step(self, grads, gname):
        self.step_count += 1
        s = self.step_count
        bc1, bc2 = 1 - 0.9 ** s, 1 - 0.999 ** s
        for k in self.names:
            g = grads[gname[k]].view_as(self.shadow[k])
            self.m[k].mul_(0.9).add_(g, alpha=0.1)
            self.v[k].mul_(0.999).addcmul_(g, g, value=0.001)
            denom = (self.v[k].sqrt() / np.sqrt(bc2)).add_(1e-15)
            self.shadow[k].addcdiv_(self.m[k], denom, value=-(self.lr[k] / bc1))


GRAD_OF = dict(means3D="dL_dmeans3D", dc="dL_ddc", sh="dL_dsh", opacity="dL_dopacity", scales="dL_dscales",
               rotations="dL_drotations")
PARAM_NAMES = ("means3D", "dc", "sh", "opacity", "scales", "rotations")


def run_reference(args, rank, world, dev):
    """Reference arm: the unmodified reference CUDA rasterizer (oracle/_ref) + LibTorch's Adam
    op sequence, rank 0 only."""
    if rank != 0:
        return None
    from oracle import ref_binding as rb
    if not rb.available():
        # the reference compiled in the build container; if the .so did not travel, fall back
        # to the CPU restatement so the arm still reports a number
        cb = cpu_baseline(args.config)
        return {"metric": METRIC, "impl": "reference", "value": cb["value"], "unit": UNIT, "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1000.0 / cb["value"], 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": args.config}, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    scene, cam = make_scene(args.config, 0)
    p = dev_params(scene, dev)
    e = torch.empty(0, dtype=torch.float32, device=dev)
    rr = rb.RefRasterizer()
    hk = host_keyframe(scene, cam, args.with_loss)
    stage = {k: torch.empty_like(v, device=dev) for k, v in hk.items()}
    cam_host, cam_dev, cam_views = camera_block(hk, dev)   # same single camera copy per step as the product arm
    stage.update(cam_views)
    img_host = [torch.empty((3, cam.H, cam.W), dtype=torch.float32).pin_memory() for _ in range(2)]
    from oracle.loss_ref import reference_mapping_loss
    opt = LibtorchAdam(p, PARAM_NAMES)

    copy_stream = torch.cuda.Stream(device=dev)    # device -> host
    h2d_stream = torch.cuda.Stream(device=dev)     # host -> device (as in the product arm)
    ev_h2d, ev_fwd = torch.cuda.Event(), torch.cuda.Event()
    ev_step = [torch.cuda.Event(), torch.cuda.Event()]
    ev_img = [torch.cuda.Event(), torch.cuda.Event()]
    for ev in ev_step + ev_img:
        ev.record(torch.cuda.current_stream(dev))
    dl_stage = [stage["dL_dout"], torch.empty_like(stage["dL_dout"])]
    counter = [0]

    def step(debug=False, host_io=True, adam=True):
        # host I/O exactly like the product's e2e step: camera block on the compute stream, the
        # image-sized gradient up and the rendered image down on two copy streams, double-buffered, the image of
        # step k awaited before step k+2
        src = stage
        main = torch.cuda.current_stream(dev)
        dl = stage["dL_dout"]
        slot = 0
        if host_io:
            slot = counter[0] & 1
            counter[0] += 1
            ev_img[slot].synchronize()
            cam_dev.copy_(cam_host, non_blocking=True)
            dl = dl_stage[slot]
            h2d_stream.wait_event(ev_step[slot])
            with torch.cuda.stream(h2d_stream):
                dl.copy_(hk["dL_dout"], non_blocking=True)
                ev_h2d.record(h2d_stream)
        # allocations exactly like the reference shim (src/rasterize_points.cu:74-76, 157-166)
        R, B, color, radii = rr.forward(src["bg"], p["means3D"], e, p["opacity"], p["scales"], p["rotations"],
                                        1.0, e, src["viewmatrix"], src["projmatrix"], cam.tanfovx,
                                        cam.tanfovy, cam.H, cam.W, p["dc"], p["sh"], scene.degree,
                                        src["campos"], debug=debug)
        if host_io:
            ev_fwd.record(main)
            copy_stream.wait_event(ev_fwd)
            with torch.cuda.stream(copy_stream):
                img_host[slot].copy_(color, non_blocking=True)
                ev_img[slot].record(copy_stream)
            color.record_stream(copy_stream)
            main.wait_event(ev_h2d)
        if args.with_loss:
            dl = reference_mapping_loss(color, dl, None, 0.2)[3]   # the reference's own loss_utils.h + autograd
        g = rr.backward(src["bg"], p["means3D"], radii, e, p["scales"], p["rotations"], 1.0, e,
                        src["viewmatrix"], src["projmatrix"], cam.tanfovx, cam.tanfovy, dl,
                        p["dc"], p["sh"], scene.degree, src["campos"], debug=debug)
        if adam:
            opt.step(g, GRAD_OF)
        if host_io:
            ev_step[slot].record(main)

    for k in hk:
        stage[k].copy_(hk[k])
    cam_dev.copy_(cam_host)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    ms_e2e = timed(lambda: step(False, True), args.steps, args.warmup)
    ms_res = timed(lambda: step(False, False), args.steps, 3)
    ms_dbg = timed(lambda: step(True, False), max(3, args.steps // 2), 2)
    ms_rast = timed(lambda: step(False, False, adam=False), max(10, min(args.steps, 30)), 3)
    h2d = sum(x.numel() * 4 for x in hk.values())
    d2h = img_host[0].numel() * 4
    line = {"metric": METRIC, "impl": "reference", "value": round(1000.0 / ms_res, 2), "unit": UNIT,
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_res, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: {scene.name}; UNMODIFIED reference CUDA rasterizer "
                                   "recompiled for sm_100a (oracle/_ref) + LibTorch Adam op sequence"
                                   + (" + the reference's loss_utils.h L1+SSIM with autograd" if args.with_loss else "") + "; "
                                   "`value` device-resident like the product's, `e2e` with the product's host I/O",
                       "reference_debug_flag": False},
            "e2e": {"value": round(1000.0 / ms_e2e, 2), "unit": UNIT, "ms_per_step": round(ms_e2e, 4),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "value_device_resident_debug_true_as_shipped": round(1000.0 / ms_dbg, 2),
            "rasterizer_only": {"value": round(1000.0 / ms_rast, 2), "unit": "rasterizer fwd+bwd it/s",
                                "ms_per_iteration": round(ms_rast, 4),
                                "api": "Rasterizer::forward / backward with the reference shim's zero-filled allocations"}}
    del rr, opt
    torch.cuda.empty_cache()
    if not args.no_cfg5:
        line["mapping_cfg5"] = reference_cfg5(args, dev)
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.config)
    return line


def reference_cfg5(args, dev):
    """configs[4] on the reference arm: ONE GPU, the 8 keyframes processed sequentially with gradient accumulation
    (the reference has no multi-GPU mode), LibTorch Adam op sequence."""
    from oracle import ref_binding as rb
    scene5 = make_cfg5(args.small_cfg5 or args.config == "small")
    p = dev_params(scene5, dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    e = torch.empty(0, dtype=torch.float32, device=dev)
    bg, dL = t(scene5.bg), t(scene5.dL_dout)
    cams = [(t(c.viewmatrix), t(c.projmatrix), t(c.campos), c) for c in scene5.extra_cams]
    opt = LibtorchAdam(p, PARAM_NAMES)
    rr = rb.RefRasterizer()

    def step():
        acc = None
        for view, proj, campos, c in cams:
            R, B, color, radii = rr.forward(bg, p["means3D"], e, p["opacity"], p["scales"], p["rotations"], 1.0, e,
                                            view, proj, c.tanfovx, c.tanfovy, c.H, c.W, p["dc"], p["sh"],
                                            scene5.degree, campos)
            g = rr.backward(bg, p["means3D"], radii, e, p["scales"], p["rotations"], 1.0, e, view, proj, c.tanfovx,
                            c.tanfovy, dL, p["dc"], p["sh"], scene5.degree, campos)
            if acc is None:
                acc = {GRAD_OF[k]: g[GRAD_OF[k]] for k in PARAM_NAMES}
            else:
                for k in PARAM_NAMES:          # .grad accumulation of K backward() calls without zero_grad
                    acc[GRAD_OF[k]].add_(g[GRAD_OF[k]])
        opt.step(acc, GRAD_OF)

    steps = max(3, min(args.steps, 6))
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    return {"workload": f"{scene5.name}: {scene5.P} Gaussians, {len(cams)} keyframes per optimiser step, ONE GPU, "
                        "sequential with gradient accumulation (reference rasterizer + LibTorch Adam ops)",
            "n_gpus": 1, "steps": steps, "ms_per_step": round(ms, 4), "value": round(1000.0 / ms, 2),
            "unit": "mapping it/s (optimiser steps of 8 keyframes)",
            "keyframes_per_s": round(1000.0 * len(cams) / ms, 1)}


def main():
    # Libraries (NCCL's version banner, torchrun notices) may write to stdout; the contract is
    # ONE JSON line there, so everything else is sent to stderr and the line is written to the
    # real stdout at the end.
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the mapping_cfg5 block (3 M Gaussians, 8 keyframes per step)")
    ap.add_argument("--small-cfg5", action="store_true", help="CI-sized stand-in for the cfg5 scene")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed exchange_parity block (N > 1)")
    ap.add_argument("--no-shim", action="store_true", help="skip the rasterizer_only block (C++ drop-in shim, N = 1)")
    ap.add_argument("--with-loss", action="store_true",
                    help="include the photometric L1+SSIM loss (SURVEY row f-1) in the step: fused kernels in "
                         "the product arm, the reference's own loss_utils.h + autograd in the reference arm; "
                         "the image-sized input is then the ground-truth image instead of dL/d(image)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rasterizer hot path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    if args.impl == "reference":
        if rank == 0:
            emit(run_reference(args, rank, world, dev))
        return

    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)
    line = run_product(args, rank, world, dev)
    if rank == 0:
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
