"""Developer tool (one GPU): gsr_adam_step_peer with n_peers = 1 — the kernel's LOCAL efficiency (shard = the whole
arena, no NVLink): time per launch and achieved HBM bandwidth (24 B per element: p, m, v read and written).
    python tools/peer_adam_bench.py [P] [rendered_fraction]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gsdf_slam_b200 import _lib

P = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
L = _lib.lib()
dev = torch.device("cuda")
widths = [3, 3, 45, 1, 3, 4]
lrs = [1.6e-4, 0.0025, 0.0025 / 20, 0.05, 0.005, 0.001]
total = sum(widths) * P
g = torch.Generator(device=dev).manual_seed(1)
param = torch.randn(total, device=dev, generator=g)
grad = torch.randn(total, device=dev, generator=g) * 1e-3
m = torch.zeros(total, device=dev); v = torch.zeros(total, device=dev)
mask = (torch.rand(P, device=dev, generator=g) < frac).to(torch.int32)
a = _lib.PeerAdamArgs()
a.n_peers, a.rank = 1, 0
a.param_peers[0] = param.data_ptr(); a.grad_peers[0] = grad.data_ptr()
a.n_segments = 6
off = 0
for i, (w, lr) in enumerate(zip(widths, lrs)):
    a.segments[i].start, a.segments[i].numel, a.segments[i].width, a.segments[i].lr = off, w * P, w, lr
    off += w * P
a.shard_lo, a.shard_hi, a.shard_chunk = 0, total, 0
a.exp_avg, a.exp_avg_sq, a.union_mask = m.data_ptr(), v.data_ptr(), mask.data_ptr()
a.beta1, a.beta2, a.eps, a.grad_scale = 0.9, 0.999, 1e-15, 1.0
a.stream = torch.cuda.current_stream().cuda_stream
for variant in ("fresh state (rows without gradient keep zero moments)", "steady state (every row has moments)"):
    if variant.startswith("steady"):
        m.normal_(generator=g).mul_(1e-3); v.normal_(generator=g).abs_().mul_(1e-6)
    step = 0
    def once():
        global step
        step += 1
        a.step = step
        _lib.check(L.gsr_adam_step_peer(C.byref(a)), "gsr_adam_step_peer")
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        once()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{variant}: P={P} elements={total} {ms:.3f} ms  {24 * total / ms / 1e6:.0f} GB/s of 24 B/element")
