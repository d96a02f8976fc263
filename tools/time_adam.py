"""Developer tool: time the fused Adam step and the two phases of the visibility-split step."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gsdf_slam_b200 import FusedAdam
P = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.155
shapes = [(P,3),(P,1,3),(P,15,3),(P,1),(P,3),(P,4)]
ps = [torch.randn(s, device="cuda") for s in shapes]
gs_ = {p: torch.randn_like(p) for p in ps}
radii = (torch.rand(P, device="cuda") < frac).to(torch.int32)
V = int(radii.sum())
opt = FusedAdam([{"params":[p],"lr":1e-3} for p in ps], eps=1e-15)
n = sum(p.numel() for p in ps)

def timeit(fn, label, nbytes):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"{label:24s} median {np.median(ts)*1000:7.1f} us  min {min(ts)*1000:7.1f} us -> {nbytes/min(ts)/1e6:6.0f} GB/s")

timeit(lambda: opt.step(grads=gs_), "fused", 28 * n)
timeit(lambda: opt.step_rows(radii, 0, new_step=True), "rows zero-grad", 24 * 59 * (P - V))
timeit(lambda: opt.step_rows(radii, 1, grads=gs_), "rows with-grad", 28 * 59 * V)
