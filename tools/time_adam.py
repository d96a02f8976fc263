import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gsdf_slam_b200 import FusedAdam
P = 500_000
shapes = [(P,3),(P,1,3),(P,15,3),(P,1),(P,3),(P,4)]
ps = [torch.randn(s, device="cuda") for s in shapes]
gs_ = {p: torch.randn_like(p) for p in ps}
opt = FusedAdam([{"params":[p],"lr":1e-3} for p in ps], eps=1e-15)
for _ in range(5): opt.step(grads=gs_)
torch.cuda.synchronize()
ts=[]
for _ in range(30):
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record(); opt.step(grads=gs_); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
n=sum(p.numel() for p in ps)
print(f"adam median {np.median(ts)*1000:.1f} us min {min(ts)*1000:.1f} us -> {28*n/min(ts)/1e6:.0f} GB/s")
