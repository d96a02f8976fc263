import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import gsdf_slam_b200 as gs
from oracle import ref_binding as rb
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for P in (100_000, 1_000_000):
    g = torch.Generator(device="cuda").manual_seed(P)
    pts = torch.randn((P,3), device="cuda", generator=g) * torch.tensor([4.0,1.0,3.0], device="cuda")
    mine = t(lambda: gs.distCUDA2(pts))
    ref = t(lambda: rb.ref_knn(pts), 3) if rb.available() else float('nan')
    same = bool(torch.equal(gs.distCUDA2(pts), rb.ref_knn(pts))) if rb.available() else None
    print(f"knn P={P}: product {mine:.3f} ms, reference {ref:.3f} ms, bit-equal {same}")
