import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import run_product_backward, run_product_forward, scenes, to_dev
scene = scenes.cfg2(); t = to_dev(scene)
for _ in range(5): f = run_product_forward(scene, t)
torch.cuda.synchronize()
ev=[]; wall=[]
for i in range(20):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); a.record(); f = run_product_forward(scene, t); b.record(); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    ev.append(a.elapsed_time(b)); wall.append(((t1-t0)*1e3,(t2-t0)*1e3))
print("event ms:", " ".join(f"{x:.2f}" for x in ev))
print("wall call/total ms:", " ".join(f"{x:.2f}/{y:.2f}" for x,y in wall))
print("mem allocated MB", torch.cuda.memory_allocated()/1e6, "reserved", torch.cuda.memory_reserved()/1e6)
# keep result alive vs not
ev=[]
keep=[]
for i in range(20):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); f2 = run_product_forward(scene, t); b.record(); torch.cuda.synchronize(); keep=[f2]
    ev.append(a.elapsed_time(b))
print("event ms (keeping prev):", " ".join(f"{x:.2f}" for x in ev))
