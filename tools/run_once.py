"""Developer tool: run product fwd+bwd a few times on one scene (used under ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import run_product_backward, run_product_forward, scenes, to_dev
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
scene = getattr(scenes, cfg)()
t = to_dev(scene)
for _ in range(iters):
    fwd = run_product_forward(scene, t)
    g = run_product_backward(scene, t, fwd)
torch.cuda.synchronize()
print("done", fwd[0])
