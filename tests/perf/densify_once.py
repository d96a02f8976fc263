"""ncu target: one prunePoints + one densifyAndPrune at P rows (SH-3), steady state."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "perf"))
from oracle import densify_ref as R
import time_densify as T
P = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
st = R.random_state(P, 15, seed=1, device="cuda")
st["denom_"].fill_(4.0); st["xyz_gradient_accum_"] = torch.rand((P, 1), device="cuda") * 0.0009
noise = torch.randn((2 * P, 3), device="cuda"); mask = torch.rand(P, device="cuda") < 0.05
m = T.model_of(st); m.prunePoints(mask)
m = T.model_of(st); print(m.densifyAndPrune(0.0002, 0.005, 4.0, 20, noise=noise)); torch.cuda.synchronize()
