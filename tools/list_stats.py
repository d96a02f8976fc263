"""Developer tool: per-tile list lengths vs the deepest contributor (max_contrib) of a scene — how much of each
depth-sorted list the blend kernels ever read.   python tools/list_stats.py cfg5"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import product_exports, run_product_forward, scenes, to_dev  # noqa: E402

for name in sys.argv[1:] or ["cfg5"]:
    sc = getattr(scenes, name)() if name != "cfg1" else scenes.cfg1(sh_coeffs=15)
    t = to_dev(sc)
    fwd = run_product_forward(sc, t)
    ex = product_exports(sc, t, fwd)
    r = ex["ranges"].long()
    ln = (r[:, 1] - r[:, 0]).cpu().numpy()
    mc = ex["max_contrib"].long().cpu().numpy()
    nz = ln > 0
    print(name, "tiles", len(ln), "R", int(ln.sum()), "sum max_contrib", int(mc.sum()),
          "len pcts", np.percentile(ln[nz], [10, 50, 90, 99, 100]).tolist(),
          "max_contrib pcts", np.percentile(mc[nz], [10, 50, 90, 99, 100]).tolist(),
          "tiles with max_contrib <= 128 / 256 / 512:", float((mc[nz] <= 128).mean()), float((mc[nz] <= 256).mean()),
          float((mc[nz] <= 512).mean()), "ratio pcts", np.percentile(mc[nz] / ln[nz], [50, 90, 99]).round(3).tolist())
