"""Measurement for BASELINE config 5 at N = 1: a batch of K keyframes per optimiser step on ONE GPU
(cfg2 scene, 8 ring cameras; P = 500k by default, --P 3000000 for the config's size).  The first keyframe's
backward pass writes the gradient arena, the others add their rendered rows to it
(gsr_backward_args.accumulate), then one masked Adam step.  Prints one JSON line."""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gsdf_slam_b200 import scenes
from gsdf_slam_b200.mapping import Keyframe, KeyframeBatchMapper

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=500_000)
ap.add_argument("--K", type=int, default=8)
ap.add_argument("--steps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda")
scene = scenes.cfg2(a.P)
cams = scenes.ring_cameras(8, scene.cam.W, scene.cam.H, 600.0, 600.0, radius=0.3)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)  # noqa: E731
params = dict(means3D=t(scene.means3D), dc=t(scene.dc), sh=t(scene.sh), opacity=t(scene.opacity),
              scales=t(scene.scales), rotations=t(scene.rotations))
m = KeyframeBatchMapper(params, scene.degree, t(scene.bg), freeze_render_inputs=True)
dL = t(scene.dL_dout)
kfs = [Keyframe(t(c.viewmatrix), t(c.projmatrix), t(c.campos), c.tanfovx, c.tanfovy, c.H, c.W, dL) for c in cams[:a.K]]
for _ in range(5):
    m.step(kfs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    m.step(kfs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(json.dumps({"workload": f"cfg2 scene, P={a.P}, {a.K} keyframes per step on one GPU", "ms_per_step": round(ms, 3),
                  "batches_per_s": round(1000.0 / ms, 1), "keyframes_per_s": round(1000.0 * a.K / ms, 1)}))
