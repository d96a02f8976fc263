"""Developer tool: where the side-stream Adam phase actually runs relative to the main stream."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gsdf_slam_b200 import scenes
from gsdf_slam_b200.mapping import Keyframe, KeyframeBatchMapper
dev = torch.device("cuda", 0)
scene = scenes.cfg2(); cam = scene.cam
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
params = dict(means3D=t(scene.means3D), dc=t(scene.dc), sh=t(scene.sh), opacity=t(scene.opacity),
              scales=t(scene.scales), rotations=t(scene.rotations))
m = KeyframeBatchMapper(params, scene.degree, t(scene.bg), freeze_render_inputs=True)
kf = Keyframe(t(cam.viewmatrix), t(cam.projmatrix), t(cam.campos), cam.tanfovx, cam.tanfovy, cam.H, cam.W,
              t(scene.dL_dout))
for _ in range(5): m.step([kf])
torch.cuda.synchronize()
for it in range(3):
    m.trace = []
    m.step([kf]); m.step([kf])
    torch.cuda.synchronize()
    t0 = m.trace[0][1]
    print(" | ".join(f"{lab} {t0.elapsed_time(ev)*1000:.0f}" for lab, ev in m.trace))
