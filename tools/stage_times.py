"""Developer tool: per-stage CUDA-event times (gsr_profile_*) of the product's fwd + bwd on one scene.
    python tools/stage_times.py --cfg cfg2 [--iters 30] [--tag name]
Prints one JSON line; GSR_LIB_PATH selects a kernel variant build (variants/*.so)."""
import argparse, ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import run_product_backward, run_product_forward, scenes, to_dev  # noqa: E402
from gsdf_slam_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--P", type=int, default=0)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--tag", default="")
a = ap.parse_args()
mk = getattr(scenes, a.cfg)
scene = (mk(a.P) if a.P else mk()) if a.cfg != "cfg1" else scenes.cfg1(sh_coeffs=15)
t = to_dev(scene)
L = _lib.lib()
for _ in range(5):
    fwd = run_product_forward(scene, t)
    run_product_backward(scene, t, fwd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    fwd = run_product_forward(scene, t)
    run_product_backward(scene, t, fwd)
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / a.iters
L.gsr_profile_enable(1)
for _ in range(a.iters):
    fwd = run_product_forward(scene, t)
    run_product_backward(scene, t, fwd)
torch.cuda.synchronize()
buf = C.create_string_buffer(8192)
L.gsr_profile_read(buf, 8192)
L.gsr_profile_enable(0)
st = {}
for line in buf.value.decode().splitlines():
    n, c, tot = line.split()
    st[n] = round(1000.0 * float(tot) / max(1, int(c)), 2)
redo = C.c_uint32(0)
if hasattr(L, "gsr_export_redo_count"):
    L.gsr_export_redo_count(C.c_void_p(fwd[6].data_ptr()), scene.cam.W, scene.cam.H, C.byref(redo), None)
st["redo_tiles"] = int(redo.value)
print(json.dumps({"tag": a.tag, "cfg": a.cfg, "lib": os.path.basename(_lib.LIB_PATH), "fwd_bwd_ms": round(total, 4),
                  "R": int(fwd[0]), "stages_us": st}))
