"""Developer tool: bit-exact / tolerance parity vs the reference build at FULL config sizes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import run_product_backward, run_product_forward, run_reference, scenes, to_dev
from test_gpu_parity import _compare_forward, _compare_grads
for name in sys.argv[1:] or ["cfg1", "cfg2", "cfg4"]:
    scene = getattr(scenes, name)() if name != "cfg1" else scenes.cfg1(sh_coeffs=15)
    t = to_dev(scene)
    ref = run_reference(scene, t)
    fwd = run_product_forward(scene, t)
    info = _compare_forward(scene, t, fwd, ref)
    g = run_product_backward(scene, t, fwd)
    worst = _compare_grads(g, ref["grads"])
    print(name, "P", scene.P, info, {k: f"{v:.1e}" for k, v in worst.items()})
    del ref, fwd, g, t
    torch.cuda.empty_cache()
