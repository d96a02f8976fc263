"""Developer tool: time product vs reference-CUDA fwd/bwd on one scene with CUDA events and
print the workload counts (P, V, R, E ...).  Not part of the bench contract."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from util import (product_exports, run_product_backward, run_product_forward,  # noqa: E402
                  run_reference, scenes, to_dev)


def ev_time(fn, iters, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg2")
    ap.add_argument("--P", type=int, default=0)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--ref", type=int, default=1)
    a = ap.parse_args()
    mk = getattr(scenes, a.cfg)
    scene = mk(a.P) if a.P else mk()
    t = to_dev(scene)
    cam = scene.cam
    print(f"scene {scene.name} P={scene.P} {cam.W}x{cam.H} degree={scene.degree}")

    fwd = run_product_forward(scene, t)
    torch.cuda.synchronize()
    ex = product_exports(scene, t, fwd)
    R = fwd[0]
    V = int((fwd[3] > 0).sum())
    E = int(ex["n_contrib"].long().sum())
    lens = (ex["ranges"][:, 1] - ex["ranges"][:, 0])
    print(f"V={V} R_v={R} T={lens.numel()} N={cam.W * cam.H} sum_n_contrib={E} "
          f"max_tile_len={int(lens.max())} mean_tile_len={float(lens.float().mean()):.1f} "
          f"sum_max_contrib={int(ex['max_contrib'].long().sum())}")

    f_med, f_min = ev_time(lambda: run_product_forward(scene, t), a.iters, 5)
    b_med, b_min = ev_time(lambda: run_product_backward(scene, t, fwd), a.iters, 5)
    print(f"product  fwd {f_med:.3f} ms (min {f_min:.3f})  bwd {b_med:.3f} ms (min {b_min:.3f})  "
          f"fwd+bwd {f_med + b_med:.3f} ms -> {1000.0 / (f_med + b_med):.1f} it/s")
    t0 = time.perf_counter()
    for _ in range(10):
        f2 = run_product_forward(scene, t)
        run_product_backward(scene, t, f2)
    torch.cuda.synchronize()
    print(f"product  wall fwd+bwd {(time.perf_counter() - t0) * 100:.3f} ms")

    if a.ref:
        from oracle import ref_binding as rb
        e = t["empty"]
        rr = rb.RefRasterizer()

        def rf(debug=False):
            return rr.forward(t["bg"], t["means3D"], e, t["opacity"], t["scales"], t["rotations"], 1.0, e,
                              t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W,
                              t["dc"], t["sh"], scene.degree, t["campos"], debug=debug)
        Rr, Br, color, radii = rf()

        def rbk(debug=False):
            return rr.backward(t["bg"], t["means3D"], radii, e, t["scales"], t["rotations"], 1.0, e,
                               t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy, t["dL_dout"],
                               t["dc"], t["sh"], scene.degree, t["campos"], debug=debug)
        print(f"reference R(incl. padding)={Rr} B={Br}")
        for dbg in (False, True):
            f_med, f_min = ev_time(lambda: rf(dbg), a.iters, 5)
            b_med, b_min = ev_time(lambda: rbk(dbg), a.iters, 5)
            print(f"reference(debug={dbg}) fwd {f_med:.3f} ms (min {f_min:.3f})  bwd {b_med:.3f} ms "
                  f"(min {b_min:.3f})  fwd+bwd {f_med + b_med:.3f} ms -> {1000.0 / (f_med + b_med):.1f} it/s")


if __name__ == "__main__":
    main()
