"""Developer tool / measurement of SURVEY.md row f-3: the map-surgery kernels (csrc/densify.cu) against the
reference's LibTorch op sequences (oracle/densify_ref.py, used here only as the timed baseline) at
P = 500k and 3M SH-3 Gaussians.  Prints one JSON line per operation with the algorithmic bytes
(one read of every source row + one write of every destination row, 724 B per row: 59 floats x
(param + 2 moments) + 4 statistics words) and the achieved fraction of the measured HBM peak."""
import json, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from gsdf_slam_b200.gaussian_model import GaussianModel
from oracle import densify_ref as R   # baseline arm only

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6546.9
ROW = 59 * 4 * 3 + 16


def timed(fn, setup, n=7):
    ts = []
    for i in range(n + 2):
        arg = setup()
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(arg); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
        del arg, out
    return float(np.median(ts))


def model_of(st):
    m = GaussianModel(sh_degree=3, device="cuda", capacity=int(st["xyz_"].shape[0] * 1.3) + 4096)
    m.load(st["xyz_"], st["features_dc_"], st["features_rest_"], st["opacity_"], st["scaling_"], st["rotation_"],
           exist_since_iter=st["exist_since_iter_"], exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"],
           xyz_gradient_accum=st["xyz_gradient_accum_"], denom=st["denom_"], max_radii2D=st["max_radii2D_"])
    m._twin(int(m.P * 1.3) + 4096)   # steady state: the twin exists (the reference's allocator is warm too)
    return m


def main():
    out = []
    for P in (500_000, 3_000_000):
        st = R.random_state(P, 15, seed=1, device="cuda")
        # a realistic densification: ~5 % clone, ~5 % split, ~3 % prune
        st["denom_"].fill_(4.0)
        st["xyz_gradient_accum_"] = torch.rand((P, 1), device="cuda") * 0.0009
        noise = torch.randn((2 * P, 3), device="cuda")
        mask = torch.rand(P, device="cuda") < 0.05
        m = model_of(st)

        def prod_prune(_):
            m._cur.rows = P if m._cur.capacity >= P else m._cur.rows
            m.prunePoints(mask); return None
        # prunePoints swaps the twin stores; reload between runs to keep the source identical
        t_prune = timed(lambda mm: mm.prunePoints(mask), lambda: model_of(st), n=3)
        kept = int((~mask).sum())
        t_prune_ref = timed(lambda s: R.prune_points(s, mask), lambda: R.clone_state(st), n=3)
        b = P * ROW + P + kept * ROW
        out.append({"op": "prunePoints", "P": P, "rows_out": kept, "ms": round(t_prune, 4),
                    "libtorch_ms": round(t_prune_ref, 4), "algorithmic_bytes": b,
                    "achieved_gbs": round(b / t_prune / 1e6, 1), "frac_of_hbm_peak": round(b / t_prune / 1e6 / PEAK, 3)})

        res = {}
        def prod_densify(mm):
            res.update(mm.densifyAndPrune(0.0002, 0.005, 4.0, 20, noise=noise))
        t_d = timed(prod_densify, lambda: model_of(st), n=3)
        t_d_ref = timed(lambda s: R.densify_and_prune(s, 0.0002, 0.005, 4.0, 20, 0.01, noise=noise),
                        lambda: R.clone_state(st), n=3)
        b = P * ROW + res["rows"] * ROW
        out.append({"op": "densifyAndPrune", "P": P, **res, "ms": round(t_d, 4), "libtorch_ms": round(t_d_ref, 4),
                    "algorithmic_bytes": b, "achieved_gbs": round(b / t_d / 1e6, 1),
                    "frac_of_hbm_peak": round(b / t_d / 1e6 / PEAK, 3)})

        grad = torch.randn((P, 3), device="cuda") * 1e-3
        radii = (torch.rand(P, device="cuda") * 60 - 40).to(torch.int32)
        m = model_of(st)
        t_s = timed(lambda _: m.addDensificationStats(grad, radii), lambda: None)
        t_s_ref = timed(lambda _: R.add_densification_stats(st, grad, radii), lambda: None)
        V = int((radii > 0).sum())
        b = 4 * P + V * (8 + 24)
        out.append({"op": "addDensificationStats", "P": P, "V": V, "ms": round(t_s, 4), "libtorch_ms": round(t_s_ref, 4),
                    "algorithmic_bytes": b, "achieved_gbs": round(b / t_s / 1e6, 1)})

        n_new = P // 10
        pts = torch.randn((n_new, 3), device="cuda") * 2.0
        cols = torch.rand((n_new, 3), device="cuda")
        import gsdf_slam_b200 as gs
        t_i = timed(lambda mm: mm.increasePcd(pts, cols, 5), lambda: model_of(st), n=3)
        t_i_ref = timed(lambda s: R.increase_pcd(s, pts, cols, 5, 3, gs.distCUDA2), lambda: R.clone_state(st), n=3)
        out.append({"op": "increasePcd", "P": P, "new_points": n_new, "ms": round(t_i, 4),
                    "libtorch_ms": round(t_i_ref, 4)})
        del m, st, noise
        torch.cuda.empty_cache()
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
