"""Measurement of BASELINE config 3 (TUM-style 640x480 mapping loop with densify / prune, the Gaussian count
changes every 10 iterations): the whole iteration of GaussianMapper::trainForOneIteration
(src/gaussian_mapper.cpp:341-466) on both arms, same schedule, same inputs.

  this repo : raw parameters -> gsr_forward (activations fused) -> gsr_l1_ssim_loss -> gsr_backward ->
              gsr_add_densification_stats -> gsr_adam_step_rows on the store's moments;
              every 10th iteration gsr_densify_and_prune instead of the optimiser step
  reference : LibTorch activations + autograd, the UNMODIFIED reference rasterizer (oracle/_ref), the LibTorch
              L1 + SSIM op sequence + autograd, index_put_ statistics, torch::optim::Adam's op sequence,
              densifyAndPrune as the LibTorch op sequence (oracle/densify_ref.py) + emptyCache

Prints one JSON line.  The oracle modules are used here only as the timed baseline arm."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import gsdf_slam_b200 as gs
from gsdf_slam_b200 import scenes
from gsdf_slam_b200.gaussian_model import GaussianModel
from gsdf_slam_b200.loss import l1_ssim_loss
from oracle.loss_ref import reference_loss
from oracle import densify_ref as R
from oracle import ref_binding as rb

ITERS, WARM, EVERY = 120, 20, 10
LR = dict(xyz=1.6e-4, f_dc=0.0025, f_rest=0.000125, opacity=0.05, scaling=0.005, rotation=0.001)
dev = torch.device("cuda")
scene = scenes.frustum_scene(100_000, 640, 480, 517.306, 516.469, seed=3, degree=3, sh_coeffs=15, name="cfg3")
cam = scene.cam
t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in scene.arrays().items()}
e = torch.empty(0, device=dev)
target = gs.RasterizeGaussiansCUDA(t["bg"], t["means3D"], e, t["opacity"], t["scales"], t["rotations"], 1.0, e,
                                   t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W, t["dc"],
                                   t["sh"], 3, t["campos"], False, False)[2].clone()
op = t["opacity"].clamp(1e-4, 1 - 1e-4)
raw = dict(xyz=t["means3D"], f_dc=t["dc"].view(-1, 1, 3) + 0.2, f_rest=t["sh"], opacity=torch.log(op / (1 - op)),
           scaling=torch.log(t["scales"]) + 0.1, rotation=t["rotations"])
P0 = raw["xyz"].shape[0]


def product():
    m = GaussianModel(sh_degree=3, device=dev, capacity=int(P0 * 1.5))
    m.load(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"])
    m.lr_ = dict(LR)
    sizes = []

    def it(i):
        f = gs.RasterizeGaussiansCUDA(t["bg"], m.xyz_, e, m.opacity_, m.scaling_, m.rotation_, 1.0, e, t["viewmatrix"],
                                      t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W, m.features_dc_.view(-1, 3),
                                      m.features_rest_, 3, t["campos"], False, False, raw_params=True)
        stats, dL = l1_ssim_loss(f[2], target, None, 0.2)
        g = gs.RasterizeGaussiansBackwardCUDA(t["bg"], m.xyz_, f[3], e, m.scaling_, m.rotation_, 1.0, e, t["viewmatrix"],
                                              t["projmatrix"], cam.tanfovx, cam.tanfovy, dL, m.features_dc_.view(-1, 3),
                                              m.features_rest_, 3, t["campos"], f[4], f[0], f[5], f[6], f[1], f[7], False,
                                              raw_opacity=m.opacity_)
        m.addDensificationStats(g[0], f[3])
        if i % EVERY == 0:
            m.densifyAndPrune(0.0002, 0.005, 6.0, 0)
            sizes.append(m.P)
        else:
            m.step([g[3], g[5], g[6], g[2], g[7], g[8]], row_mask=f[3])
    return it, sizes, lambda: m.P


def reference():
    st = dict(xyz_=raw["xyz"].clone(), features_dc_=raw["f_dc"].clone(), features_rest_=raw["f_rest"].clone(),
              opacity_=raw["opacity"].clone(), scaling_=raw["scaling"].clone(), rotation_=raw["rotation"].clone())
    st["exp_avg"] = [torch.zeros_like(st[n]) for n in R.PARAMS]
    st["exp_avg_sq"] = [torch.zeros_like(st[n]) for n in R.PARAMS]
    st["exist_since_iter_"] = torch.zeros(P0, dtype=torch.int32, device=dev)
    st["xyz_gradient_accum_"] = torch.zeros((P0, 1), device=dev)
    st["denom_"] = torch.zeros((P0, 1), device=dev)
    st["max_radii2D_"] = torch.zeros(P0, device=dev)
    box = {"st": st, "step": 0}
    rr = rb.RefRasterizer()
    sizes = []
    lrs = [LR["xyz"], LR["f_dc"], LR["f_rest"], LR["opacity"], LR["scaling"], LR["rotation"]]

    def it(i):
        st = box["st"]
        ro, rs, rq = (st[n].detach().requires_grad_(True) for n in ("opacity_", "scaling_", "rotation_"))
        op_a, sc_a, rot_a = torch.sigmoid(ro), torch.exp(rs), torch.nn.functional.normalize(rq)   # gaussian_model.cpp:48-71
        dc, sh = st["features_dc_"].view(-1, 3), st["features_rest_"]
        Rn, B, color, radii = rr.forward(t["bg"], st["xyz_"], e, op_a.detach(), sc_a.detach(), rot_a.detach(), 1.0, e,
                                         t["viewmatrix"], t["projmatrix"], cam.tanfovx, cam.tanfovy, cam.H, cam.W, dc, sh,
                                         3, t["campos"], debug=False)
        x = color.detach().requires_grad_(True)
        reference_loss(x, target, None, 0.2).backward()
        g = rr.backward(t["bg"], st["xyz_"], radii, e, sc_a.detach(), rot_a.detach(), 1.0, e, t["viewmatrix"],
                        t["projmatrix"], cam.tanfovx, cam.tanfovy, x.grad, dc, sh, 3, t["campos"], debug=False)
        torch.autograd.backward([op_a, sc_a, rot_a], [g["dL_dopacity"].view_as(op_a), g["dL_dscales"], g["dL_drotations"]])
        R.add_densification_stats(st, g["dL_dmeans2D"], radii)
        if i % EVERY == 0:
            box["st"], _ = R.densify_and_prune(st, 0.0002, 0.005, 6.0, 0, 0.01)
            torch.cuda.empty_cache()                                   # gaussian_model.cpp:664
            sizes.append(box["st"]["xyz_"].shape[0])
        else:
            box["step"] += 1
            s = box["step"]
            bc1, bc2 = 1 - 0.9 ** s, 1 - 0.999 ** s
            grads = [g["dL_dmeans3D"], g["dL_ddc"], g["dL_dsh"], ro.grad, rs.grad, rq.grad]
            for k, n in enumerate(R.PARAMS):                           # torch::optim::Adam, single-tensor path
                gk = grads[k].view_as(st[n])
                mk, vk = st["exp_avg"][k], st["exp_avg_sq"][k]
                mk.mul_(0.9).add_(gk, alpha=0.1)
                vk.mul_(0.999).addcmul_(gk, gk, value=0.001)
                denom = (vk.sqrt() / np.sqrt(bc2)).add_(1e-15)
                st[n].addcdiv_(mk, denom, value=-(lrs[k] / bc1))
    return it, sizes, lambda: box["st"]["xyz_"].shape[0]


def run(make):
    it, sizes, count = make()
    for i in range(1, WARM + 1):
        it(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(WARM + 1, WARM + ITERS + 1):
        it(i)
    b.record()
    torch.cuda.synchronize()
    return {"it_per_s": round(ITERS / (a.elapsed_time(b) / 1e3), 1), "wall_it_per_s": round(ITERS / (time.perf_counter() - t0), 1),
            "P_final": count(), "P_after_each_densify": sizes[-6:]}


out = {"workload": "cfg3: 640x480 mapping loop, P0=100k SH-3, densifyAndPrune every 10 iterations", "iters": ITERS,
       "product": run(product)}
if rb.available():
    out["reference"] = run(reference)
    out["speedup"] = round(out["product"]["it_per_s"] / out["reference"]["it_per_s"], 2)
print(json.dumps(out))
