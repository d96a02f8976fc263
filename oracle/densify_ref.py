"""TEST INFRASTRUCTURE — oracle for SURVEY.md row f-3 (densify / prune / increasePcd).

A restatement, op for op, of the LibTorch sequences in /root/reference/src/gaussian_model.cpp with the
same ATen operators (torch's Python API dispatches to the kernels LibTorch's C++ API does), each
function citing the lines it follows.  It runs on CPU tensors (the `not gpu` tests pin it against
hand-derived expectations, tests/test_densify_oracle.py) and on CUDA tensors (the `gpu` parity
tests compare the product's kernels with it on identical inputs).  Only tests/ may import this
module; nothing under gsdf_slam_b200/ does.

Parity status: PINNED.  The reference has no tests or golden vectors for these functions, but its own code
is compiled here: oracle/build_ref_densify.py extracts the bodies of GaussianModel::{prunePoints,
densificationPostfix, densifyAndSplit, densifyAndClone, densifyAndPrune, resetOpacity, addDensificationStats, ...}
from /root/reference/src/gaussian_model.cpp at build time and compiles them (oracle/_ref/gsdf_ref_densify.so,
`ref_module()` / `reference_*` below).  tests/test_densify_oracle.py checks this restatement against that module on
CPU tensors, tests/test_gpu_densify.py checks the product's kernels against it on the GPU.  (increasePcd /
createFromPcd need the whole GaussianModel class and stay restated.)

The model state is a dict of tensors with the reference's member names and shapes:
  xyz_ [P,3], features_dc_ [P,1,3], features_rest_ [P,M,3], opacity_ [P,1], scaling_ [P,3],
  rotation_ [P,4]; exp_avg / exp_avg_sq: lists of six tensors shaped like the parameters (the
  AdamParamState of each group); exist_since_iter_ [P] i32, xyz_gradient_accum_ [P,1], denom_ [P,1],
  max_radii2D_ [P].
"""
from __future__ import annotations

import importlib.util
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "gsdf_ref_densify.so")
_mod = None


def ref_module():
    """The reference's own densification code, compiled (None when oracle/_ref/gsdf_ref_densify.so is absent)."""
    global _mod
    if _mod is None and os.path.exists(_SO):
        spec = importlib.util.spec_from_file_location("gsdf_ref_densify", _SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def _flat(st):
    return ([st[n] for n in PARAMS], list(st["exp_avg"]), list(st["exp_avg_sq"]), st["exist_since_iter_"],
            st["xyz_gradient_accum_"], st["denom_"], st["max_radii2D_"])


def _unflat(out):
    st = {n: out[g] for g, n in enumerate(PARAMS)}
    st["exp_avg"] = list(out[6:12])
    st["exp_avg_sq"] = list(out[12:18])
    st["exist_since_iter_"], st["xyz_gradient_accum_"], st["denom_"], st["max_radii2D_"] = out[18:22]
    return st


def reference_densify_and_prune(st, max_grad, min_opacity, extent, max_screen_size, percent_dense, seed, step=7):
    """The compiled reference's densifyAndPrune on a copy of `st`.  Returns (new_state, noise): `noise` is the
    standard-normal tensor its at::normal(means, stds) call drew (at::normal(Tensor, Tensor) fills the output with
    N(0,1) from the default generator, then multiplies by stds and adds means), reproduced by re-seeding."""
    m = ref_module()
    dev = st["xyz_"].device
    torch.manual_seed(seed)
    out = m.densify_and_prune(*_flat(st), float(max_grad), float(min_opacity), float(extent), int(max_screen_size),
                              float(percent_dense), int(step))
    # number of rows densifyAndSplit selected == rows at::normal drew for (N = 2 copies each)
    grads = st["xyz_gradient_accum_"] / st["denom_"]
    grads[grads.isnan()] = 0.0
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))  # noqa: E731
    n_clone = int((torch.logical_and(torch.linalg.vector_norm(grads, dim=-1) >= f32(max_grad),
                                     torch.exp(st["scaling_"]).max(dim=1).values <= f32(percent_dense) * f32(extent))).sum())
    padded = torch.zeros(st["xyz_"].shape[0] + n_clone, device=dev)
    padded[: grads.shape[0]] = grads.squeeze()
    scal = torch.cat([st["scaling_"], st["scaling_"][:0]])   # clones are appended: recompute on the grown state below
    st2, _ = densify_and_clone(clone_state(st), grads, f32(max_grad), f32(extent), f32(percent_dense))
    sel = torch.logical_and(padded >= f32(max_grad),
                            torch.exp(st2["scaling_"]).max(dim=1).values > f32(percent_dense) * f32(extent))
    rows = 2 * int(sel.sum())
    torch.manual_seed(seed)
    noise = torch.normal(torch.zeros((rows, 3), device=dev), torch.ones((rows, 3), device=dev))
    del scal
    return _unflat(out), noise


def reference_densify_and_clone(st, max_grad, extent, percent_dense, step=7):
    """densifyAndClone of the compiled reference (runs on CPU too; densifyAndSplit needs CUDA: the reference's
    build_rotation allocates on torch::kCUDA, include/general_utils.h:46)."""
    return _unflat(ref_module().densify_and_clone(*_flat(st), float(max_grad), float(extent), float(percent_dense),
                                                  int(step)))


def reference_prune_points(st, mask, step=7):
    return _unflat(ref_module().prune_points(*_flat(st), mask, int(step)))


def reference_reset_opacity(st, step=7):
    return _unflat(ref_module().reset_opacity(*_flat(st), int(step)))


def reference_add_densification_stats(st, viewspace_grad, update_filter):
    """GaussianModel::addDensificationStats alone (the max_radii2D line lives in the mapper)."""
    return ref_module().add_densification_stats(st["xyz_gradient_accum_"], st["denom_"], viewspace_grad, update_filter)

PARAMS = ("xyz_", "features_dc_", "features_rest_", "opacity_", "scaling_", "rotation_")
C0 = 0.28209479177387814  # include/sh_utils.h:33 (a float constant there)


def clone_state(st):
    out = {}
    for k, v in st.items():
        out[k] = [t.clone() for t in v] if isinstance(v, list) else v.clone()
    return out


def prune_points(st, mask):
    """GaussianModel::prunePoints, src/gaussian_model.cpp:431-490."""
    valid = ~mask
    out = {}
    for g, name in enumerate(PARAMS):
        out[name] = st[name][valid]
    out["exp_avg"] = [m[valid].clone() for m in st["exp_avg"]]
    out["exp_avg_sq"] = [m[valid].clone() for m in st["exp_avg_sq"]]
    out["exist_since_iter_"] = st["exist_since_iter_"][valid]
    out["xyz_gradient_accum_"] = st["xyz_gradient_accum_"][valid]
    out["denom_"] = st["denom_"][valid]
    out["max_radii2D_"] = st["max_radii2D_"][valid]
    return out


def densification_postfix(st, new_xyz, new_features_dc, new_features_rest, new_opacities, new_scaling,
                          new_rotation, new_exist_since_iter):
    """GaussianModel::densificationPostfix, src/gaussian_model.cpp:492-569."""
    new = (new_xyz, new_features_dc, new_features_rest, new_opacities, new_scaling, new_rotation)
    out = {}
    out["exp_avg"], out["exp_avg_sq"] = [], []
    for g, name in enumerate(PARAMS):
        ext = new[g]
        out["exp_avg"].append(torch.cat([st["exp_avg"][g].clone(), torch.zeros_like(ext)], dim=0))
        out["exp_avg_sq"].append(torch.cat([st["exp_avg_sq"][g].clone(), torch.zeros_like(ext)], dim=0))
        out[name] = torch.cat([st[name], ext], dim=0)
    out["exist_since_iter_"] = torch.cat([st["exist_since_iter_"], new_exist_since_iter], dim=0)
    n = out["xyz_"].shape[0]
    dev = out["xyz_"].device
    out["xyz_gradient_accum_"] = torch.zeros((n, 1), device=dev)
    out["denom_"] = torch.zeros((n, 1), device=dev)
    out["max_radii2D_"] = torch.zeros((n,), device=dev)
    return out


def build_rotation(r):
    """general_utils::build_rotation, include/general_utils.h:32-57."""
    r0, r1, r2, r3 = r[:, 0], r[:, 1], r[:, 2], r[:, 3]
    norm = torch.sqrt(r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3)
    q = r / norm.unsqueeze(1)
    rr, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.zeros((q.shape[0], 3, 3), device=r.device)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - rr * z)
    R[:, 0, 2] = 2 * (x * z + rr * y)
    R[:, 1, 0] = 2 * (x * y + rr * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - rr * x)
    R[:, 2, 0] = 2 * (x * z - rr * y)
    R[:, 2, 1] = 2 * (y * z + rr * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def densify_and_split(st, grads, grad_threshold, scene_extent, percent_dense, N=2, noise=None):
    """GaussianModel::densifyAndSplit, src/gaussian_model.cpp:571-612.  `noise`: the standard-normal
    draws at::normal(means, stds) would make ([N*S,3]; at::normal computes z * std + mean)."""
    n_init = st["xyz_"].shape[0]
    dev = st["xyz_"].device
    padded_grad = torch.zeros((n_init,), device=dev)
    padded_grad[: grads.shape[0]] = grads.squeeze()
    sel = torch.where(padded_grad >= grad_threshold, True, False)
    scaling_act = torch.exp(st["scaling_"])
    sel = torch.logical_and(sel, scaling_act.max(dim=1).values > percent_dense * scene_extent)
    stds = scaling_act[sel].repeat(N, 1)
    means = torch.zeros((stds.shape[0], 3), device=dev)
    if noise is None:
        samples = torch.normal(means, stds)
    else:
        samples = noise[: stds.shape[0]].to(dev).clone().mul_(stds).add_(means)
    r_masked = st["rotation_"][sel]
    rots = build_rotation(r_masked).repeat(N, 1, 1)
    new_xyz = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + st["xyz_"][sel].repeat(N, 1)
    new_scaling = torch.log(scaling_act[sel].repeat(N, 1) / (0.8 * N))
    new_rotation = st["rotation_"][sel].repeat(N, 1)
    new_features_dc = st["features_dc_"][sel].repeat(N, 1, 1)
    new_features_rest = st["features_rest_"][sel].repeat(N, 1, 1)
    new_opacity = st["opacity_"][sel].repeat(N, 1)
    new_exist = st["exist_since_iter_"][sel].repeat(N)
    st = densification_postfix(st, new_xyz, new_features_dc, new_features_rest, new_opacity, new_scaling,
                               new_rotation, new_exist)
    prune_filter = torch.cat([sel, torch.zeros((N * int(sel.sum().item()),), device=dev, dtype=torch.bool)])
    return prune_points(st, prune_filter), int(sel.sum().item())


def densify_and_clone(st, grads, grad_threshold, scene_extent, percent_dense):
    """GaussianModel::densifyAndClone, src/gaussian_model.cpp:614-642."""
    sel = torch.where(torch.linalg.vector_norm(grads, dim=-1) >= grad_threshold, True, False)
    scaling_act = torch.exp(st["scaling_"])
    sel = torch.logical_and(sel, scaling_act.max(dim=1).values <= percent_dense * scene_extent)
    new = [st[name][sel] for name in PARAMS]
    return densification_postfix(st, *new, st["exist_since_iter_"][sel]), int(sel.sum().item())


def densify_and_prune(st, max_grad, min_opacity, extent, max_screen_size, percent_dense, noise=None, N=2):
    """GaussianModel::densifyAndPrune, src/gaussian_model.cpp:644-665.  Returns (state, counts)."""
    grads = st["xyz_gradient_accum_"] / st["denom_"]
    grads[grads.isnan()] = 0.0
    # float products as the C++ forms them: percentDense() * scene_extent and 0.1f * extent are floats
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))  # noqa: E731
    st, n_clone = densify_and_clone(st, grads, f32(max_grad), f32(extent), f32(percent_dense))
    st, n_split = densify_and_split(st, grads, f32(max_grad), f32(extent), f32(percent_dense), N, noise)
    prune_mask = (torch.sigmoid(st["opacity_"]) < f32(min_opacity)).squeeze(-1)
    if max_screen_size:
        big_vs = st["max_radii2D_"] > max_screen_size
        big_ws = torch.exp(st["scaling_"]).max(dim=1).values > float(
            torch.tensor(0.1, dtype=torch.float32) * torch.tensor(extent, dtype=torch.float32))
        prune_mask = torch.logical_or(torch.logical_or(prune_mask, big_vs), big_ws)
    return prune_points(st, prune_mask), {"cloned": n_clone, "split": n_split}


def add_densification_stats(st, viewspace_grad, radii):
    """src/gaussian_mapper.cpp:443-447 + GaussianModel::addDensificationStats, src/gaussian_model.cpp:667-681.
    In place on st."""
    vis = radii > 0
    st["max_radii2D_"][vis] = torch.max(st["max_radii2D_"][vis], radii[vis])
    st["xyz_gradient_accum_"].index_put_(
        (vis,), torch.linalg.vector_norm(viewspace_grad[vis, :2], dim=-1, keepdim=True), accumulate=True)
    st["denom_"].index_put_((vis,), st["denom_"][vis] + 1)
    return st


def reset_opacity(st):
    """GaussianModel::resetOpacity, src/gaussian_model.cpp:399-408 (the cap is ones_like(x * 0.01) = 1)."""
    act = torch.sigmoid(st["opacity_"])
    x = torch.min(act, torch.ones_like(act * 0.01))
    st["opacity_"] = torch.log(x / (1 - x))
    st["exp_avg"][3] = torch.zeros_like(st["opacity_"])
    st["exp_avg_sq"][3] = torch.zeros_like(st["opacity_"])
    return st


def voxel_keep_indices(existing_xyz, new_points, voxel_size=0.05):
    """The duplicate test of increasePcd, src/gaussian_model.cpp:214-233."""
    vs = float(torch.tensor(voxel_size, dtype=torch.float32))

    def keys(p):
        v = torch.floor(p / vs).to(torch.int32)
        return (v[:, 0].to(torch.int64) * 73856093 + v[:, 1].to(torch.int64) * 19349663
                + v[:, 2].to(torch.int64) * 83492791)

    isin = torch.isin(keys(new_points), keys(existing_xyz))
    return (~isin).nonzero().squeeze(-1)


def increase_pcd(st, new_points, new_colors, iteration, max_sh_degree, dist_fn, voxel_size=0.05):
    """GaussianModel::increasePcd, src/gaussian_model.cpp:207-317.  dist_fn = distCUDA2."""
    if new_points.shape[0] == 0:
        return st, 0
    dev = st["xyz_"].device
    keep = voxel_keep_indices(st["xyz_"], new_points, voxel_size)
    pts = new_points.index_select(0, keep)
    cols = new_colors.index_select(0, keep)
    fused = (cols - 0.5) / float(torch.tensor(C0, dtype=torch.float32))
    k = (max_sh_degree + 1) ** 2
    features = torch.zeros((fused.shape[0], 3, k), dtype=torch.float32, device=dev)
    features[:, :3, 0] = fused
    features[:, 3:, 1:] = 0.0
    dist2 = torch.clamp_min(dist_fn(pts.clone()), 0.0000001)
    scales = torch.log(torch.sqrt(dist2)).unsqueeze(1).repeat(1, 3)
    rots = torch.zeros((pts.shape[0], 4), device=dev)
    rots[:, 0] = 1
    x = 0.1 * torch.ones((pts.shape[0], 1), dtype=torch.float32, device=dev)
    opac = torch.log(x / (1 - x))
    exist = torch.full((pts.shape[0],), iteration, dtype=torch.int32, device=dev)
    f_dc = features[:, :, 0:1].transpose(1, 2).contiguous()
    f_rest = features[:, :, 1:].transpose(1, 2).contiguous()
    return densification_postfix(st, pts, f_dc, f_rest, opac, scales, rots, exist), int(pts.shape[0])


def create_from_pcd(point_cloud, color, max_sh_degree, dist_fn):
    """GaussianModel::createFromPcd, src/gaussian_model.cpp:130-199 (the voxel-unique selection at :135-146 is
    computed and then discarded by the reference: index_select's result is not assigned)."""
    dev = point_cloud.device
    fused = (color - 0.5) / float(torch.tensor(C0, dtype=torch.float32))
    k = (max_sh_degree + 1) ** 2
    n = point_cloud.shape[0]
    features = torch.zeros((n, 3, k), dtype=torch.float32, device=dev)
    features[:, :3, 0] = fused
    dist2 = torch.clamp_min(dist_fn(point_cloud.clone()), 0.0000001)
    scales = torch.log(torch.sqrt(dist2)).unsqueeze(1).repeat(1, 3)
    rots = torch.zeros((n, 4), device=dev)
    rots[:, 0] = 1
    x = 0.1 * torch.ones((n, 1), dtype=torch.float32, device=dev)
    st = {"xyz_": point_cloud, "features_dc_": features[:, :, 0:1].transpose(1, 2).contiguous(),
          "features_rest_": features[:, :, 1:].transpose(1, 2).contiguous(), "opacity_": torch.log(x / (1 - x)),
          "scaling_": scales, "rotation_": rots,
          "exist_since_iter_": torch.zeros((n,), dtype=torch.int32, device=dev),
          "max_radii2D_": torch.zeros((n,), device=dev)}
    return st


def random_state(P, M, seed, device="cpu", with_moments=True):
    """A seeded model state whose densification decisions are all exercised."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    st = {
        "xyz_": r(P, 3) * 2.0,
        "features_dc_": r(P, 1, 3),
        "features_rest_": r(P, M, 3) * 0.1,
        "opacity_": r(P, 1) * 2.5,
        "scaling_": torch.log(torch.tensor(0.03)) + r(P, 3) * 0.9,
        "rotation_": r(P, 4),
        "exist_since_iter_": torch.randint(0, 500, (P,), generator=g, dtype=torch.int32),
        "xyz_gradient_accum_": torch.rand(P, 1, generator=g) * 0.004,
        "denom_": torch.randint(0, 6, (P, 1), generator=g).to(torch.float32),
        "max_radii2D_": torch.rand(P, generator=g) * 40.0,
    }
    shapes = [st[n].shape for n in PARAMS]
    if with_moments:
        st["exp_avg"] = [r(*s) * 1e-3 for s in shapes]
        st["exp_avg_sq"] = [torch.rand(*s, generator=g) * 1e-6 for s in shapes]
    else:
        st["exp_avg"] = [torch.zeros(*s) for s in shapes]
        st["exp_avg_sq"] = [torch.zeros(*s) for s in shapes]
    return {k: ([t.to(device) for t in v] if isinstance(v, list) else v.to(device)) for k, v in st.items()}
