"""CPU checks of the row f-3 oracle (oracle/densify_ref.py) against hand-derived expectations, and of the
host-side mirror's argument handling.  The reference has no tests for these functions, so
the expectations below are derived by hand from src/gaussian_model.cpp:431-665 and :207-317."""
import math

import pytest

import torch

from oracle import densify_ref as R
from oracle import densify_ref as dr


def _tiny():
    P, M = 5, 1
    st = R.random_state(P, M, seed=7)
    # row: 0 keep, 1 clone, 2 split, 3 pruned (opacity), 4 clone-selected but transparent
    st["xyz_gradient_accum_"] = torch.tensor([[0.0001], [0.01], [0.02], [0.0], [0.03]])
    st["denom_"] = torch.tensor([[1.0], [2.0], [2.0], [0.0], [3.0]])   # 0/0 = NaN -> 0 for row 3
    small, large = math.log(0.005), math.log(0.5)
    st["scaling_"] = torch.tensor([[small] * 3, [small] * 3, [small, large, small], [small] * 3, [small] * 3])
    st["opacity_"] = torch.tensor([[2.0], [1.0], [0.5], [-9.0], [-9.0]])
    st["rotation_"] = torch.tensor([[1.0, 0, 0, 0]] * P)
    return st


def test_densify_and_prune_row_order_and_values():
    st = _tiny()
    ref = R.clone_state(st)
    noise = torch.zeros((2 * 5, 3))
    out, counts = R.densify_and_prune(st, max_grad=0.0002, min_opacity=0.005, extent=4.0, max_screen_size=0,
                                      percent_dense=0.01, noise=noise)
    assert counts == {"cloned": 2, "split": 1}
    # survivors: row 0, row 1, clone of 1, two copies of row 2 (rows 3, 4 and 4's clone are transparent)
    assert out["xyz_"].shape[0] == 5
    assert torch.equal(out["xyz_"][:3], ref["xyz_"][[0, 1, 1]])
    assert torch.equal(out["xyz_"][3:], ref["xyz_"][[2, 2]])          # zero noise: copies sit on the parent
    exp_scale = torch.log(torch.exp(ref["scaling_"][2]) / 1.6)
    assert torch.equal(out["scaling_"][3], exp_scale) and torch.equal(out["scaling_"][4], exp_scale)
    assert torch.equal(out["features_rest_"][2], ref["features_rest_"][1])
    assert torch.equal(out["exist_since_iter_"], ref["exist_since_iter_"][[0, 1, 1, 2, 2]])
    # Adam moments: kept rows keep theirs, new rows start at zero
    for g in range(6):
        assert torch.equal(out["exp_avg"][g][:2], ref["exp_avg"][g][:2])
        assert float(out["exp_avg"][g][2:].abs().sum()) == 0.0
        assert float(out["exp_avg_sq"][g][2:].abs().sum()) == 0.0
    # densificationPostfix resets the statistics of every row
    for k in ("xyz_gradient_accum_", "denom_", "max_radii2D_"):
        assert out[k].shape[0] == 5 and float(out[k].abs().sum()) == 0.0


def test_size_test_prunes_world_space_giants_only():
    st = _tiny()
    st["max_radii2D_"] = torch.full((5,), 500.0)   # ignored: the statistics were zeroed by the postfix
    out, _ = R.densify_and_prune(st, 0.0002, 0.005, extent=4.0, max_screen_size=20, percent_dense=0.01,
                                 noise=torch.zeros((10, 3)))
    # 0.1 * extent = 0.4: the split copies have max scale 0.5 / 1.6 = 0.3125 and survive
    assert out["xyz_"].shape[0] == 5
    st = _tiny()
    out, _ = R.densify_and_prune(st, 0.0002, 0.005, extent=3.0, max_screen_size=20, percent_dense=0.01,
                                 noise=torch.zeros((10, 3)))
    assert out["xyz_"].shape[0] == 3                # 0.3125 > 0.3: both copies go


def test_prune_points_keeps_order_and_state():
    st = R.random_state(9, 2, seed=1)
    mask = torch.tensor([0, 1, 0, 0, 1, 1, 0, 0, 1], dtype=torch.bool)
    out = R.prune_points(R.clone_state(st), mask)
    keep = [0, 2, 3, 6, 7]
    for name in R.PARAMS + ("exist_since_iter_", "xyz_gradient_accum_", "denom_", "max_radii2D_"):
        assert torch.equal(out[name], st[name][keep])
    assert torch.equal(out["exp_avg_sq"][2], st["exp_avg_sq"][2][keep])


def test_voxel_keep_indices():
    existing = torch.tensor([[0.01, 0.01, 0.01], [1.0, 1.0, 1.0]])
    new = torch.tensor([[0.04, 0.04, 0.04],      # same voxel as existing[0]
                        [0.06, 0.01, 0.01],      # neighbour voxel
                        [1.02, 1.01, 1.04],      # same voxel as existing[1]
                        [-0.01, 0.01, 0.01],     # floor(-0.2) = -1: different voxel
                        [0.06, 0.01, 0.01]])     # new points are not de-duplicated among themselves
    assert R.voxel_keep_indices(existing, new).tolist() == [1, 3, 4]


def test_reset_opacity_is_a_round_trip_in_the_reference():
    st = R.random_state(64, 0, seed=3)
    before = st["opacity_"].clone()
    st = R.reset_opacity(st)
    assert torch.allclose(st["opacity_"], before, rtol=0, atol=2e-3)  # cap = ones_like(...) = 1: no clamping
    assert float(st["exp_avg"][3].abs().sum()) == 0.0


def test_stats_accumulate_only_visible_rows():
    st = R.random_state(6, 0, seed=4)
    st["xyz_gradient_accum_"].zero_(); st["denom_"].zero_(); st["max_radii2D_"].fill_(3.0)
    grad = torch.tensor([[3.0, 4.0, 9.0]] * 6)
    radii = torch.tensor([0, 5, 0, 2, 7, 0], dtype=torch.int32)
    R.add_densification_stats(st, grad, radii)
    assert st["xyz_gradient_accum_"].squeeze().tolist() == [0, 5, 0, 5, 5, 0]
    assert st["denom_"].squeeze().tolist() == [0, 1, 0, 1, 1, 0]
    assert st["max_radii2D_"].tolist() == [3, 5, 3, 3, 7, 3]


def test_store_struct_layout_matches_header():
    import ctypes as C
    from gsdf_slam_b200 import _lib
    # int64, int64, int32 (+pad), 18 pointers, 4 pointers
    assert C.sizeof(_lib.GaussianStoreC) == 8 + 8 + 8 + 18 * 8 + 4 * 8
    assert C.sizeof(_lib.DensifyArgs) == 4 * 4 + 2 * 4 + 8 + 8 + 8


# ---- the restatement against the reference's OWN code, compiled (oracle/build_ref_densify.py) -----------------
def _same_state(a, b, tol=0.0):
    for name in dr.PARAMS + ("exist_since_iter_", "xyz_gradient_accum_", "denom_", "max_radii2D_"):
        assert a[name].shape == b[name].shape, name
        if tol == 0.0 or a[name].dtype != torch.float32:
            assert torch.equal(a[name], b[name]), name
        else:
            assert torch.allclose(a[name], b[name], rtol=tol, atol=tol), name
    for key in ("exp_avg", "exp_avg_sq"):
        for g in range(6):
            assert torch.equal(a[key][g].reshape(b[key][g].shape), b[key][g]), (key, g)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_restatement_matches_the_compiled_reference_densify_and_clone(seed):
    """densifyAndClone + densificationPostfix (the whole densifyAndPrune needs CUDA in the reference — its
    build_rotation allocates on torch::kCUDA — and is compared on the GPU box: tests/test_gpu_densify.py)."""
    if dr.ref_module() is None:
        pytest.skip("oracle/_ref/gsdf_ref_densify.so not built")
    st = dr.random_state(1500, 15, seed=seed)
    ref_st = dr.reference_densify_and_clone(dr.clone_state(st), 0.0002, 4.0, 0.01)
    grads = st["xyz_gradient_accum_"] / st["denom_"]
    grads[grads.isnan()] = 0.0
    f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))  # noqa: E731
    mine, n_clone = dr.densify_and_clone(dr.clone_state(st), grads, f32(0.0002), f32(4.0), f32(0.01))
    assert n_clone > 0 and mine["xyz_"].shape[0] == st["xyz_"].shape[0] + n_clone
    _same_state(mine, ref_st)


def test_restatement_matches_the_compiled_reference_prune_reset_and_stats():
    if dr.ref_module() is None:
        pytest.skip("oracle/_ref/gsdf_ref_densify.so not built")
    st = dr.random_state(800, 15, seed=9)
    g = torch.Generator().manual_seed(4)
    mask = torch.rand(800, generator=g) < 0.3
    _same_state(dr.prune_points(dr.clone_state(st), mask), dr.reference_prune_points(dr.clone_state(st), mask))
    _same_state(dr.reset_opacity(dr.clone_state(st)), dr.reference_reset_opacity(dr.clone_state(st)))
    radii = torch.where(torch.rand(800, generator=g) < 0.6, torch.rand(800, generator=g) * 30, torch.zeros(800))
    grad = torch.randn(800, 3, generator=g)
    mine = dr.add_densification_stats(dr.clone_state(st), grad, radii)
    accum, denom = dr.reference_add_densification_stats(dr.clone_state(st), grad, radii > 0)
    assert torch.equal(mine["xyz_gradient_accum_"], accum) and torch.equal(mine["denom_"], denom)
