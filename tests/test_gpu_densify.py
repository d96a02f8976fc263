"""GPU parity of the map-surgery kernels (csrc/densify.cu) against the reference's
LibTorch op sequences restated in oracle/densify_ref.py (src/gaussian_model.cpp:201-317, 399-681), run
on the same device with the same ATen kernels the reference calls.

Bars: every copied value (parameters, Adam moments, statistics, exist_since_iter, row ORDER and row
count) bit-exact; values the kernels compute themselves — split positions (cuBLAS bmm sums the 3-term dot
product in an unspecified order), log/exp of the split scaling, sigmoid/logit of resetOpacity, the
2-norm of the statistics — within the tolerance written next to each assert."""
import pytest
import torch

from oracle import densify_ref as R

pytestmark = pytest.mark.gpu

STAT_KEYS = ("exist_since_iter_", "xyz_gradient_accum_", "denom_", "max_radii2D_")


def _model(st, M, **kw):
    from gsdf_slam_b200.gaussian_model import GaussianModel
    deg = {0: 0, 3: 1, 8: 2, 15: 3}[M]
    m = GaussianModel(sh_degree=deg, device="cuda", **kw)
    m.load(st["xyz_"], st["features_dc_"], st["features_rest_"], st["opacity_"], st["scaling_"], st["rotation_"],
           exist_since_iter=st["exist_since_iter_"], exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"],
           xyz_gradient_accum=st["xyz_gradient_accum_"], denom=st["denom_"], max_radii2D=st["max_radii2D_"])
    return m


def _model_state(m):
    st = {n: t for n, t in zip(R.PARAMS, m.params())}
    shapes = [t.shape for t in m.params()]
    st["exp_avg"] = [m.exp_avg(g).view(s) for g, s in enumerate(shapes)]
    st["exp_avg_sq"] = [m.exp_avg_sq(g).view(s) for g, s in enumerate(shapes)]
    st["exist_since_iter_"] = m.exist_since_iter_
    st["xyz_gradient_accum_"] = m.xyz_gradient_accum_
    st["denom_"] = m.denom_
    st["max_radii2D_"] = m.max_radii2D_
    return st


def _assert_equal_states(got, ref, skip=()):
    assert got["xyz_"].shape[0] == ref["xyz_"].shape[0], (got["xyz_"].shape, ref["xyz_"].shape)
    for name in R.PARAMS + STAT_KEYS:
        if name in skip:
            continue
        assert got[name].shape == ref[name].shape, name
        assert torch.equal(got[name], ref[name]), f"{name} differs"
    for g in range(6):
        assert torch.equal(got["exp_avg"][g], ref["exp_avg"][g]), f"exp_avg[{g}] differs"
        assert torch.equal(got["exp_avg_sq"][g], ref["exp_avg_sq"][g]), f"exp_avg_sq[{g}] differs"


@pytest.mark.parametrize("P,M,frac", [(50_000, 15, 0.3), (1, 0, 0.0), (1024, 3, 1.0), (1025, 0, 0.5),
                                      (300_001, 15, 0.03), (4097, 8, 0.0)])
def test_prune_points_is_bit_exact(P, M, frac):
    st = R.random_state(P, M, seed=P % 97, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(P)
    mask = torch.rand(P, device="cuda", generator=g) < frac
    if frac == 1.0:
        mask[:] = True
    m = _model(st, M)
    m.prunePoints(mask)
    ref = R.prune_points(R.clone_state(st), mask)
    assert m.P == int((~mask).sum())
    _assert_equal_states(_model_state(m), ref)


@pytest.mark.parametrize("P,M,screen,extent", [(200_000, 15, 0, 4.0), (200_000, 15, 20, 4.0), (70_001, 0, 20, 1.0),
                                               (3000, 3, 0, 4.0), (1, 0, 0, 4.0)])
def test_densify_and_prune_matches_the_libtorch_sequence(P, M, screen, extent):
    st = R.random_state(P, M, seed=11 + P % 13, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    noise = torch.randn((2 * P, 3), device="cuda", generator=g)
    m = _model(st, M)
    counts = m.densifyAndPrune(0.0002, 0.005, extent, screen, noise=noise)
    ref, rc = R.densify_and_prune(R.clone_state(st), 0.0002, 0.005, extent, screen, 0.01, noise=noise)
    got = _model_state(m)
    assert counts["rows"] == ref["xyz_"].shape[0]
    assert counts["split"] == rc["split"]
    if P >= 3000 and extent == 4.0:
        assert rc["cloned"] > P // 50 and rc["split"] > P // 50          # both branches are exercised
        assert ref["xyz_"].shape[0] < P + rc["cloned"] + rc["split"]     # and the final prune
    _assert_equal_states(got, ref, skip=("xyz_", "scaling_"))
    # rows that are not split copies are copies: bit-exact
    n_copy = counts["kept"] + counts["cloned"]
    assert torch.equal(got["xyz_"][:n_copy], ref["xyz_"][:n_copy])
    assert torch.equal(got["scaling_"][:n_copy], ref["scaling_"][:n_copy])
    # split copies: log(exp(s) / 1.6) — the same libdevice routines, 1 ulp allowed
    assert torch.allclose(got["scaling_"][n_copy:], ref["scaling_"][n_copy:], rtol=2e-7, atol=2e-7)
    # positions: R * (z * std) + xyz, summation order of the 3-term dot product differs from cuBLAS bmm
    d = (got["xyz_"][n_copy:] - ref["xyz_"][n_copy:]).abs()
    scale = ref["xyz_"][n_copy:].abs().clamp_min(1.0)
    assert float((d / scale).max()) <= 1e-6 if d.numel() else True


def test_densify_grows_a_too_small_twin_and_keeps_parity():
    P, M = 40_000, 3
    st = R.random_state(P, M, seed=3, device="cuda")
    st["xyz_gradient_accum_"].fill_(1.0)          # every row is selected: 2x rows needed
    st["denom_"].fill_(1.0)
    st["opacity_"].fill_(3.0)
    noise = torch.zeros((2 * P, 3), device="cuda")
    m = _model(st, M, growth=1.0)
    counts = m.densifyAndPrune(0.0002, 0.005, 4.0, 0, noise=noise)
    ref, _ = R.densify_and_prune(R.clone_state(st), 0.0002, 0.005, 4.0, 0, 0.01, noise=noise)
    assert counts["rows"] == ref["xyz_"].shape[0] == 2 * P
    _assert_equal_states(_model_state(m), ref, skip=("xyz_", "scaling_"))
    assert torch.allclose(_model_state(m)["xyz_"], ref["xyz_"], rtol=1e-6, atol=1e-6)


def test_device_drawn_split_noise_is_standard_normal():
    P, M = 100_000, 0
    st = R.random_state(P, M, seed=21, device="cuda")
    st["xyz_gradient_accum_"].fill_(1.0)
    st["denom_"].fill_(1.0)
    st["opacity_"].fill_(3.0)
    st["scaling_"].fill_(0.0)                     # std = 1 on every axis, everything is split
    st["rotation_"] = torch.tensor([[1.0, 0, 0, 0]], device="cuda").repeat(P, 1)
    m = _model(st, M)
    counts = m.densifyAndPrune(0.0002, 0.005, 4.0, 0)
    assert counts["split"] == P and counts["rows"] == 2 * P
    z = m.xyz_ - st["xyz_"].repeat(2, 1)
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert abs(float((z ** 4).mean()) - 3.0) < 0.1                     # kurtosis of a normal
    assert abs(float((z[:P] * z[P:]).mean())) < 0.01                   # the two copies are independent
    assert abs(float((z[:, 0] * z[:, 1]).mean())) < 0.01
    m2 = _model(st, M)
    m2.densifyAndPrune(0.0002, 0.005, 4.0, 0)
    assert torch.equal(m2.xyz_, m.xyz_)                                # same seed, same draws


def test_densification_postfix_appends_and_resets_statistics():
    P, M, n = 30_000, 15, 7_777
    st = R.random_state(P, M, seed=2, device="cuda")
    new = R.random_state(n, M, seed=5, device="cuda")
    m = _model(st, M)                              # capacity < P + n: the store regrows
    args = [new[k] for k in R.PARAMS] + [new["exist_since_iter_"]]
    m.densificationPostfix(*args)
    ref = R.densification_postfix(R.clone_state(st), *args)
    _assert_equal_states(_model_state(m), ref)


def test_add_densification_stats():
    P = 123_457
    st = R.random_state(P, 0, seed=8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    grad = torch.randn((P, 3), device="cuda", generator=g) * 1e-3
    radii = (torch.rand(P, device="cuda", generator=g) * 60 - 20).to(torch.int32)
    m = _model(st, 0)
    m.addDensificationStats(grad, radii)
    ref = R.add_densification_stats(R.clone_state(st), grad, radii)
    assert torch.equal(m.max_radii2D_, ref["max_radii2D_"])
    assert torch.equal(m.denom_, ref["denom_"])
    # sqrt(gx^2 + gy^2): ATen's reduction may or may not contract the sum into an FMA
    assert torch.allclose(m.xyz_gradient_accum_, ref["xyz_gradient_accum_"], rtol=3e-7, atol=0)
    vis = radii > 0
    assert torch.equal(m.xyz_gradient_accum_[~vis], st["xyz_gradient_accum_"][~vis])


def test_reset_opacity():
    P = 50_000
    st = R.random_state(P, 0, seed=9, device="cuda")
    m = _model(st, 0)
    m.resetOpacity()
    ref = R.reset_opacity(R.clone_state(st))
    assert torch.allclose(m.opacity_, ref["opacity_"], rtol=1e-6, atol=1e-7)
    assert float(m.exp_avg(3).abs().sum()) == 0.0 and float(m.exp_avg_sq(3).abs().sum()) == 0.0
    assert torch.equal(m.exp_avg(0), st["exp_avg"][0].reshape(-1))     # other groups untouched
    m.resetOpacity(cap=0.01)
    assert float(torch.sigmoid(m.opacity_).max()) <= 0.01 + 1e-6


def test_increase_pcd_matches_the_libtorch_sequence():
    import gsdf_slam_b200 as gs
    P, M, n = 60_000, 15, 25_000
    st = R.random_state(P, M, seed=4, device="cuda")
    st["xyz_"] = st["xyz_"] * 0.3                  # dense enough that many new points hit occupied voxels
    g = torch.Generator(device="cuda").manual_seed(2)
    pts = torch.randn((n, 3), device="cuda", generator=g) * 0.6
    pts[:500] = st["xyz_"][:500] + 1e-4            # certain duplicates
    cols = torch.rand((n, 3), device="cuda", generator=g)
    m = _model(st, M)
    added = m.increasePcd(pts, cols, iteration=42)
    ref, n_ref = R.increase_pcd(R.clone_state(st), pts, cols, 42, 3, gs.distCUDA2)
    assert added == n_ref and 0 < added < n - 400
    got = _model_state(m)
    _assert_equal_states(got, ref, skip=("scaling_", "opacity_", "features_dc_"))
    assert torch.allclose(got["features_dc_"], ref["features_dc_"], rtol=3e-7, atol=1e-7)
    assert torch.allclose(got["scaling_"], ref["scaling_"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(got["opacity_"], ref["opacity_"], rtol=1e-6, atol=0)
    # the dedup itself: exact set equality with torch.isin on the int64 voxel hashes
    keep_ref = R.voxel_keep_indices(st["xyz_"], pts)
    assert torch.equal(got["xyz_"][P:], pts[keep_ref])


def test_create_from_pcd_matches_the_libtorch_sequence():
    import gsdf_slam_b200 as gs
    from gsdf_slam_b200.gaussian_model import GaussianModel
    g = torch.Generator(device="cuda").manual_seed(12)
    n = 80_001
    pts = torch.randn((n, 3), device="cuda", generator=g) * 3.0
    pts[100:200] = pts[:100]                       # duplicates: distance 0 -> the 1e-7 clamp
    cols = torch.rand((n, 3), device="cuda", generator=g)
    m = GaussianModel(sh_degree=3, device="cuda")
    m.createFromPcd(pts, cols, 1.0)
    ref = R.create_from_pcd(pts, cols, 3, gs.distCUDA2)
    assert m.P == n
    assert torch.equal(m.xyz_, ref["xyz_"]) and torch.equal(m.rotation_, ref["rotation_"])
    assert torch.equal(m.features_rest_, ref["features_rest_"]) and torch.equal(m.exist_since_iter_, ref["exist_since_iter_"])
    assert torch.allclose(m.features_dc_, ref["features_dc_"], rtol=3e-7, atol=1e-7)
    assert torch.allclose(m.scaling_, ref["scaling_"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(m.opacity_, ref["opacity_"], rtol=1e-6, atol=0)
    assert float(m.max_radii2D_.abs().sum()) == 0.0 and float(m.exp_avg(2).abs().sum()) == 0.0


def test_voxel_dedup_edge_cases():
    from gsdf_slam_b200.gaussian_model import GaussianModel
    m = GaussianModel(sh_degree=0, device="cuda")
    pts = torch.tensor([[0.01, 0.01, 0.01], [0.02, 0.02, 0.02], [5.0, 5.0, 5.0]], device="cuda")
    cols = torch.full((3, 3), 0.5, device="cuda")
    assert m.increasePcd(pts, cols, 1) == 3        # empty map: nothing is a duplicate, not even among new points
    assert m.increasePcd(pts, cols, 2) == 0        # now all of them are
    assert m.increasePcd(pts[:0], cols[:0], 3) == 0
    neg = torch.tensor([[-0.01, 0.01, 0.01]], device="cuda")
    assert m.increasePcd(neg, cols[:1], 4) == 1    # floor, not truncation
    assert m.P == 4 and m.exist_since_iter_.tolist() == [1, 1, 1, 4]


def test_mapping_loop_with_varying_row_count_stays_in_lockstep():
    """cfg3-style sequence: statistics -> densifyAndPrune -> increasePcd -> prune, four rounds; the product
    model and the LibTorch-op oracle must hold the same rows in the same order after every call."""
    import gsdf_slam_b200 as gs
    P, M = 20_000, 3
    st = R.random_state(P, M, seed=6, device="cuda")
    st["xyz_gradient_accum_"].zero_(); st["denom_"].zero_(); st["max_radii2D_"].zero_()
    m = _model(st, M)
    ref = R.clone_state(st)
    g = torch.Generator(device="cuda").manual_seed(77)
    sizes = [P]
    for it in range(4):
        n = ref["xyz_"].shape[0]
        for _ in range(3):
            grad = torch.randn((n, 3), device="cuda", generator=g) * 4e-4
            radii = (torch.rand(n, device="cuda", generator=g) * 50 - 15).to(torch.int32)
            m.addDensificationStats(grad, radii)
            R.add_densification_stats(ref, grad, radii)
        # identical statistics on both sides from here on (the 2-norm may differ in the last bit)
        m.xyz_gradient_accum_.copy_(ref["xyz_gradient_accum_"])
        noise = torch.randn((2 * n, 3), device="cuda", generator=g)
        m.densifyAndPrune(0.0002, 0.005, 4.0, 20 if it >= 2 else 0, noise=noise)
        ref, _ = R.densify_and_prune(ref, 0.0002, 0.005, 4.0, 20 if it >= 2 else 0, 0.01, noise=noise)
        _assert_equal_states(_model_state(m), ref, skip=("xyz_", "scaling_"))
        # carry the oracle's computed values over so that later decisions see identical inputs
        m.xyz_.copy_(ref["xyz_"]); m.scaling_.copy_(ref["scaling_"])
        pts = torch.randn((5000, 3), device="cuda", generator=g) * 2.5
        cols = torch.rand((5000, 3), device="cuda", generator=g)
        m.increasePcd(pts, cols, it)
        ref, _ = R.increase_pcd(ref, pts, cols, it, 1, gs.distCUDA2)
        m.scaling_.copy_(ref["scaling_"]); m.opacity_.copy_(ref["opacity_"]); m.features_dc_.copy_(ref["features_dc_"])
        n = ref["xyz_"].shape[0]
        mask = torch.rand(n, device="cuda", generator=g) < 0.1
        m.prunePoints(mask)
        ref = R.prune_points(ref, mask)
        _assert_equal_states(_model_state(m), ref)
        sizes.append(m.P)
    assert len(set(sizes)) == len(sizes)           # the row count really changed every round


# ---- against the reference's OWN code, compiled (oracle/build_ref_densify.py -> oracle/_ref/gsdf_ref_densify.so) -----
@pytest.mark.parametrize("P,M,screen", [(120_000, 15, 20), (30_001, 3, 0)])
def test_densify_and_prune_against_the_compiled_reference(P, M, screen):
    """densifyAndPrune of src/gaussian_model.cpp:644-665 itself (clone, split with at::normal + build_rotation + bmm,
    the two prunes, the Adam-state surgery keyed by TensorImpl): (a) pins the ATen restatement oracle/densify_ref.py on
    the GPU, bit for bit; (b) the product's fused kernel against it with the bars of this file."""
    if R.ref_module() is None:
        pytest.skip("oracle/_ref/gsdf_ref_densify.so not built")
    st = R.random_state(P, M, seed=17 + P % 7, device="cuda")
    ref, noise = R.reference_densify_and_prune(R.clone_state(st), 0.0002, 0.005, 4.0, screen, 0.01, seed=1234 + P)
    mine, rc = R.densify_and_prune(R.clone_state(st), 0.0002, 0.005, 4.0, screen, 0.01, noise=noise)
    assert rc["cloned"] > P // 50 and rc["split"] > P // 50 and noise.shape[0] == 2 * rc["split"]
    _assert_equal_states(mine, ref)                               # (a) restatement == reference code, sampled rows too
    m = _model(st, M)
    pad = torch.zeros((2 * P, 3), device="cuda")
    pad[: noise.shape[0]] = noise
    counts = m.densifyAndPrune(0.0002, 0.005, 4.0, screen, noise=pad)
    got = _model_state(m)
    assert counts["rows"] == ref["xyz_"].shape[0] and counts["split"] == rc["split"]
    _assert_equal_states(got, ref, skip=("xyz_", "scaling_"))     # (b) product == reference code
    n_copy = counts["kept"] + counts["cloned"]
    assert torch.equal(got["xyz_"][:n_copy], ref["xyz_"][:n_copy])
    assert torch.equal(got["scaling_"][:n_copy], ref["scaling_"][:n_copy])
    assert torch.allclose(got["scaling_"][n_copy:], ref["scaling_"][n_copy:], rtol=2e-7, atol=2e-7)
    d = (got["xyz_"][n_copy:] - ref["xyz_"][n_copy:]).abs()
    assert float((d / ref["xyz_"][n_copy:].abs().clamp_min(1.0)).max()) <= 1e-6


def test_prune_reset_and_statistics_against_the_compiled_reference():
    if R.ref_module() is None:
        pytest.skip("oracle/_ref/gsdf_ref_densify.so not built")
    P, M = 60_000, 15
    st = R.random_state(P, M, seed=5, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(8)
    mask = torch.rand(P, device="cuda", generator=g) < 0.25
    m = _model(st, M)
    m.prunePoints(mask)
    _assert_equal_states(_model_state(m), R.reference_prune_points(R.clone_state(st), mask))
    m = _model(st, M)
    m.resetOpacity()
    ref = R.reference_reset_opacity(R.clone_state(st))
    got = _model_state(m)
    assert torch.allclose(got["opacity_"], ref["opacity_"], rtol=1e-6, atol=1e-6)
    assert float(got["exp_avg"][3].abs().max()) == 0.0 and float(ref["exp_avg"][3].abs().max()) == 0.0
    radii = torch.where(torch.rand(P, device="cuda", generator=g) < 0.6,
                        torch.randint(1, 40, (P,), device="cuda", generator=g), torch.zeros(P, device="cuda", dtype=torch.int64))
    grad = torch.randn(P, 3, device="cuda", generator=g)
    m = _model(st, M)
    m.addDensificationStats(grad, radii.to(torch.int32))
    accum, denom = R.reference_add_densification_stats(R.clone_state(st), grad, radii > 0)
    got = _model_state(m)
    assert torch.equal(got["denom_"], denom)
    assert torch.allclose(got["xyz_gradient_accum_"], accum, rtol=1e-6, atol=1e-9)
