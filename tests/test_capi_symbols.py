"""CPU tests: the C-ABI library loads and exports every symbol include/gsr_capi.h declares
(no compute calls: there is no GPU here), and argument validation works without a device."""
import ctypes as C
import os
import re

import pytest

import gsdf_slam_b200 as gs
from gsdf_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    out = set()
    for name in ("gsr_capi.h", "gsr_mapper.h"):
        hdr = open(os.path.join(ROOT, "include", name)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        out |= set(re.findall(r"\b(gsr_[a-z0-9_]+)\s*\(", hdr))
    return sorted(out)


def test_header_symbols_are_exported():
    L = _lib.lib()
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/*.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in L.gsr_version()


def test_library_is_in_tree_and_no_cpu_fallback():
    assert os.path.dirname(_lib.LIB_PATH) == os.path.join(ROOT, "gsdf_slam_b200")
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            gs.RasterizeGaussiansCUDA(torch.zeros(3), torch.zeros((4, 3)), torch.zeros(0), torch.zeros((4, 1)),
                                      torch.zeros((4, 3)), torch.zeros((4, 4)), 1.0, torch.zeros(0),
                                      torch.eye(4), torch.eye(4), 1.0, 1.0, 16, 16, torch.zeros((4, 1, 3)),
                                      torch.zeros((4, 15, 3)), 3, torch.zeros(3), False, False)


def test_argument_validation_without_device():
    L = _lib.lib()
    assert L.gsr_forward(None) == -1                      # GSR_ERR_BAD_ARG
    assert L.gsr_backward(None) == -1
    a = _lib.ForwardArgs()
    a.P, a.W, a.H = -1, 16, 16
    assert L.gsr_forward(C.byref(a)) == -1
    assert L.gsr_sort_u64(None, None, None, -5, 0, 64, None, None) == -1
    assert L.gsr_sort_u64(None, None, None, 0, 0, 64, None, None) == 0      # empty input is fine
    assert L.gsr_adam_step(None, 3, 0.9, 0.999, 1e-15, 1.0, None) == -1
    assert L.gsr_mark_visible(-1, None, None, None, None, None) == -1
    assert L.gsr_knn_mean_dist2(0, None, None, None) == 0
    assert int(L.gsr_sort_temp_bytes(1_000_000)) > 0


def test_store_and_ingest_argument_validation_without_device():
    """Rows f-3 / f-4: bad arguments are rejected before any CUDA call, empty inputs are fine."""
    L = _lib.lib()
    bad = -1
    assert L.gsr_prune_points(None, None, None, None, None, None) == bad
    assert L.gsr_densify_and_prune(None, None, None, None, None, None) == bad
    st = _lib.GaussianStoreC()
    st.rows, st.capacity, st.n_rest = 10, 5, 15                     # rows > capacity
    assert L.gsr_prune_points(C.byref(st), C.byref(st), None, None, None, None) == bad
    assert L.gsr_densification_postfix(C.byref(st), None, None, 1, None) == bad
    assert L.gsr_add_densification_stats(-1, None, None, None, None, None, None) == bad
    assert L.gsr_add_densification_stats(0, None, None, None, None, None, None) == 0
    assert L.gsr_add_densification_stats(5, None, None, None, None, None, None) == bad
    assert L.gsr_reset_opacity(0, None, None, None, 1.0, None) == 0
    assert L.gsr_reset_opacity(3, None, None, None, 1.0, None) == bad
    assert L.gsr_voxel_dedup(None, 0, None, 0, 0.05, None, None, None) == 0
    assert L.gsr_voxel_dedup(None, 0, None, 4, 0.0, None, None, None) == bad      # voxel size must be > 0
    assert L.gsr_voxel_dedup(None, 0, None, 4, 0.05, None, None, None) == bad     # missing buffers
    assert L.gsr_select_rows(None, -1, 0, None, None, None, None, None) == bad
    assert L.gsr_colorize_points(None, 0, None, None, None, None, 64, 48, None, None, None, None, None) == 0
    assert L.gsr_colorize_points(None, 8, None, None, None, None, 64, 48, None, None, None, None, None) == bad
    assert L.gsr_colorize_points(None, 8, None, None, None, None, 0, 48, None, None, None, None, None) == bad
    assert L.gsr_ply_pack_rows(0, 15, None, None, None, None, None, None, None, None) == 0
    assert L.gsr_ply_pack_rows(4, 15, None, None, None, None, None, None, None, None) == bad
    assert L.gsr_peer_rank_mask_u8(None, 2, 10, None, None) == bad
    assert int(L.gsr_store_workspace_bytes(1_000_000)) >= 1_000_000
    assert int(L.gsr_voxel_dedup_workspace_bytes(1_000_000)) >= 16_000_000          # >= 2 slots of 8 bytes per key
    assert int(L.gsr_colorize_workspace_bytes(1000)) > 5000


def test_means3d_shape_error_matches_reference_message():
    import torch
    with pytest.raises(RuntimeError, match="means3D must have dimensions"):
        gs.RasterizeGaussiansCUDA(torch.zeros(3), torch.zeros(12), torch.zeros(0), torch.zeros((4, 1)),
                                  torch.zeros((4, 3)), torch.zeros((4, 4)), 1.0, torch.zeros(0), torch.eye(4),
                                  torch.eye(4), 1.0, 1.0, 16, 16, torch.zeros((4, 1, 3)),
                                  torch.zeros((4, 15, 3)), 3, torch.zeros(3), False, False)


def test_rasterizer_module_argument_errors():
    import torch
    rs = gs.GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3,
                                          torch.zeros(3))
    r = gs.GaussianRasterizer(rs)
    m = torch.zeros((4, 3))
    with pytest.raises(RuntimeError, match="exactly one of either scale/rotation"):
        r(m, m, torch.zeros((4, 1)), dc=torch.zeros((4, 1, 3)), shs=torch.zeros((4, 15, 3)))
    with pytest.raises(RuntimeError, match="SHs or precomputed colors"):
        r(m, m, torch.zeros((4, 1)), scales=m, rotations=torch.zeros((4, 4)))
