"""ctypes binding of the C-ABI product library ``libgsr.so`` (include/gsr_capi.h).

The library is hand-written CUDA for sm_100a; there is NO CPU fallback.  If the
shared object is missing this module raises at first use (loudly), it never
routes to the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSR_LIB_PATH: developer override to A/B kernel variants (tools/); default is the in-tree build
LIB_PATH = os.environ.get("GSR_LIB_PATH", os.path.join(_HERE, "libgsr.so"))

RESIZE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class ForwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("W", C.c_int), ("H", C.c_int),
        ("background", C.c_void_p), ("means3D", C.c_void_p), ("dc", C.c_void_p),
        ("shs", C.c_void_p), ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p),
        ("scales", C.c_void_p), ("scale_modifier", C.c_float), ("rotations", C.c_void_p),
        ("cov3D_precomp", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
        ("campos", C.c_void_p), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("prefiltered", C.c_int), ("debug", C.c_int),
        ("out_color", C.c_void_p), ("radii", C.c_void_p),
        ("geom_resize", RESIZE_FN), ("geom_user", C.c_void_p),
        ("binning_resize", RESIZE_FN), ("binning_user", C.c_void_p),
        ("image_resize", RESIZE_FN), ("image_user", C.c_void_p),
        ("sample_resize", RESIZE_FN), ("sample_user", C.c_void_p),
        ("stream", C.c_void_p),
        ("num_rendered", C.POINTER(C.c_int)), ("num_buckets", C.POINTER(C.c_int)),
        ("raw_params", C.c_int), ("radii_ready_event", C.c_void_p), ("context", C.c_void_p),
    ]


class BackwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("R", C.c_int), ("B", C.c_int),
        ("W", C.c_int), ("H", C.c_int),
        ("background", C.c_void_p), ("means3D", C.c_void_p), ("dc", C.c_void_p),
        ("shs", C.c_void_p), ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p),
        ("scales", C.c_void_p), ("scale_modifier", C.c_float), ("rotations", C.c_void_p),
        ("cov3D_precomp", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
        ("campos", C.c_void_p), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("radii", C.c_void_p),
        ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p),
        ("image_buffer", C.c_void_p), ("sample_buffer", C.c_void_p),
        ("dL_dpix", C.c_void_p),
        ("dL_dmean2D", C.c_void_p), ("dL_dconic", C.c_void_p), ("dL_dopacity", C.c_void_p),
        ("dL_dcolor", C.c_void_p), ("dL_dmean3D", C.c_void_p), ("dL_dcov3D", C.c_void_p),
        ("dL_ddc", C.c_void_p), ("dL_dsh", C.c_void_p), ("dL_dscale", C.c_void_p),
        ("dL_drot", C.c_void_p),
        ("debug", C.c_int), ("stream", C.c_void_p), ("raw_params", C.c_int), ("rendered_rows_only", C.c_int),
        ("accumulate", C.c_int), ("accumulate_rows", C.c_void_p), ("grad_rows", C.c_void_p),
    ]


class AdamGroup(C.Structure):
    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p),
        ("exp_avg_sq", C.c_void_p), ("numel", C.c_int64), ("lr", C.c_double),
        ("step", C.c_int64),
    ]


GSR_MAX_PEERS = 8
GSR_PEER_MAX_SEGMENTS = 12


class PeerSegment(C.Structure):
    _fields_ = [("start", C.c_int64), ("numel", C.c_int64), ("width", C.c_int32), ("rule", C.c_int32), ("lr", C.c_double)]


class PeerAdamArgs(C.Structure):
    _fields_ = [
        ("n_peers", C.c_int), ("rank", C.c_int),
        ("param_peers", C.c_void_p * GSR_MAX_PEERS), ("grad_peers", C.c_void_p * GSR_MAX_PEERS),
        ("n_segments", C.c_int), ("segments", PeerSegment * GSR_PEER_MAX_SEGMENTS),
        ("shard_lo", C.c_int64), ("shard_hi", C.c_int64),
        ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("union_mask", C.c_void_p),
        ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("grad_scale", C.c_float),
        ("step", C.c_int64), ("stream", C.c_void_p), ("shard_chunk", C.c_int64),
        ("param_multicast", C.c_void_p), ("traffic_counters", C.c_void_p),
    ]


class MapperConfig(C.Structure):
    """gsr_mapper_config (include/gsr_mapper.h)."""
    _fields_ = [("P", C.c_int64), ("M", C.c_int), ("sh_degree", C.c_int), ("rank", C.c_int), ("world", C.c_int),
                ("lr", C.c_double * 6), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("raw_params", C.c_int), ("with_stats", C.c_int), ("lambda_dssim", C.c_float)]


class MapperKeyframe(C.Structure):
    """gsr_mapper_keyframe (include/gsr_mapper.h)."""
    _fields_ = [("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("W", C.c_int), ("H", C.c_int),
                ("dL_dpix", C.c_void_p), ("gt_image", C.c_void_p), ("mask", C.c_void_p), ("out_color", C.c_void_p)]


GSR_PEER_RULE_ADAM, GSR_PEER_RULE_SUM, GSR_PEER_RULE_MAX = 0, 1, 2
GSR_MAPPER_HANDLE_BYTES = 256
GSR_STORE_GROUPS = 6


class GaussianStoreC(C.Structure):
    """gsr_gaussian_store (include/gsr_capi.h)."""
    _fields_ = [
        ("rows", C.c_int64), ("capacity", C.c_int64), ("n_rest", C.c_int32),
        ("param", C.c_void_p * GSR_STORE_GROUPS), ("exp_avg", C.c_void_p * GSR_STORE_GROUPS),
        ("exp_avg_sq", C.c_void_p * GSR_STORE_GROUPS),
        ("exist_since_iter", C.c_void_p), ("xyz_gradient_accum", C.c_void_p), ("denom", C.c_void_p),
        ("max_radii2D", C.c_void_p),
    ]


class DensifyArgs(C.Structure):
    _fields_ = [
        ("max_grad", C.c_float), ("min_opacity", C.c_float), ("extent", C.c_float),
        ("percent_dense", C.c_float), ("max_screen_size", C.c_int32), ("n_split", C.c_int32),
        ("noise", C.c_void_p), ("seed", C.c_uint64), ("noise_rows", C.c_int64),
    ]


# every symbol include/gsr_capi.h declares (checked by tests/test_capi_symbols.py)
EXPORTED_SYMBOLS = [
    "gsr_forward", "gsr_backward", "gsr_mark_visible", "gsr_knn_mean_dist2", "gsr_adam_step", "gsr_adam_step_rows", "gsr_adam_step_peer", "gsr_peer_union_mask", "gsr_peer_rank_mask_u8",
    "gsr_peer_local_stats", "gsr_peer_stats_exchange", "gsr_export_redo_count",
    "gsr_l1_ssim_loss", "gsr_l1_ssim_workspace_bytes",
    "gsr_store_workspace_bytes", "gsr_prune_points", "gsr_densify_and_prune", "gsr_densification_postfix",
    "gsr_add_densification_stats", "gsr_reset_opacity", "gsr_voxel_dedup_workspace_bytes", "gsr_voxel_dedup",
    "gsr_select_rows", "gsr_increase_pcd_rows", "gsr_colorize_workspace_bytes", "gsr_colorize_points", "gsr_pack_xyz",
    "gsr_ply_pack_rows", "gsr_ply_unpack_rows",
    "gsr_export_binning", "gsr_export_image", "gsr_export_geometry", "gsr_export_pair_counts", "gsr_sort_temp_bytes",
    "gsr_sort_u64", "gsr_launch_count", "gsr_version", "gsr_profile_enable", "gsr_profile_read",
    "gsr_context_create", "gsr_context_destroy",
    # include/gsr_mapper.h
    "gsr_mapper_create", "gsr_mapper_connect", "gsr_mapper_param", "gsr_mapper_stat", "gsr_mapper_set_lr",
    "gsr_mapper_step", "gsr_mapper_steps_done", "gsr_mapper_destroy",
]

_lib = None


class _Missing:
    """Stand-in for a symbol an older developer build (GSR_LIB_PATH) does not export."""
    argtypes = restype = None

    def __init__(self, name):
        self.name = name

    def __call__(self, *a, **k):
        raise RuntimeError(f"{LIB_PATH} does not export {self.name}")


def build(verbose: bool = False) -> str:
    """Compile libgsr.so in-tree with nvcc for sm_100a (no GPU needed)."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    res = subprocess.run(["bash", script], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libgsr.so failed:\n" + res.stderr[-4000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the product library; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "There is no CPU fallback for the rasterizer hot path.")
    L = C.CDLL(LIB_PATH)
    if "GSR_LIB_PATH" in os.environ:
        # developer A/B builds (variants/*.so) may predate a symbol: calling it raises, loading does not
        for name in EXPORTED_SYMBOLS:
            if not hasattr(L, name):
                setattr(L, name, _Missing(name))
    L.gsr_forward.argtypes = [C.POINTER(ForwardArgs)]
    L.gsr_forward.restype = C.c_int
    L.gsr_backward.argtypes = [C.POINTER(BackwardArgs)]
    L.gsr_backward.restype = C.c_int
    L.gsr_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsr_mark_visible.restype = C.c_int
    L.gsr_knn_mean_dist2.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsr_knn_mean_dist2.restype = C.c_int
    L.gsr_adam_step.argtypes = [C.POINTER(AdamGroup), C.c_int, C.c_double, C.c_double, C.c_double,
                                C.c_float, C.c_void_p]
    L.gsr_adam_step.restype = C.c_int
    L.gsr_adam_step_rows.argtypes = [C.POINTER(AdamGroup), C.c_int, C.c_double, C.c_double, C.c_double,
                                     C.c_float, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    L.gsr_adam_step_rows.restype = C.c_int
    L.gsr_adam_step_peer.argtypes = [C.POINTER(PeerAdamArgs)]
    L.gsr_adam_step_peer.restype = C.c_int
    L.gsr_peer_union_mask.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.gsr_peer_union_mask.restype = C.c_int
    L.gsr_peer_rank_mask_u8.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.gsr_peer_rank_mask_u8.restype = C.c_int
    L.gsr_peer_local_stats.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
    L.gsr_peer_local_stats.restype = C.c_int
    L.gsr_peer_stats_exchange.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int64, C.c_int, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsr_peer_stats_exchange.restype = C.c_int
    L.gsr_l1_ssim_workspace_bytes.argtypes = [C.c_int, C.c_int]
    L.gsr_l1_ssim_workspace_bytes.restype = C.c_size_t
    L.gsr_l1_ssim_loss.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsr_l1_ssim_loss.restype = C.c_int
    SP = C.POINTER(GaussianStoreC)
    L.gsr_store_workspace_bytes.argtypes = [C.c_int64]
    L.gsr_store_workspace_bytes.restype = C.c_size_t
    L.gsr_prune_points.argtypes = [SP, SP, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
    L.gsr_prune_points.restype = C.c_int
    L.gsr_densify_and_prune.argtypes = [SP, SP, C.POINTER(DensifyArgs), C.c_void_p, C.POINTER(C.c_int64),
                                        C.c_void_p]
    L.gsr_densify_and_prune.restype = C.c_int
    L.gsr_densification_postfix.argtypes = [SP, C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_void_p]
    L.gsr_densification_postfix.restype = C.c_int
    L.gsr_add_densification_stats.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
    L.gsr_add_densification_stats.restype = C.c_int
    L.gsr_reset_opacity.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    L.gsr_reset_opacity.restype = C.c_int
    L.gsr_voxel_dedup_workspace_bytes.argtypes = [C.c_int64]
    L.gsr_voxel_dedup_workspace_bytes.restype = C.c_size_t
    L.gsr_voxel_dedup.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_float, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
    L.gsr_voxel_dedup.restype = C.c_int
    L.gsr_select_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.POINTER(C.c_int64), C.c_void_p]
    L.gsr_select_rows.restype = C.c_int
    L.gsr_increase_pcd_rows.argtypes = [SP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                        C.c_int32, C.c_void_p]
    L.gsr_increase_pcd_rows.restype = C.c_int
    L.gsr_colorize_workspace_bytes.argtypes = [C.c_int64]
    L.gsr_colorize_workspace_bytes.restype = C.c_size_t
    L.gsr_colorize_points.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
    L.gsr_colorize_points.restype = C.c_int
    L.gsr_pack_xyz.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    L.gsr_pack_xyz.restype = C.c_int
    L.gsr_ply_pack_rows.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 8
    L.gsr_ply_pack_rows.restype = C.c_int
    L.gsr_ply_unpack_rows.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 8
    L.gsr_ply_unpack_rows.restype = C.c_int
    L.gsr_export_binning.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsr_export_binning.restype = C.c_int
    L.gsr_export_image.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsr_export_image.restype = C.c_int
    L.gsr_export_redo_count.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.c_void_p]
    L.gsr_export_redo_count.restype = C.c_int
    L.gsr_export_geometry.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gsr_export_geometry.restype = C.c_int
    L.gsr_export_pair_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
    L.gsr_export_pair_counts.restype = C.c_int
    L.gsr_sort_temp_bytes.argtypes = [C.c_int64]
    L.gsr_sort_temp_bytes.restype = C.c_size_t
    L.gsr_sort_u64.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p]
    L.gsr_sort_u64.restype = C.c_int
    L.gsr_launch_count.argtypes = []
    L.gsr_launch_count.restype = C.c_int64
    L.gsr_profile_enable.argtypes = [C.c_int]
    L.gsr_profile_enable.restype = None
    L.gsr_profile_read.argtypes = [C.c_char_p, C.c_size_t]
    L.gsr_profile_read.restype = C.c_int
    L.gsr_context_create.argtypes = []
    L.gsr_context_create.restype = C.c_void_p
    L.gsr_context_destroy.argtypes = [C.c_void_p]
    L.gsr_context_destroy.restype = None
    L.gsr_mapper_create.argtypes = [C.POINTER(MapperConfig), C.c_void_p]
    L.gsr_mapper_create.restype = C.c_void_p
    L.gsr_mapper_connect.argtypes = [C.c_void_p, C.c_void_p]
    L.gsr_mapper_connect.restype = C.c_int
    L.gsr_mapper_param.argtypes = [C.c_void_p, C.c_int]
    L.gsr_mapper_param.restype = C.c_void_p
    L.gsr_mapper_stat.argtypes = [C.c_void_p, C.c_int]
    L.gsr_mapper_stat.restype = C.c_void_p
    L.gsr_mapper_set_lr.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.gsr_mapper_set_lr.restype = C.c_int
    L.gsr_mapper_step.argtypes = [C.c_void_p, C.POINTER(MapperKeyframe), C.c_int, C.c_void_p, C.c_float, C.c_void_p,
                                  C.c_void_p]
    L.gsr_mapper_step.restype = C.c_int
    L.gsr_mapper_steps_done.argtypes = [C.c_void_p]
    L.gsr_mapper_steps_done.restype = C.c_int64
    L.gsr_mapper_destroy.argtypes = [C.c_void_p]
    L.gsr_mapper_destroy.restype = None
    L.gsr_version.argtypes = []
    L.gsr_version.restype = C.c_char_p
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    """Translate a non-zero C-ABI return code into RuntimeError (the reference's
    shim throws std::runtime_error from CHECK_CUDA, auxiliary.h:196-207)."""
    if rc != 0:
        kind = "gsr error" if rc < 0 else "cudaError"
        raise RuntimeError(f"{what} failed: {kind} {rc}")
