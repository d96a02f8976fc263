// Parameter blocks and launchers shared between the C-ABI orchestration
// (gsr_capi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gsr_capi.h"

namespace gsr {
struct PreprocessParams {
  int P, D, M, W, H;
  const float* means3D;
  const float* scales;
  const float* rotations;
  const float* opacities;
  const float* dc;
  const float* shs;
  const float* cov3D_precomp;
  const float* colors_precomp;
  float scale_modifier;
  const float* view;
  const float* proj;
  const float* campos;
  float tan_fovx, tan_fovy, focal_x, focal_y;
  int grid_x, grid_y;
  int prefiltered;
  int vec_ok;
  int activate;           // opacities / scales / rotations are RAW parameters (row f-2)
  int* radii;
  float* depths;
  float2* means2D;
  float4* conic_opacity;
  float* rgb;
  uint8_t* clamped;
  uint32_t* tiles_touched;
  float4* splat_rec;      // [P][3] {x, y, conic.x, conic.y | conic.z, opacity, reject_power, idx | r, g, b, 0}
  uint64_t* vis_items;    // compacted (depth_bits<<32 | idx) of the rendered Gaussians
  uint32_t* num_visible;  // its length (zeroed by the caller)
  uint32_t* cand;         // [P] scratch: Gaussians that pass the conservative cull (preprocess.cu, stage 1)
  uint32_t* num_cand;     // its length (zeroed by the caller)
  uint32_t* large_list;   // work queue of the large rectangles: (chunk << 24 | position in vis_items)
  uint32_t* large_count;  // its length (zeroed by the caller; tile_meta_kernel zeroes it again for the emit stage)
  uint32_t large_capacity;
  uint32_t* tile_count;   // [T] += 1 per (Gaussian, tile) that passes the exact test (zeroed by the caller)
};
struct PreBwdParams {
  int P, D, M;
  int activate;                 // gradients w.r.t. the RAW opacity / scale / rotation (row f-2)
  int accumulate;               // add the parameter gradients to the rows' current values (keyframe batch)
  const int* accumulate_rows;   // optional with accumulate: only rows with a value > 0 are added to, the rest overwritten
  int* grad_rows;               // optional in/out row mask kept by the kernel (gsr_backward_args.grad_rows)
  int dense_zeroed;             // the outputs were zero-filled before this kernel (dense API): rows may be skipped
  const float* opacities;       // raw opacities (needed only when activate)
  const float* means3D;
  const int* radii;
  const float* dc;
  const float* shs;
  const uint8_t* clamped;
  const float* scales;
  const float* rotations;
  const float* cov3D_precomp;
  float scale_modifier;
  const float* view;
  const float* proj;
  const float* campos;
  float focal_x, focal_y, tan_fovx, tan_fovy;
  const float* acc;             // [P][12] moment sums from blend_bwd
  const float4* conic_opacity;  // [P] from the forward's geometry buffer
  int W, H;
  float* dL_dmean2D;
  float* dL_dconic;
  float* dL_dopacity;
  float* dL_dcolor;
  float* dL_dmean3D;
  float* dL_dcov3D;
  float* dL_ddc;
  float* dL_dsh;
  float* dL_dscale;
  float* dL_drot;
};

int launch_precull(const PreprocessParams& p, cudaStream_t s);
int launch_preprocess(const PreprocessParams& p, cudaStream_t s);
int launch_mark_visible(int P, const float* means3D, const float* view, uint8_t* present,
                        cudaStream_t s);
struct GeomHeader;
struct BinHeader;
int launch_tile_meta(const uint32_t* tile_count, uint32_t T, uint2* ranges, GeomHeader* hdr,
                     uint32_t* order, uint32_t* max_contrib, cudaStream_t s);
int launch_emit(int P, const uint64_t* vis_items, const uint32_t* num_visible,
                const uint32_t* tiles_touched, const float2* means2D, const float4* conic_opacity,
                const int* radii, int grid_x, int grid_y, const uint2* ranges, uint32_t* tile_fill,
                uint64_t* items, uint32_t capacity, uint32_t* large_count, uint32_t* large_list,
                cudaStream_t s);
int launch_tile_sort_pack(int P, uint32_t T, const uint2* ranges, const uint32_t* order, const GeomHeader* hdr,
                          BinHeader* bin_hdr, uint32_t capacity, uint64_t* items, uint64_t* scratch,
                          const float4* splat_rec, float4* rec, bool do_mid, bool do_long, uint32_t* sorted_len,
                          uint32_t* redo_flag, uint32_t* redo_list, uint32_t* redo_count, cudaStream_t s);
int launch_tile_sort_full(int P, uint32_t T, const uint2* ranges, const uint32_t* order, const GeomHeader* hdr,
                          uint32_t capacity, uint64_t* items, uint64_t* scratch, const float4* splat_rec, float4* rec,
                          bool do_mid, bool do_long, const uint32_t* redo_list, const uint32_t* redo_count,
                          cudaStream_t s);
int launch_export_binning(uint32_t T, const uint2* ranges, const uint64_t* items, uint64_t* keys,
                          uint32_t* point_list, cudaStream_t s);
// sorted_len != nullptr: lists of kSortSmallCap+ entries are only sorted / packed up to sorted_len[tile]; a tile whose
// pixels are not saturated there is flagged (redo_flag / redo_list / redo_count).  redo: blend the flagged tiles
// (fully sorted by launch_tile_sort_full) instead of all tiles.
int launch_blend_fwd(int W, int H, const uint2* ranges, const uint32_t* tile_order, const float4* rec,
                     const float* bg, float* out_color, float* final_T,
                     uint32_t* n_contrib, float* accum_c, uint32_t* max_contrib, uint32_t capacity,
                     unsigned long long* pair_counts, const uint32_t* sorted_len, uint32_t* redo_flag,
                     uint32_t* redo_list, uint32_t* redo_count, bool redo, cudaStream_t s);
int launch_blend_bwd(int W, int H, const uint2* ranges, const uint32_t* tile_order,
                     const uint32_t* max_contrib, const float4* rec, const float* bg,
                     const float* final_T, const uint32_t* n_contrib, const float* accum_c,
                     const float* dL_dpix, float* acc, cudaStream_t s);
int launch_preprocess_bwd(const PreBwdParams& p, const uint64_t* vis_items, const uint32_t* num_visible,
                          uint32_t* list, uint32_t* num_list, bool zero_fill, cudaStream_t s);
int launch_adam(const gsr_adam_group* groups, int n_groups, double beta1, double beta2, double eps,
                float grad_scale, cudaStream_t s);
size_t knn_temp_bytes(int64_t P);
int launch_adam_rows(const gsr_adam_group* groups, int n_groups, double beta1, double beta2, double eps,
                     float grad_scale, int64_t rows, const int* row_mask, int phase, cudaStream_t s);
int launch_knn(int P, const float* points, float* mean_dist2, char* tmp, cudaStream_t s);

}  // namespace gsr
