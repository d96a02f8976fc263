/*
 * gsr_capi.h — C ABI of the B200-native Gaussian-splatting rasterizer hot path.
 *
 * This is the drop-in boundary for GSDF-SLAM's mapping hot path.  Each entry
 * point below replaces one reference interface (file:line under /root/reference):
 *
 *   gsr_forward        <- CudaRasterizer::Rasterizer::forward   cuda_rasterizer/rasterizer.h:27-55
 *                         (called by RasterizeGaussiansCUDA,     src/rasterize_points.cu:38-117)
 *   gsr_backward       <- CudaRasterizer::Rasterizer::backward  cuda_rasterizer/rasterizer.h:57-94
 *                         (called by RasterizeGaussiansBackwardCUDA, src/rasterize_points.cu:120-205)
 *   gsr_mark_visible   <- CudaRasterizer::Rasterizer::markVisible cuda_rasterizer/rasterizer.h:22-26
 *                         (called by markVisible,                src/rasterize_points.cu:207-226)
 *   gsr_knn_mean_dist2 <- SimpleKNN::knn                          submodules/simple-knn/simple_knn.h:17
 *                         (called by distCUDA2,                  submodules/simple-knn/spatial.cu:15-26)
 *   gsr_adam_step,     <- torch::optim::Adam::step on the 6 parameter groups
 *   gsr_adam_step_rows    (src/gaussian_model.cpp:319-352, stepped at src/gaussian_mapper.cpp:465)
 *   gsr_prune_points, gsr_densify_and_prune, gsr_densification_postfix, gsr_reset_opacity,
 *   gsr_add_densification_stats, gsr_voxel_dedup, gsr_select_rows, gsr_increase_pcd_rows
 *                      <- GaussianModel::prunePoints / densifyAndPrune / densificationPostfix /
 *                         resetOpacity / addDensificationStats / increasePcd
 *                         (src/gaussian_model.cpp:201-317, 399-681)
 *   gsr_colorize_points <- colorize_launcher, cuda_colorize/colorize_pointcloud.cu:57-107 (row f-4)
 *   gsr_export_*       <- test/debug decoders of the opaque scratch buffers (the
 *                         reference's GeometryState/BinningState/ImageState::fromChunk,
 *                         cuda_rasterizer/rasterizer_impl.cu:274-343)
 *
 * Conventions (same as the reference): every pointer is a DEVICE pointer to
 * contiguous fp32/int32 data unless stated; viewmatrix/projmatrix are stored
 * transposed (element m[4*col+row]); optional inputs are NULL when absent.
 * Scratch memory is obtained through C callbacks — the C analogue of the
 * reference's std::function<char*(size_t)> (rasterizer.h:28-31).
 * All functions return 0 on success, a negative gsr error code or a positive
 * cudaError_t otherwise; nothing throws across this ABI.  All work is enqueued
 * on `stream` (a cudaStream_t).  gsr_forward sizes the per-duplicate buffer
 * optimistically from its previous call and only waits for the 24-byte count block
 * after everything is enqueued (first call of a context: one blocking wait; see
 * DESIGN.md).  That host-side state lives in a gsr_context (below): pass your own, or
 * leave `context` NULL for a default one per (host thread, device).  There is no other
 * hidden state between calls; gsr_forward / gsr_backward may be entered from several
 * host threads (the reference's mapper and viewer threads) on different streams.
 * Limits: P < 2^24 Gaussians per call (GSR_ERR_BAD_ARG beyond).
 */
#ifndef GSR_CAPI_H_
#define GSR_CAPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_OK 0
#define GSR_ERR_BAD_ARG (-1)
#define GSR_ERR_ALLOC (-2)
#define GSR_ERR_NO_DEVICE (-3)

/* Host-side state of gsr_forward (read-back event, pinned words, buffer-size and sort-kernel hints),
 * bound to the device that is current at creation.  One context must not be used by two threads at once. */
typedef struct gsr_context gsr_context;
gsr_context* gsr_context_create(void);
void gsr_context_destroy(gsr_context* ctx);

/* Resizes the caller-owned scratch buffer to `bytes` and returns its device address. */
typedef char* (*gsr_resize_fn)(void* user, size_t bytes);

typedef struct gsr_forward_args {
  int P;              /* number of Gaussians */
  int D;              /* active SH degree (0..3) */
  int M;              /* number of "rest" SH coefficients per Gaussian in `shs` (0 or 15) */
  int W, H;           /* image size */
  const float* background;     /* [3] */
  const float* means3D;        /* [P,3] */
  const float* dc;             /* [P,3]   SH band 0 */
  const float* shs;            /* [P,M,3] or NULL */
  const float* colors_precomp; /* [P,3] or NULL */
  const float* opacities;      /* [P] */
  const float* scales;         /* [P,3] or NULL (then cov3D_precomp) */
  float scale_modifier;
  const float* rotations;      /* [P,4] or NULL */
  const float* cov3D_precomp;  /* [P,6] or NULL */
  const float* viewmatrix;     /* [16] */
  const float* projmatrix;     /* [16] */
  const float* campos;         /* [3] */
  float tan_fovx, tan_fovy;
  int prefiltered;
  int debug;                   /* !=0: synchronise + check after the call */
  float* out_color;            /* [3,H,W] out */
  int* radii;                  /* [P] out */
  gsr_resize_fn geom_resize;    void* geom_user;
  gsr_resize_fn binning_resize; void* binning_user;
  gsr_resize_fn image_resize;   void* image_user;
  gsr_resize_fn sample_resize;  void* sample_user;
  void* stream;                /* cudaStream_t */
  int* num_rendered;           /* host out: number of (Gaussian,tile) duplicates R */
  int* num_buckets;            /* host out: kept for API parity (number of 32-splat buckets) */
  int raw_params;              /* !=0 (extension): opacities / scales / rotations are the
                                  model's RAW parameters; the callee applies sigmoid / exp / normalize
                                  (GaussianModel::get*Activation, src/gaussian_model.cpp:48-71) itself */
  void* radii_ready_event;     /* optional cudaEvent_t, recorded on `stream` right after the kernel
                                  that writes `radii` (lets a side stream start gsr_adam_step_rows
                                  phase 0 while the rest of the forward pass runs); may be NULL */
  void* context;               /* gsr_context* or NULL (default: one per host thread and device) */
} gsr_forward_args;

typedef struct gsr_backward_args {
  int P, D, M, R, B, W, H;
  const float* background;
  const float* means3D;
  const float* dc;
  const float* shs;
  const float* colors_precomp;
  const float* opacities;      /* unused by the maths; kept for symmetry (may be NULL) */
  const float* scales;
  float scale_modifier;
  const float* rotations;
  const float* cov3D_precomp;
  const float* viewmatrix;
  const float* projmatrix;
  const float* campos;
  float tan_fovx, tan_fovy;
  const int* radii;
  char* geom_buffer;           /* the four buffers filled by gsr_forward */
  char* binning_buffer;
  char* image_buffer;
  char* sample_buffer;
  const float* dL_dpix;        /* [3,H,W] */
  /* outputs: fully written by the callee (no pre-zeroing needed) */
  float* dL_dmean2D;           /* [P,3] */
  float* dL_dconic;            /* [P,4] or NULL */
  float* dL_dopacity;          /* [P] */
  float* dL_dcolor;            /* [P,3] */
  float* dL_dmean3D;           /* [P,3] */
  float* dL_dcov3D;            /* [P,6] or NULL */
  float* dL_ddc;               /* [P,3] */
  float* dL_dsh;               /* [P,M,3] (may be NULL when M==0) */
  float* dL_dscale;            /* [P,3] */
  float* dL_drot;              /* [P,4] */
  int debug;
  void* stream;
  int raw_params;              /* !=0: the forward ran with raw_params; `opacities` (raw, required),
                                  `scales`, `rotations` are raw and dL_dopacity / dL_dscale / dL_drot are
                                  returned w.r.t. the RAW parameters (chain rule through the activations) */
  int rendered_rows_only;      /* !=0: write only the rows of the rendered Gaussians (radii > 0 and at
                                  least one tile) and leave every other row of the outputs UNTOUCHED
                                  instead of zero-filling it — for callers that consume the gradients
                                  through the radii mask (gsr_adam_step_rows); saves the dense zero
                                  stream (276 B per Gaussian) */
  int accumulate;              /* !=0 (implies rendered_rows_only): the six parameter gradients — dL_dmean3D,
                                  dL_ddc, dL_dsh, dL_dopacity, dL_dscale, dL_drot — are ADDED to what the
                                  rendered rows already hold, i.e. a second backward() without zero_grad
                                  (the keyframe-batch sum of SURVEY.md §8e) without a scratch arena and a
                                  dense add.  The per-view outputs (dL_dmean2D, dL_dconic, dL_dcolor,
                                  dL_dcov3D) are overwritten as usual. */
  const int32_t* accumulate_rows; /* optional with accumulate, [P]: only the rows with a value > 0 (written by an
                                  earlier keyframe of the batch, e.g. the running maximum of their radii) are
                                  added to, the other rendered rows are overwritten — so NO keyframe of a
                                  batch needs the dense zero fill; NULL = add to every rendered row */
  int32_t* grad_rows;          /* optional, [P], in/out — the row mask of a keyframe batch kept by the callee
                                  itself (the caller zeroes it at the start of the batch):
                                    in : with `accumulate`, the rows with grad_rows > 0 are added to and the other
                                         rendered rows overwritten (takes precedence over accumulate_rows);
                                    out: grad_rows[row] = 1 for every rendered row that RECEIVED a gradient (some
                                         pixel blended it).  A rendered row that received none (a splat hidden
                                         behind opaque ones: all of its gradients are exactly zero — 85 % of the
                                         rendered rows of a 3 M-Gaussian room view) is not flagged, and with
                                         rendered_rows_only its six PARAMETER gradient rows are not written at all
                                         (the per-view outputs dL_dmean2D / dL_dconic / dL_dcolor are, as zeros).
                                  Consumers that go through this mask — gsr_adam_step_rows(row_mask = grad_rows),
                                  gsr_peer_local_stats(..., grad_rows) -> gsr_adam_step_peer — treat every other row
                                  as zero gradient, which is what it is. */
} gsr_backward_args;

int gsr_forward(const gsr_forward_args* a);
int gsr_backward(const gsr_backward_args* a);

int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);

/* mean of the squared distances to the 3 nearest neighbours; points [P,3]. */
int gsr_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* stream);

/* One fused multi-tensor Adam step (torch::optim::Adam semantics, no amsgrad,
 * no weight decay).  Bias corrections are computed in double on the host from
 * `step` exactly like LibTorch does. */
typedef struct gsr_adam_group {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
  double lr;
  int64_t step;  /* step count AFTER this update (>=1) */
} gsr_adam_group;
int gsr_adam_step(const gsr_adam_group* groups, int n_groups, double beta1,
                  double beta2, double eps, float grad_scale, void* stream);

/* The same step, split by row visibility so that most of it can overlap the rasterizer.
 * Every group is a [rows, width] tensor (numel % rows == 0) and row_mask[r] > 0 marks the rows
 * that received a gradient in this iteration (the `radii` output of gsr_forward, or the
 * element-wise maximum of several keyframes' / ranks' radii).
 *   phase GSR_ADAM_ROWS_ZERO_GRAD: updates the rows with row_mask <= 0, whose gradient is
 *     exactly +0 (the reference's dense torch::optim::Adam still decays their moments and moves
 *     them); `grad` is not read, so this phase may run on a side stream as soon as the forward
 *     pass has produced radii, concurrently with the blend / backward kernels;
 *   phase GSR_ADAM_ROWS_WITH_GRAD: updates the rows with row_mask > 0 from `grad`.
 * Running both phases (in either order) is bit-identical to one gsr_adam_step whenever the
 * masked-out rows of `grad` hold +0, which gsr_backward guarantees. */
#define GSR_ADAM_ROWS_ZERO_GRAD 0
#define GSR_ADAM_ROWS_WITH_GRAD 1
/* phase GSR_ADAM_ROWS_ALL: one pass over every row; `grad` is read only where row_mask > 0 and
 * taken as +0 elsewhere, so the masked-out rows of `grad` may hold anything (see
 * gsr_backward_args.rendered_rows_only).  Same result as gsr_adam_step on a zero-filled grad. */
#define GSR_ADAM_ROWS_ALL 2
int gsr_adam_step_rows(const gsr_adam_group* groups, int n_groups, double beta1,
                       double beta2, double eps, float grad_scale, int64_t rows,
                       const int32_t* row_mask, int phase, void* stream);

/* ---- fused L1 + SSIM photometric loss, forward + backward ----
 * Replaces loss_utils::l1_loss + loss_utils::ssim and their autograd backward as used by
 * GaussianMapper::trainForOneIteration (include/loss_utils.h:28-124, src/gaussian_mapper.cpp:414-431):
 *   masked = where(mask > 0, image, 1);  loss = (1 - l) * mean|masked - where(mask > 0, gt, 1)|
 *                                             + l * (1 - SSIM(masked, gt))      (11x11, sigma 1.5)
 * image, gt, dL_dimage: [3,H,W]; mask: [H,W] or NULL; loss_out: device [3] {loss, L1, SSIM} or NULL.
 * dL_dimage = d loss / d image, ready to be passed to gsr_backward as dL_dpix.
 * workspace: gsr_l1_ssim_workspace_bytes(H, W) bytes of device memory. */
size_t gsr_l1_ssim_workspace_bytes(int H, int W);
int gsr_l1_ssim_loss(const float* image, const float* gt, const float* mask, int H, int W,
                     float lambda_dssim, float* loss_out, float* dL_dimage, char* workspace,
                     void* stream);

/* ---- multi-GPU optimiser step over NVLink peer memory (one node, N <= GSR_MAX_PEERS ranks) ----
 * Fuses the gradient reduction, the Adam update and the parameter broadcast of a data-parallel
 * mapping step into one kernel: rank `rank` owns the element range [shard_lo, shard_hi) of a flat
 * arena, sums the N peers' gradient arenas over it (remote loads, skipped for rows whose
 * union_mask is 0), updates ITS copy of the parameters with shard-local Adam moments and stores the
 * result into every rank's parameter arena (remote stores).  All arenas share one layout, described
 * by `segments` (ascending, 4-element aligned starts; `width` = elements per Gaussian row, `lr` per
 * segment).  param_peers / grad_peers are the N peer-mapped base pointers of the arenas (e.g. from
 * torch.distributed._symmetric_memory).  The caller must bracket the call with cross-rank barriers:
 * all gradients written before, all ranks' kernels finished before parameters are read again.
 * Replaces torch::optim::Adam::step + the gradient all-reduce a multi-GPU port of
 * src/gaussian_mapper.cpp:465 would need. */
#define GSR_MAX_PEERS 8
#define GSR_PEER_MAX_SEGMENTS 12
#define GSR_PEER_RULE_ADAM 0   /* param <- Adam(param, sum over ranks of grad), the row's GRADIENT mask bits */
#define GSR_PEER_RULE_SUM 1    /* param <- param + sum over ranks of grad (rank order), the row's RENDERED mask bits */
#define GSR_PEER_RULE_MAX 2    /* param <- max(param, max over ranks of grad),           the row's RENDERED mask bits */
typedef struct gsr_peer_segment {
  int64_t start;   /* first arena element */
  int64_t numel;
  int32_t width;   /* elements per row */
  int32_t rule;    /* GSR_PEER_RULE_*: SUM / MAX carry the densification statistics of a keyframe batch through the
                      same kernel as three one-float segments — "param" arena = the replicated statistic, "grad" arena
                      = this rank's per-step delta (gsr_peer_local_stats); they have no Adam moments */
  double lr;
} gsr_peer_segment;
typedef struct gsr_peer_adam_args {
  int n_peers, rank;
  float* param_peers[GSR_MAX_PEERS];
  const float* grad_peers[GSR_MAX_PEERS];
  int n_segments;
  gsr_peer_segment segments[GSR_PEER_MAX_SEGMENTS];
  int64_t shard_lo, shard_hi;   /* multiples of 4 elements */
  float* exp_avg;               /* [shard_hi - shard_lo], local */
  float* exp_avg_sq;
  const int32_t* union_mask;    /* [rows]: bit k <=> rank k holds a valid GRADIENT row, bit 8 + k <=> rank k RENDERED
                                   the row (gsr_peer_rank_mask_u8; gsr_peer_union_mask sets both alike) */
  double beta1, beta2, eps;
  float grad_scale;
  int64_t step;                 /* step count AFTER this update (>= 1) */
  void* stream;
  /* 0: the rank owns the contiguous range [shard_lo, shard_hi).  > 0 (a multiple of 4): interleaved
   * sharding — the arena [0, shard_hi) is cut into chunks of shard_chunk elements and rank r owns the
   * chunks c with c % n_peers == r (shard_lo must be 0); exp_avg / exp_avg_sq hold the owned chunks back
   * to back.  Every rank then gets the same mix of narrow and wide parameter rows, i.e. the same work. */
  int64_t shard_chunk;
  /* Optional MULTICAST address of the parameter arena (NVLS: one mapping that the NVSwitch replicates into the
   * arena of every rank, e.g. torch symmetric memory's multicast_ptr).  When non-NULL the new parameters leave
   * the SM as ONE multimem store per 16 bytes instead of n_peers unicast stores: 1/n_peers of the store
   * instructions and of the NVLink egress.  NULL: unicast stores to param_peers[0..n_peers). */
  float* param_multicast;
  /* Optional (device, two uint64 the caller zeroes): measurement of the kernel's NVLink traffic.  [0] += 16-byte
   * gradient loads issued to OTHER ranks' arenas, [1] += 16-byte parameter groups that changed and were stored
   * (each leaves as n_peers - 1 remote stores, or one multicast store).  NULL: nothing is counted (the timed path). */
  unsigned long long* traffic_counters;
} gsr_peer_adam_args;
int gsr_adam_step_peer(const gsr_peer_adam_args* a);
/* union_out[row] = bitmask of the ranks that rendered the row: bit k set <=> radii_peers[k][row] > 0
 * (peer-mapped radii buffers).  gsr_adam_step_peer reads rank k's gradient of a row only when bit k is
 * set, so a rank may leave the rows it did not render unwritten (gsr_backward rendered_rows_only). */
int gsr_peer_union_mask(const int32_t* const* radii_peers, int n_peers, int64_t rows,
                        int32_t* union_out, void* stream);
/* The same mask from one-byte flags (flags_peers[k][row]: bit 0 <=> rank k rendered the row, bit 1 <=> its gradient
 * row is valid — gsr_peer_local_stats; 4-byte aligned peer-mapped buffers, mask_out 16-byte aligned): a quarter of the
 * NVLink traffic of the radii version.  mask_out[row] = gradient bits | rendered bits << 8. */
int gsr_peer_rank_mask_u8(const uint8_t* const* flags_peers, int n_peers, int64_t rows,
                          int32_t* mask_out, void* stream);

/* Densification statistics of a keyframe batch split over several ranks (the reference's
 * per-view update is GaussianMapper::trainForOneIteration, src/gaussian_mapper.cpp:443-447, and
 * GaussianModel::addDensificationStats, src/gaussian_model.cpp:667-681).
 * gsr_peer_local_stats: after each LOCAL keyframe's backward pass; for the rows with radii > 0 it sets
 *   flags[row] = 1 (rendered) | 2 (its gradient row is valid: grad_rows[row] > 0, or grad_rows == NULL — see
 *   gsr_backward_args.grad_rows; only those rows are pulled over NVLink by gsr_adam_step_peer) and, when stats_local != NULL, adds that view's terms to the rank's per-step deltas
 *   stats_local = [accum | denom | max_radii], three arrays of stats_stride >= P floats each (stats_stride a
 *   multiple of 4, base 16-byte aligned; the caller zeroes flags — P + 16 bytes, 4-byte aligned — and stats_local
 *   at the start of the step; both live in peer-mapped memory).
 * The mappers of this repository then carry the statistics THROUGH gsr_adam_step_peer: stats_local is the tail of
 *   the rank's gradient arena, the replicated statistics the tail of its parameter arena, described by three
 *   one-float segments with GSR_PEER_RULE_SUM / _SUM / _MAX — each owner reduces its rows and broadcasts the result
 *   (rank order, bit-identical replicas), 1/n_peers of the traffic of every rank pulling every rank's deltas.
 * gsr_peer_stats_exchange is that all-pull alternative for callers that keep the statistics outside the arenas:
 *   after the cross-rank barrier, mask_out[row] = gradient bits | rendered bits << 8 (what gsr_adam_step_peer
 *   consumes) and the replicated statistics of THIS rank receive xyz_gradient_accum += sum_k accum_k,
 *   denom += sum_k denom_k, max_radii2D = max(max_radii2D, max_k ...) over the ranks that rendered the row, in
 *   rank order.  (Measured at 3 M Gaussians: 0.10 ms at 2 ranks, 0.41 ms at 8; the fused form costs 0.05 ms.) */
int gsr_peer_local_stats(int64_t P, const int32_t* radii, const float* dL_dmeans2D, float* stats_local,
                         int64_t stats_stride, uint8_t* flags, const int32_t* grad_rows, void* stream);
int gsr_peer_stats_exchange(const uint8_t* const* flags_peers, const float* const* stats_peers, int64_t stats_stride,
                            int n_peers, int64_t P, int32_t* mask_out, float* max_radii2D, float* xyz_gradient_accum,
                            float* denom, void* stream);

/* ---- Gaussian-store surgery: densify / prune / increasePcd ----
 * The reference keeps the map as six parameter tensors, twelve Adam moment tensors and four
 * per-Gaussian statistics, and re-creates ALL of them with index / cat / zeros_like on every
 * densification (src/gaussian_model.cpp:431-665) and every new frame (increasePcd, :201-317).
 * Here the store is one struct of row arrays with spare capacity and every reference method is one
 * pass: classify each source row, scan, scatter every surviving / new row to its final place.
 * Row order, values and Adam state of the result equal the reference's op sequence (bit-exact for
 * copied rows; see tests/test_gpu_densify.py for the tolerance of the computed ones). */
#define GSR_STORE_GROUPS 6
typedef struct gsr_gaussian_store {
  int64_t rows;       /* P: Gaussians held */
  int64_t capacity;   /* rows every array can hold */
  int32_t n_rest;     /* M: features_rest_ is [P,M,3] */
  /* xyz [P,3], features_dc [P,1,3], features_rest [P,M,3], opacity [P,1], scaling [P,3],
   * rotation [P,4]: the reference's param_groups order (src/gaussian_model.cpp:325-345) */
  float* param[GSR_STORE_GROUPS];
  float* exp_avg[GSR_STORE_GROUPS];     /* Adam moments, same shapes; NULL = no state yet */
  float* exp_avg_sq[GSR_STORE_GROUPS];
  int32_t* exist_since_iter;            /* [P] */
  float* xyz_gradient_accum;            /* [P,1] */
  float* denom;                         /* [P,1] */
  float* max_radii2D;                   /* [P] */
} gsr_gaussian_store;

/* Bytes of device scratch the store operations below need for a store of `rows` rows. */
size_t gsr_store_workspace_bytes(int64_t rows);

/* GaussianModel::prunePoints (src/gaussian_model.cpp:431-490): dst = the rows of src whose
 * prune_mask byte is 0, in order — parameters, both Adam moments and the four statistics.
 * src and dst must not overlap.  counts_out_host[0] = rows kept (the call waits for it, like the
 * reference's boolean-mask index does); dst->rows is NOT written (dst is const: the caller sets it). */
int gsr_prune_points(const gsr_gaussian_store* src, const gsr_gaussian_store* dst,
                     const uint8_t* prune_mask, char* workspace, int64_t* counts_out_host,
                     void* stream);

/* GaussianModel::densifyAndPrune (src/gaussian_model.cpp:644-665) = densifyAndClone (:614-642) +
 * densifyAndSplit (:571-612, N copies, originals removed) + prunePoints(opacity / size mask), fused.
 * Resulting row order = the reference's: surviving originals, surviving clones, surviving split
 * copies (copy 0 of every selected row, then copy 1, ...).  New rows get zero Adam moments; the three
 * statistics of EVERY row are zero afterwards (densificationPostfix resets them, :566-568), which also
 * makes the reference's max_radii2D > max_screen_size test vacuous — kept that way.
 * Split positions: xyz + R(rotation) * (z * exp(scaling)), z ~ N(0,1).  `noise` supplies z explicitly:
 * row (c * S + j) of noise[., 3] belongs to copy c of the j-th split-selected Gaussian (S = number
 * selected — the layout at::normal sees in the reference, so a caller that does not know S yet passes
 * n_split * rows rows and the first n_split * S are used); with noise == NULL z is drawn on the
 * device from Philox-4x32-10 keyed by (seed, c * S + j).
 * dst must not overlap src.  If dst->capacity is below the worst case (2 * rows, or n_split * rows)
 * the real size is checked before anything is written: GSR_ERR_ALLOC with counts_out_host[0] = rows
 * needed means "grow dst and call again".
 * counts_out_host[4] = {rows in dst, originals kept, clones kept, split-selected S} (waited for). */
typedef struct gsr_densify_args {
  float max_grad;         /* densify_grad_threshold; must be > 0 */
  float min_opacity;
  float extent;           /* scene extent (cameras_extent_) */
  float percent_dense;
  int32_t max_screen_size;/* 0 = skip the size tests (src/gaussian_model.cpp:655-660) */
  int32_t n_split;        /* N, the reference uses 2; 1..4 */
  const float* noise;     /* [noise_rows, 3] standard normal, or NULL */
  uint64_t seed;
  int64_t noise_rows;     /* rows `noise` holds; draws beyond them come from the device generator */
} gsr_densify_args;
int gsr_densify_and_prune(const gsr_gaussian_store* src, const gsr_gaussian_store* dst,
                          const gsr_densify_args* args, char* workspace, int64_t* counts_out_host,
                          void* stream);

/* GaussianModel::densificationPostfix (src/gaussian_model.cpp:492-569) for rows that already exist
 * as arrays: appends n_new rows (new_param[g] in the store's group order, new_exist_since_iter) after
 * store->rows, zero Adam moments for them, and zeroes the three statistics of all rows.  Needs
 * store->rows + n_new <= store->capacity. */
int gsr_densification_postfix(const gsr_gaussian_store* store, const float* const* new_param,
                              const int32_t* new_exist_since_iter, int64_t n_new, void* stream);

/* The per-iteration bookkeeping of GaussianMapper::trainForOneIteration (src/gaussian_mapper.cpp:
 * 443-447) + GaussianModel::addDensificationStats (src/gaussian_model.cpp:667-681), one pass: for
 * rows with radii > 0: max_radii2D = max(max_radii2D, radii); xyz_gradient_accum += |dL_dmeans2D.xy|;
 * denom += 1.  dL_dmeans2D is [P,3] (z unused). */
int gsr_add_densification_stats(int64_t P, const int32_t* radii, const float* dL_dmeans2D,
                                float* max_radii2D, float* xyz_gradient_accum, float* denom,
                                void* stream);

/* GaussianModel::resetOpacity (src/gaussian_model.cpp:399-408) incl. replaceTensorToOptimizer
 * (:410-429): opacity = inverse_sigmoid(min(sigmoid(opacity), cap)), both Adam moments of the group
 * zeroed.  NOTE the reference builds its cap as ones_like(opacity * 0.01) == 1, i.e. cap = 1 (the
 * values only make a sigmoid / logit round trip); pass cap = 1.0f for reference behaviour, 0.01f
 * for what the original 3DGS code does. */
int gsr_reset_opacity(int64_t P, float* opacity, float* exp_avg, float* exp_avg_sq, float cap,
                      void* stream);

/* The voxel-hash duplicate test of GaussianModel::increasePcd (src/gaussian_model.cpp:214-232):
 * keep_out[i] = 0 if a Gaussian of `existing` lies in the same voxel hash
 * (int64 floor(x/v)*73856093 + floor(y/v)*19349663 + floor(z/v)*83492791) as new point i, else 1.
 * Replaces torch::isin on two int64 key tensors by a device hash set.
 * workspace: gsr_voxel_dedup_workspace_bytes(n_existing). */
size_t gsr_voxel_dedup_workspace_bytes(int64_t n_existing);
int gsr_voxel_dedup(const float* existing_xyz, int64_t n_existing, const float* new_xyz,
                    int64_t n_new, float voxel_size, uint8_t* keep_out, char* workspace,
                    void* stream);

/* Compacted ascending list of the rows whose mask byte is non-zero (invert: zero):
 * idx_out[0..count), count written to the device word count_out_dev and, if counts_out_host is
 * non-NULL, waited for and stored there.  (nonzero().squeeze() of src/gaussian_model.cpp:233.)
 * workspace: gsr_store_workspace_bytes(rows). */
int gsr_select_rows(const uint8_t* mask, int64_t rows, int invert, uint32_t* idx_out,
                    uint32_t* count_out_dev, char* workspace, int64_t* counts_out_host, void* stream);

/* The new-Gaussian initialisation of increasePcd (src/gaussian_model.cpp:235-290) for the points
 * selected by idx[0..n): gathers xyz / colour, features_dc = RGB2SH(colour), features_rest = 0,
 * scaling = log(sqrt(max(dist2, 1e-7))) (dist2 = gsr_knn_mean_dist2 of the gathered points),
 * rotation = (1,0,0,0), opacity = inverse_sigmoid(0.1), exist_since_iter = iteration — written
 * straight into rows [store->rows, store->rows + n) of the store with zero Adam moments, and the
 * statistics of all rows zeroed (the densificationPostfix at :305).  Two calls: stage 0 gathers the
 * points into the store's xyz rows (so that the caller can run gsr_knn_mean_dist2 on them), stage 1
 * needs dist2 and fills the rest. */
int gsr_increase_pcd_rows(const gsr_gaussian_store* store, const float* new_xyz,
                          const float* new_rgb, const uint32_t* idx, int64_t n, int stage,
                          const float* dist2, int32_t iteration, void* stream);

/* PLY interchange rows (GaussianModel::savePly / loadPly, src/gaussian_model.cpp:688-953): one vertex =
 * x y z nx ny nz f_dc_0..2 f_rest_0..3M-1 opacity scale_0..2 rot_0..3 = 17 + 3M little-endian floats, normals
 * zero, f_rest channel-major (file index ch*M + m holds features_rest[m][ch]).  pack builds the [P, 17+3M]
 * rows in device memory from the model's arrays, unpack is the inverse (normals ignored). */
int gsr_ply_pack_rows(int64_t P, int n_rest, const float* xyz, const float* features_dc,
                      const float* features_rest, const float* opacity, const float* scaling,
                      const float* rotation, float* rows, void* stream);
int gsr_ply_unpack_rows(int64_t P, int n_rest, const float* rows, float* xyz, float* features_dc,
                        float* features_rest, float* opacity, float* scaling, float* rotation,
                        void* stream);

/* ---- frame ingest ----
 * colorize_launcher (cuda_colorize/colorize_pointcloud.cu:57-107): every LiDAR point is taken to the
 * camera (Tr_velo_to_cam, row-major 4x4, rows 0..2 used) and projected (proj_mat, row-major 3x4); points
 * with camera z <= 0.2 or outside the image are dropped; the survivors are moved to the world frame
 * (lidar_pose, row-major 4x4) and take the colour of their pixel (image: [H,W,3] float).  Outputs are the
 * COMPACTED survivors in input order: points_out / colors_out [count,3], both with room for n rows.
 * The count is waited for and stored in *count_out_host when that is non-NULL.
 * workspace: gsr_colorize_workspace_bytes(n). */
/* The point tensor of a LiDAR frame (submodules/data_loader/data_loader.cpp:172-179: the reference fills a CUDA
 * tensor with one `tensor[i][j] = value` assignment per coordinate, i.e. 3 n tiny device operations): `records` is
 * the scan as read from disk and copied to the device in ONE transfer — n records of stride_floats floats whose
 * first three are x, y, z (KITTI .bin: 4) — and xyz_out [n,3] the dense point tensor colorize consumes. */
int gsr_pack_xyz(const float* records, int64_t n, int stride_floats, float* xyz_out, void* stream);
size_t gsr_colorize_workspace_bytes(int64_t n);
int gsr_colorize_points(const float* points, int64_t n, const float* lidar_pose, const float* proj_mat,
                        const float* Tr_velo_to_cam, const float* image, int img_width, int img_height,
                        float* points_out, float* colors_out, char* workspace, int64_t* count_out_host,
                        void* stream);

/* ---- test / debug exports (device -> device copies out of the opaque buffers) ---- */
/* Any output pointer may be NULL.  keys are reconstructed 64-bit (tile<<32 | depth bits).
 * gsr_export_binning returns the COMPLETE depth-sorted per-tile lists (what the reference's radix sort produces):
 * gsr_forward itself orders and packs, of a list of 512 or more entries, only the prefix the blend kernels read
 * (see gsr_export_redo_count), so the export first finishes the sort of those lists on `stream` (the packed
 * records keep their prefix; the geometry buffer is needed for that). */
int gsr_export_binning(const char* geom_buffer, const char* binning_buffer,
                       const char* image_buffer, int P, int R, int W, int H,
                       uint64_t* keys_sorted, uint32_t* point_list, void* stream);
/* How many tiles of the last gsr_forward took the full sort + the second blend launch (which continues their half
 * tiles) because the depth-sorted PREFIX of their list (lists of 512+ entries are ordered only as far as the blend is
 * expected to read) did not saturate their pixels; host value, synchronises the stream.  0 also when the context
 * had the prefix sort switched off. */
int gsr_export_redo_count(const char* image_buffer, int W, int H, uint32_t* count_host, void* stream);
int gsr_export_image(const char* image_buffer, int W, int H, float* final_T,
                     uint32_t* n_contrib, uint32_t* ranges /*[T,2]*/,
                     uint32_t* max_contrib /*[T]*/, void* stream);
/* Measurement aid (bench.py's roofline numerators): replays the forward blend of a finished gsr_forward with an
 * instrumented kernel and returns counts_out[3] (device) = (splat, pixel) pairs {visited while the pixel was not
 * done, passing the power gate (exp evaluated), blended}.  scratch_color: [3,H,W] floats, overwritten. */
int gsr_export_pair_counts(const char* binning_buffer, char* image_buffer, int R, int W, int H,
                           const float* background, float* scratch_color, uint64_t* counts_out, void* stream);
int gsr_export_geometry(const char* geom_buffer, int P, float* depths,
                        float* means2D, float* conic_opacity, float* rgb,
                        uint32_t* tiles_touched, uint8_t* clamped, void* stream);

/* Stand-alone stable LSD onesweep radix sort of 64-bit items on bits
 * [begin_bit, end_bit) (exposed for unit tests and benchmarks).
 * `tmp` must hold gsr_sort_temp_bytes(n) bytes. Result ends in `items_out`. */
size_t gsr_sort_temp_bytes(int64_t n);
int gsr_sort_u64(const uint64_t* items_in, uint64_t* items_out, uint64_t* items_alt,
                 int64_t n, int begin_bit, int end_bit, char* tmp, void* stream);

/* Optional per-stage CUDA-event timing (bench.py roofline measurement): enable, run steps,
 * then read "stage count total_ms" lines.  Off by default; adds two event records per stage. */
void gsr_profile_enable(int on);
int gsr_profile_read(char* out, size_t cap);

/* Number of kernels this library has launched since load (bench bookkeeping). */
int64_t gsr_launch_count(void);
const char* gsr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GSR_CAPI_H_ */
