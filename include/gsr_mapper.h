/*
 * gsr_mapper.h — C entry for the keyframe-batch mapping step (single- and multi-GPU) of the B200-native hot path.
 *
 * What it replaces: the body of GaussianMapper::trainForOneIteration between "pick a keyframe" and
 * "optimizer_->step()" (/root/reference/src/gaussian_mapper.cpp:372-466) — render, loss, backward,
 * densification statistics (:443-447 + GaussianModel::addDensificationStats, src/gaussian_model.cpp:667-681), Adam
 * (src/gaussian_model.cpp:319-352) — for a BATCH of keyframes that is split over `world` ranks, one process (or
 * thread) per GPU on one NVLink node.  The semantics are the ones pinned there: batch gradient =
 * sum over the keyframes of all ranks, one Adam step per batch, statistics of every view, replicas bit-identical.
 *
 * A C++ gaussian_mapper.cpp needs nothing but this header, libgsr.so and a way to all-gather
 * GSR_MAPPER_HANDLE_BYTES bytes per rank once at start-up (MPI_Allgather, ncclAllGather on a staging buffer, a
 * file, a socket ...):
 *
 *     gsr_mapper_config cfg = {...};                       // P, M, rank, world, learning rates
 *     char mine[GSR_MAPPER_HANDLE_BYTES], all[8 * GSR_MAPPER_HANDLE_BYTES];
 *     gsr_mapper* m = gsr_mapper_create(&cfg, mine);       // allocates this rank's arenas (peer-shareable)
 *     my_allgather(mine, all, GSR_MAPPER_HANDLE_BYTES);    // the caller's transport
 *     gsr_mapper_connect(m, all);                          // maps the peers' arenas (CUDA IPC over NVLink)
 *     cudaMemcpy(gsr_mapper_param(m, GSR_SEG_XYZ), xyz, ...);   // initial parameters, identical on every rank
 *     for (...) gsr_mapper_step(m, my_keyframes, n_mine, background, 1.0f, stream);
 *
 * The exchange is the same fused peer-memory kernel the Python mapper uses (gsr_adam_step_peer: gradient reduction
 * over the ranks that rendered a row + Adam on an interleaved 1/world shard + parameter broadcast), bracketed by
 * two on-device cross-GPU barriers; no NCCL call is on the path.  world == 1 uses the masked single-GPU Adam.
 * All functions return 0 / a pointer on success, a negative gsr error code or a positive cudaError_t otherwise.
 */
#ifndef GSR_MAPPER_H_
#define GSR_MAPPER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_MAPPER_HANDLE_BYTES 256
#define GSR_MAPPER_MAX_RANKS 8

/* parameter segments, in the order they lie in the flat arena */
enum { GSR_SEG_XYZ = 0, GSR_SEG_FDC = 1, GSR_SEG_OPACITY = 2, GSR_SEG_SCALING = 3, GSR_SEG_ROTATION = 4, GSR_SEG_FREST = 5 };
enum { GSR_STAT_MAX_RADII2D = 0, GSR_STAT_XYZ_GRADIENT_ACCUM = 1, GSR_STAT_DENOM = 2 };

typedef struct gsr_mapper gsr_mapper;

typedef struct gsr_mapper_config {
  int64_t P;            /* Gaussians (fixed for the life of the mapper; densification = destroy + create) */
  int M;                /* "rest" SH coefficients per Gaussian (0 or 15) */
  int sh_degree;        /* active degree */
  int rank, world;      /* world <= GSR_MAPPER_MAX_RANKS; the current CUDA device is this rank's GPU */
  double lr[6];         /* per segment (GSR_SEG_*) */
  double beta1, beta2, eps;   /* the reference: 0.9, 0.999, 1e-15 */
  int raw_params;       /* !=0: opacity / scaling / rotation segments hold the model's RAW parameters (row f-2) */
  int with_stats;       /* !=0: maintain the densification statistics */
  float lambda_dssim;   /* weight of the SSIM term when a keyframe carries a ground-truth image */
} gsr_mapper_config;

typedef struct gsr_mapper_keyframe {
  const float* viewmatrix;   /* [16] device, transposed like everywhere in this API */
  const float* projmatrix;   /* [16] device */
  const float* campos;       /* [3] device */
  float tanfovx, tanfovy;
  int W, H;
  const float* dL_dpix;      /* [3,H,W] device: dL/d(rendered image), or NULL ... */
  const float* gt_image;     /* ... then [3,H,W] ground truth for the fused L1 + SSIM loss */
  const float* mask;         /* [H,W] or NULL */
  float* out_color;          /* [3,H,W] device: receives the rendered image, or NULL */
} gsr_mapper_keyframe;

/* Allocates this rank's buffers on the current device and writes the blob its peers need to `handle_out`
 * (GSR_MAPPER_HANDLE_BYTES; ignored when world == 1, may then be NULL). */
gsr_mapper* gsr_mapper_create(const gsr_mapper_config* cfg, void* handle_out);
/* all_handles: world blobs in rank order (own blob included).  No-op for world == 1. */
int gsr_mapper_connect(gsr_mapper* m, const void* all_handles);
/* Device pointers into this rank's replica: parameter segment [P * width] and statistic [P] (NULL without stats). */
float* gsr_mapper_param(gsr_mapper* m, int segment);
float* gsr_mapper_stat(gsr_mapper* m, int which);
/* learning-rate schedule (GaussianModel::updateLearningRate, src/gaussian_model.cpp:354-396) */
int gsr_mapper_set_lr(gsr_mapper* m, int segment, double lr);
/* One optimiser step over this rank's `n` keyframes of the batch (n may be 0).  Every rank must call it once per
 * batch.  loss_out (optional, device [3]): {loss, L1, SSIM} of the last keyframe that carried a gt_image. */
int gsr_mapper_step(gsr_mapper* m, const gsr_mapper_keyframe* keyframes, int n, const float* background,
                    float grad_scale, float* loss_out, void* stream);
int64_t gsr_mapper_steps_done(const gsr_mapper* m);
void gsr_mapper_destroy(gsr_mapper* m);

#ifdef __cplusplus
}
#endif
#endif /* GSR_MAPPER_H_ */
