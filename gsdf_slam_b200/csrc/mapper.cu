// C entry for the keyframe-batch mapping step (include/gsr_mapper.h): the host-side orchestration that a C++
// gaussian_mapper.cpp needs for the multi-GPU step — symmetric (peer-mapped) arenas set up with CUDA IPC, the
// on-device cross-GPU barrier, the per-keyframe forward / loss / backward / statistics calls and the fused
// peer-memory exchange — with no Python and no NCCL call on the path.  Everything here only sequences entry
// points of gsr_capi.h on the caller's stream.
#include "gsr_mapper.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "gsr_capi.h"
#include "gsr_common.cuh"

namespace {

constexpr int kSegs = 6;
constexpr int64_t kShardChunk = 65536;   // interleaved sharding of the arena (see gsr_adam_step_peer)

#define MAP_CUDA(x)                         \
  do {                                      \
    cudaError_t _e = (x);                   \
    if (_e != cudaSuccess) return (int)_e;  \
  } while (0)
#define MAP_RC(x)            \
  do {                       \
    int _r = (x);            \
    if (_r != 0) return _r;  \
  } while (0)

struct DevBuf {   // growable device buffer behind a gsr_resize_fn
  char* p = nullptr;
  size_t cap = 0;
  static char* resize(void* user, size_t bytes) {
    DevBuf* b = static_cast<DevBuf*>(user);
    if (bytes > b->cap) {
      if (b->p) cudaFree(b->p);   // (synchronises; happens only while the buffers still grow)
      b->p = nullptr;
      const size_t want = bytes + bytes / 4 + 4096;
      if (cudaMalloc(reinterpret_cast<void**>(&b->p), want) != cudaSuccess) {
        b->cap = 0;
        return nullptr;
      }
      b->cap = want;
    }
    return b->p;
  }
  void free() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// what one rank publishes to its peers
struct HandleBlob {
  uint32_t magic;
  int32_t rank;
  uint64_t bytes;
  cudaIpcMemHandle_t mem;
  int32_t pid;
};
static_assert(sizeof(HandleBlob) <= GSR_MAPPER_HANDLE_BYTES, "handle blob too large");

// Cross-GPU barrier on signal pads in peer memory: thread k posts `epoch` to slot [rank] of peer k's pad (release,
// system scope) and then waits until slot [k] of the own pad has reached `epoch`.  Stream order puts it behind the
// kernels whose writes it publishes; one rank per GPU, so the waiting kernels always run concurrently.
struct PadPtrs {
  uint32_t* p[GSR_MAPPER_MAX_RANKS];
};
__global__ void peer_barrier_kernel(PadPtrs pads, int rank, int world, uint32_t epoch) {
  const int k = threadIdx.x;
  if (k >= world) return;
  __threadfence_system();
  uint32_t* remote = pads.p[k] + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const uint32_t* mine = pads.p[rank] + k;
  uint32_t v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
  } while ((int32_t)(v - epoch) < 0);
}

}  // namespace

struct gsr_mapper {
  gsr_mapper_config cfg;
  int device = -1;
  int64_t P = 0;
  int widths[kSegs] = {3, 3, 1, 3, 4, 0};
  int64_t seg_off[kSegs] = {0}, seg_n[kSegs] = {0};
  int64_t total = 0;          // floats of the six parameter segments
  int64_t arena = 0;          // floats per arena: total + the three statistics segments [accum | denom | max_radii]
  int64_t stats_stride = 0;   // floats between the statistics segments
  double lr[kSegs];
  // symmetric block of this rank: [param arena (+ replicated statistics) | gradient arena (+ this rank's statistics
  // deltas of the step) | flags P+16 | signal pad]
  char* sym = nullptr;
  size_t sym_bytes = 0;
  size_t off_param = 0, off_grad = 0, off_flags = 0, off_pad = 0;
  char* peer[GSR_MAPPER_MAX_RANKS] = {nullptr};   // base of every rank's block as mapped here (own: sym)
  bool connected = false;
  uint32_t epoch = 0;
  // local state
  float* exp_avg = nullptr;     // world > 1: the owned chunks; world == 1: the whole arena
  float* exp_avg_sq = nullptr;
  int32_t* union_mask = nullptr;   // [P + 4]
  bool stats = false;              // densification statistics kept (in the tail of the parameter arena)
  int32_t* radii = nullptr;        // [P] of the current keyframe
  int32_t* mask = nullptr;         // [P] rows of the batch that received a gradient (gsr_backward_args.grad_rows)
  float* dmean2D = nullptr;        // [3P] per-view output
  float* dcolor = nullptr;         // [3P] per-view output
  float* dL = nullptr;             // [3HW] dL/d(image) from the fused loss
  float* color = nullptr;          // [3HW] when the caller does not want the image
  size_t img_floats = 0;
  DevBuf geom, binning, image, loss_ws;
  gsr_context* ctx = nullptr;
  int64_t adam_step = 0;

  float* param_ptr(int k) const { return reinterpret_cast<float*>(peer[k] + off_param); }
  float* grad_ptr(int k) const { return reinterpret_cast<float*>(peer[k] + off_grad); }
  float* stats_local_ptr(int k) const { return grad_ptr(k) + total; }                 // [accum | denom | max_radii]
  float* stat_ptr(int j) const { return param_ptr(cfg.rank) + total + j * stats_stride; }   // j: 0 accum, 1 denom, 2 max
  uint8_t* flags_ptr(int k) const { return reinterpret_cast<uint8_t*>(peer[k] + off_flags); }
  uint32_t* pad_ptr(int k) const { return reinterpret_cast<uint32_t*>(peer[k] + off_pad); }
};

static int mapper_barrier(gsr_mapper* m, cudaStream_t s) {
  PadPtrs pads;
  for (int k = 0; k < m->cfg.world; ++k) pads.p[k] = m->pad_ptr(k);
  m->epoch += 1;
  peer_barrier_kernel<<<1, 32, 0, s>>>(pads, m->cfg.rank, m->cfg.world, m->epoch);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

static int ensure_image(gsr_mapper* m, int W, int H) {
  const size_t need = (size_t)3 * W * H;
  if (need <= m->img_floats) return 0;
  if (m->dL) cudaFree(m->dL);
  if (m->color) cudaFree(m->color);
  m->dL = m->color = nullptr;
  MAP_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->dL), need * sizeof(float)));
  MAP_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->color), need * sizeof(float)));
  m->img_floats = need;
  return 0;
}

extern "C" {

gsr_mapper* gsr_mapper_create(const gsr_mapper_config* cfg, void* handle_out) {
  if (!cfg || cfg->P <= 0 || cfg->P >= (1 << 24) || cfg->world < 1 || cfg->world > GSR_MAPPER_MAX_RANKS ||
      cfg->rank < 0 || cfg->rank >= cfg->world || (cfg->M != 0 && cfg->M != 15) || (cfg->world > 1 && !handle_out))
    return nullptr;
  gsr_mapper* m = new (std::nothrow) gsr_mapper();
  if (!m) return nullptr;
  m->cfg = *cfg;
  m->P = cfg->P;
  m->widths[GSR_SEG_FREST] = 3 * cfg->M;
  int64_t off = 0;
  for (int k = 0; k < kSegs; ++k) {
    m->seg_off[k] = off;
    m->seg_n[k] = m->P * m->widths[k];
    off += (m->seg_n[k] + 3) / 4 * 4;
    m->lr[k] = cfg->lr[k];
  }
  m->total = off;
  m->stats_stride = (m->P + 3) / 4 * 4;
  m->stats = cfg->with_stats != 0;
  m->arena = m->total + (m->stats ? 3 * m->stats_stride : 0);
  bool ok = cudaGetDevice(&m->device) == cudaSuccess;
  // one allocation for everything the peers touch, so one IPC handle describes it
  size_t o = 0;
  auto carve_off = [&](size_t bytes) { const size_t r = o; o = (o + bytes + 255) / 256 * 256; return r; };
  m->off_param = carve_off((size_t)m->arena * 4);
  m->off_grad = carve_off((size_t)m->arena * 4);
  m->off_flags = carve_off((size_t)m->P + 16);
  m->off_pad = carve_off(256);
  m->sym_bytes = o;
  const int64_t n_chunks = (m->arena + kShardChunk - 1) / kShardChunk;
  const int64_t owned = n_chunks > cfg->rank ? (n_chunks - cfg->rank + cfg->world - 1) / cfg->world : 0;
  const size_t n_mom = cfg->world > 1 ? (size_t)(owned > 0 ? owned : 1) * kShardChunk : (size_t)m->total;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&m->sym), m->sym_bytes) == cudaSuccess &&
       cudaMemset(m->sym, 0, m->sym_bytes) == cudaSuccess &&
       cudaMalloc(reinterpret_cast<void**>(&m->exp_avg), n_mom * 4) == cudaSuccess &&
       cudaMalloc(reinterpret_cast<void**>(&m->exp_avg_sq), n_mom * 4) == cudaSuccess &&
       cudaMemset(m->exp_avg, 0, n_mom * 4) == cudaSuccess && cudaMemset(m->exp_avg_sq, 0, n_mom * 4) == cudaSuccess &&
       cudaMalloc(reinterpret_cast<void**>(&m->union_mask), ((size_t)m->P + 4) * 4) == cudaSuccess &&
       cudaMalloc(reinterpret_cast<void**>(&m->radii), (size_t)m->P * 4) == cudaSuccess &&
       cudaMalloc(reinterpret_cast<void**>(&m->mask), (size_t)m->P * 4) == cudaSuccess &&
       cudaMalloc(reinterpret_cast<void**>(&m->dmean2D), (size_t)m->P * 12) == cudaSuccess &&
       cudaMalloc(reinterpret_cast<void**>(&m->dcolor), (size_t)m->P * 12) == cudaSuccess;
  if (ok) {
    m->ctx = gsr_context_create();
    ok = m->ctx != nullptr;
  }
  if (!ok) {
    gsr_mapper_destroy(m);
    return nullptr;
  }
  m->peer[cfg->rank] = m->sym;
  if (cfg->world > 1) {
    HandleBlob blob;
    std::memset(&blob, 0, sizeof(blob));
    blob.magic = 0x67737231u;
    blob.rank = cfg->rank;
    blob.bytes = m->sym_bytes;
    if (cudaIpcGetMemHandle(&blob.mem, m->sym) != cudaSuccess) {
      gsr_mapper_destroy(m);
      return nullptr;
    }
    std::memset(handle_out, 0, GSR_MAPPER_HANDLE_BYTES);
    std::memcpy(handle_out, &blob, sizeof(blob));
  } else {
    m->connected = true;
  }
  cudaDeviceSynchronize();
  return m;
}

int gsr_mapper_connect(gsr_mapper* m, const void* all_handles) {
  if (!m) return GSR_ERR_BAD_ARG;
  if (m->cfg.world == 1) return GSR_OK;
  if (!all_handles || m->connected) return GSR_ERR_BAD_ARG;
  for (int k = 0; k < m->cfg.world; ++k) {
    if (k == m->cfg.rank) continue;
    HandleBlob blob;
    std::memcpy(&blob, static_cast<const char*>(all_handles) + (size_t)k * GSR_MAPPER_HANDLE_BYTES, sizeof(blob));
    if (blob.magic != 0x67737231u || blob.rank != k || blob.bytes != m->sym_bytes) return GSR_ERR_BAD_ARG;
    void* p = nullptr;
    MAP_CUDA(cudaIpcOpenMemHandle(&p, blob.mem, cudaIpcMemLazyEnablePeerAccess));
    m->peer[k] = static_cast<char*>(p);
  }
  m->connected = true;
  return GSR_OK;
}

float* gsr_mapper_param(gsr_mapper* m, int segment) {
  if (!m || segment < 0 || segment >= kSegs || m->seg_n[segment] == 0) return nullptr;
  return m->param_ptr(m->cfg.rank) + m->seg_off[segment];
}

float* gsr_mapper_stat(gsr_mapper* m, int which) {
  if (!m || !m->stats || which < 0 || which > 2) return nullptr;
  static const int tail[3] = {2, 0, 1};   // API order: max_radii2D, xyz_gradient_accum, denom
  return m->stat_ptr(tail[which]);
}

int gsr_mapper_set_lr(gsr_mapper* m, int segment, double lr) {
  if (!m || segment < 0 || segment >= kSegs) return GSR_ERR_BAD_ARG;
  m->lr[segment] = lr;
  return GSR_OK;
}

int64_t gsr_mapper_steps_done(const gsr_mapper* m) { return m ? m->adam_step : 0; }

int gsr_mapper_step(gsr_mapper* m, const gsr_mapper_keyframe* kfs, int n, const float* background, float grad_scale,
                    float* loss_out, void* stream) {
  if (!m || n < 0 || (n > 0 && !kfs) || !background || !m->connected) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int world = m->cfg.world, rank = m->cfg.rank;
  const int P = (int)m->P;
  float* params = m->param_ptr(rank);
  float* grads = m->grad_ptr(rank);
  auto seg = [&](float* base, int k) { return m->seg_n[k] ? base + m->seg_off[k] : nullptr; };
  float* stat_accum = m->stats ? m->stat_ptr(0) : nullptr;
  float* stat_denom = m->stats ? m->stat_ptr(1) : nullptr;
  float* stat_maxr = m->stats ? m->stat_ptr(2) : nullptr;

  if (world > 1) {
    // this rank's "rendered" flags and statistics deltas of the step (the previous step's second barrier
    // guarantees that no peer still reads them)
    MAP_CUDA(cudaMemsetAsync(m->flags_ptr(rank), 0, (size_t)P + 16, s));
    if (m->stats) MAP_CUDA(cudaMemsetAsync(m->stats_local_ptr(rank), 0, (size_t)3 * m->stats_stride * 4, s));
  }
  // the batch's row mask: kept by gsr_backward itself (rows that received a gradient so far)
  MAP_CUDA(cudaMemsetAsync(m->mask, 0, (size_t)P * 4, s));
  for (int i = 0; i < n; ++i) {
    const gsr_mapper_keyframe& kf = kfs[i];
    if (!kf.viewmatrix || !kf.projmatrix || !kf.campos || kf.W <= 0 || kf.H <= 0 || (!kf.dL_dpix && !kf.gt_image))
      return GSR_ERR_BAD_ARG;
    MAP_RC(ensure_image(m, kf.W, kf.H));
    float* color = kf.out_color ? kf.out_color : m->color;
    int R = 0, B = 0;
    gsr_forward_args fa;
    std::memset(&fa, 0, sizeof(fa));
    fa.P = P; fa.D = m->cfg.sh_degree; fa.M = m->cfg.M; fa.W = kf.W; fa.H = kf.H;
    fa.background = background;
    fa.means3D = seg(params, GSR_SEG_XYZ); fa.dc = seg(params, GSR_SEG_FDC); fa.shs = seg(params, GSR_SEG_FREST);
    fa.opacities = seg(params, GSR_SEG_OPACITY); fa.scales = seg(params, GSR_SEG_SCALING);
    fa.rotations = seg(params, GSR_SEG_ROTATION);
    fa.scale_modifier = 1.0f;
    fa.viewmatrix = kf.viewmatrix; fa.projmatrix = kf.projmatrix; fa.campos = kf.campos;
    fa.tan_fovx = kf.tanfovx; fa.tan_fovy = kf.tanfovy;
    fa.out_color = color; fa.radii = m->radii;
    fa.geom_resize = DevBuf::resize; fa.geom_user = &m->geom;
    fa.binning_resize = DevBuf::resize; fa.binning_user = &m->binning;
    fa.image_resize = DevBuf::resize; fa.image_user = &m->image;
    fa.stream = s; fa.num_rendered = &R; fa.num_buckets = &B;
    fa.raw_params = m->cfg.raw_params;
    fa.context = m->ctx;
    MAP_RC(gsr_forward(&fa));

    const float* dL = kf.dL_dpix;
    if (!dL) {
      // photometric loss of the reference's iteration (src/gaussian_mapper.cpp:414-431), forward + backward fused
      char* ws = DevBuf::resize(&m->loss_ws, gsr_l1_ssim_workspace_bytes(kf.H, kf.W));
      if (!ws) return GSR_ERR_ALLOC;
      MAP_RC(gsr_l1_ssim_loss(color, kf.gt_image, kf.mask, kf.H, kf.W, m->cfg.lambda_dssim, loss_out, m->dL, ws, s));
      dL = m->dL;
    }
    gsr_backward_args ba;
    std::memset(&ba, 0, sizeof(ba));
    ba.P = P; ba.D = m->cfg.sh_degree; ba.M = m->cfg.M; ba.R = R; ba.B = B; ba.W = kf.W; ba.H = kf.H;
    ba.background = background;
    ba.means3D = fa.means3D; ba.dc = fa.dc; ba.shs = fa.shs; ba.opacities = fa.opacities;
    ba.scales = fa.scales; ba.scale_modifier = 1.0f; ba.rotations = fa.rotations;
    ba.viewmatrix = kf.viewmatrix; ba.projmatrix = kf.projmatrix; ba.campos = kf.campos;
    ba.tan_fovx = kf.tanfovx; ba.tan_fovy = kf.tanfovy;
    ba.radii = m->radii;
    ba.geom_buffer = m->geom.p; ba.binning_buffer = m->binning.p; ba.image_buffer = m->image.p;
    ba.dL_dpix = dL;
    ba.dL_dmean2D = m->dmean2D; ba.dL_dcolor = m->dcolor;
    ba.dL_dmean3D = seg(grads, GSR_SEG_XYZ); ba.dL_ddc = seg(grads, GSR_SEG_FDC); ba.dL_dsh = seg(grads, GSR_SEG_FREST);
    ba.dL_dopacity = seg(grads, GSR_SEG_OPACITY); ba.dL_dscale = seg(grads, GSR_SEG_SCALING);
    ba.dL_drot = seg(grads, GSR_SEG_ROTATION);
    ba.stream = s;
    ba.raw_params = m->cfg.raw_params;
    // no keyframe zero-fills: the first writes the rows that received a gradient, the others add to the rows
    // already written (grad_rows) and overwrite the rest; rows without gradient are never touched
    ba.rendered_rows_only = 1;
    ba.accumulate = i > 0;
    ba.grad_rows = m->mask;
    MAP_RC(gsr_backward(&ba));

    if (world > 1) {
      MAP_RC(gsr_peer_local_stats(P, m->radii, m->dmean2D, m->stats ? m->stats_local_ptr(rank) : nullptr,
                                  m->stats_stride, m->flags_ptr(rank), m->mask, s));
    } else if (m->stats) {
      MAP_RC(gsr_add_densification_stats(P, m->radii, m->dmean2D, stat_maxr, stat_accum, stat_denom, s));
    }
  }

  m->adam_step += 1;
  if (world == 1) {
    gsr_adam_group groups[kSegs];
    int ng = 0;
    for (int k = 0; k < kSegs; ++k) {
      if (m->seg_n[k] == 0) continue;
      groups[ng].param = params + m->seg_off[k];
      groups[ng].grad = grads + m->seg_off[k];
      groups[ng].exp_avg = m->exp_avg + m->seg_off[k];
      groups[ng].exp_avg_sq = m->exp_avg_sq + m->seg_off[k];
      groups[ng].numel = m->seg_n[k];
      groups[ng].lr = m->lr[k];
      groups[ng].step = m->adam_step;
      ++ng;
    }
    return gsr_adam_step_rows(groups, ng, m->cfg.beta1, m->cfg.beta2, m->cfg.eps, grad_scale, P, m->mask,
                              GSR_ADAM_ROWS_ALL, s);
  }

  MAP_RC(mapper_barrier(m, s));   // every rank's gradients, flags and statistics deltas are written
  const uint8_t* fl[GSR_MAPPER_MAX_RANKS];
  for (int k = 0; k < world; ++k) fl[k] = m->flags_ptr(k);
  MAP_RC(gsr_peer_rank_mask_u8(fl, world, P, m->union_mask, s));   // per row: who holds a gradient, who rendered it
  gsr_peer_adam_args a;
  std::memset(&a, 0, sizeof(a));
  a.n_peers = world;
  a.rank = rank;
  for (int k = 0; k < world; ++k) {
    a.param_peers[k] = m->param_ptr(k);
    a.grad_peers[k] = m->grad_ptr(k);
  }
  int ns = 0;
  for (int k = 0; k < kSegs; ++k) {
    if (m->seg_n[k] == 0) continue;
    a.segments[ns].start = m->seg_off[k];
    a.segments[ns].numel = m->seg_n[k];
    a.segments[ns].width = m->widths[k];
    a.segments[ns].lr = m->lr[k];
    ++ns;
  }
  if (m->stats) {   // the statistics ride along: accum, denom summed over the ranks, max_radii their maximum
    static const int rules[3] = {GSR_PEER_RULE_SUM, GSR_PEER_RULE_SUM, GSR_PEER_RULE_MAX};
    for (int j = 0; j < 3; ++j, ++ns) {
      a.segments[ns].start = m->total + j * m->stats_stride;
      a.segments[ns].numel = P;
      a.segments[ns].width = 1;
      a.segments[ns].rule = rules[j];
    }
  }
  a.n_segments = ns;
  a.shard_lo = 0;
  a.shard_hi = m->arena;
  a.shard_chunk = kShardChunk;
  a.exp_avg = m->exp_avg;
  a.exp_avg_sq = m->exp_avg_sq;
  a.union_mask = m->union_mask;
  a.beta1 = m->cfg.beta1; a.beta2 = m->cfg.beta2; a.eps = m->cfg.eps;
  a.grad_scale = grad_scale;
  a.step = m->adam_step;
  a.stream = s;
  a.param_multicast = nullptr;   // CUDA IPC mappings have no multicast address (torch symmetric memory does)
  MAP_RC(gsr_adam_step_peer(&a));
  return mapper_barrier(m, s);    // every rank's parameter stores have landed
}

void gsr_mapper_destroy(gsr_mapper* m) {
  if (!m) return;
  cudaDeviceSynchronize();
  for (int k = 0; k < m->cfg.world; ++k)
    if (k != m->cfg.rank && m->peer[k]) cudaIpcCloseMemHandle(m->peer[k]);
  void* bufs[] = {m->sym, m->exp_avg, m->exp_avg_sq, m->union_mask, m->radii, m->mask,
                  m->dmean2D, m->dcolor, m->dL, m->color};
  for (void* b : bufs)
    if (b) cudaFree(b);
  m->geom.free(); m->binning.free(); m->image.free(); m->loss_ws.free();
  if (m->ctx) gsr_context_destroy(m->ctx);
  delete m;
}

}  // extern "C"
