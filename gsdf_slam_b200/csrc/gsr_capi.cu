// C ABI of the B200-native rasterizer hot path (see include/gsr_capi.h).
// Host-side orchestration only: every entry point validates its arguments,
// carves the caller-provided scratch buffers and enqueues the kernels of
// preprocess.cu / radix_sort.cu / binning.cu / blend_fwd.cu / blend_bwd.cu /
// preprocess_bwd.cu / adam.cu / knn.cu on the caller's stream.
//
// Forward pipeline (reference: Rasterizer::forward, rasterizer_impl.cu:366-538):
//   memset(counters) -> precull (conservative cull -> candidates; skipped while most of the map is rendered) ->
//   preprocess (geometry, SH colours, EXACT per-tile splat counts) [-> count_large] -> tile_meta (scan of the T
//   counts: tile ranges, R, launch order) -> emit (every duplicate dropped into its tile's range) ->
//   tile_sort_pack (per-tile depth sort + packed records; of lists of 512+ entries only the nearest < 512 are
//   selected, sorted and packed) -> blend_fwd [-> full sort of the tiles whose pixels were not saturated at the end
//   of that prefix + a second blend_fwd launch that continues their half tiles behind it; a context in which more
//   than 1/16 of the long lists need that orders its long lists in full].
//   No global sort, no scan over Gaussians, no mid-pipeline host synchronisation: the per-duplicate
//   buffer is sized optimistically from the previous call of the same context and R (24 bytes, read
//   back asynchronously) is only checked once everything has been enqueued (the reference has two
//   blocking copies plus, as shipped with debug=true, a device-wide sync after every kernel).
// Host-side state (read-back event, pinned words, sizing hints) lives in a gsr_context: caller-owned
// (gsr_context_create) or, by default, one per (host thread, device) — never shared between threads
// or devices.
#include "gsr_capi.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "gsr_common.cuh"

namespace gsr {
std::atomic<long long> g_launch_count{0};

}  // namespace gsr

#include "gsr_params.cuh"

namespace gsr {

}  // namespace gsr

// Host-side state of gsr_forward: everything that used to be implicit per-thread state.  Bound to the
// device that was current when it was created.
// the prefix sort is switched off when more than 1 / GSR_PARTIAL_REDO_DIV of the long lists needed the second pass
#ifndef GSR_PARTIAL_REDO_DIV
#define GSR_PARTIAL_REDO_DIV 16
#endif
struct gsr_context {
  int device = -1;
  uint32_t* pinned = nullptr;     // 64 bytes of pinned host memory: R and the list-length class counts
  cudaEvent_t ev = nullptr;       // read-back complete
  // the per-duplicate buffer is sized from the previous call with the same (P, W, H); 0 = unknown
  uint32_t cap_hint = 0;
  int hint_P = -1, hint_W = -1, hint_H = -1;
  // per-tile sort kernels for mid-length (512+) / long (4092+) lists are launched only if recent frames
  // had such lists (-1: unknown, n > 0: seen within the last n calls)
  int hint_mid = -1, hint_long = -1;
  // the previous call rendered more than half of the map: the conservative cull has nothing to remove, skip it
  int hint_dense = 0;
  // Prefix sort of the lists of 512+ entries (binning.cu: tile_select_kernel): pays off where pixels saturate
  // within the prefix (dense views); where many tiles end up in the full sort anyway it is switched off for
  // partial_off calls.  The count of redone tiles of call k is read back asynchronously and looked at in call k + 1.
  int partial_off = 0;
  bool redo_pending = false;
  uint32_t redo_long_tiles = 0;
};

namespace gsr {

static gsr_context* context_new() {
  gsr_context* c = new gsr_context();
  if (cudaGetDevice(&c->device) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void**>(&c->pinned), 64, cudaHostAllocPortable) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev, cudaEventDisableTiming) != cudaSuccess) {
    if (c->pinned) cudaFreeHost(c->pinned);
    delete c;
    return nullptr;
  }
  return c;
}

// default context of the calling thread for the current device
static gsr_context* default_context() {
  static thread_local std::vector<gsr_context*> per_device;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return nullptr;
  if ((size_t)dev >= per_device.size()) per_device.resize((size_t)dev + 1, nullptr);
  if (!per_device[dev]) per_device[dev] = context_new();
  return per_device[dev];
}

// ---- optional per-stage CUDA-event profiler (bench.py's live roofline measurement) ----
struct ProfPending { int slot; cudaEvent_t a, b; };
struct Profiler {
  bool on = false;
  std::vector<std::string> names;
  std::vector<double> ms;
  std::vector<long long> count;
  std::vector<ProfPending> pending;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
  int slot_of(const char* n) {
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == n) return (int)i;
    names.push_back(n); ms.push_back(0.0); count.push_back(0);
    return (int)names.size() - 1;
  }
  void drain() {
    for (auto& p : pending) {
      float t = 0.f;
      if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) {
        ms[p.slot] += t; count[p.slot] += 1;
      }
      pool.push_back({p.a, p.b});
    }
    pending.clear();
  }
};
static Profiler g_prof;
static std::mutex g_prof_mutex;     // the profiler is a developer aid; when on, its bookkeeping is serialised
struct ProfScope {
  cudaStream_t s; int slot = -1; cudaEvent_t a{}, b{};
  ProfScope(const char* name, cudaStream_t st) : s(st) {
    if (!g_prof.on) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    slot = g_prof.slot_of(name);
    if (!g_prof.pool.empty()) { a = g_prof.pool.back().first; b = g_prof.pool.back().second; g_prof.pool.pop_back(); }
    else { cudaEventCreate(&a); cudaEventCreate(&b); }
    cudaEventRecord(a, s);
  }
  ~ProfScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    cudaEventRecord(b, s);
    g_prof.pending.push_back({slot, a, b});
  }
};
}  // namespace gsr

using namespace gsr;
#define GSR_PROF(name, call)        \
  do {                              \
    ProfScope _ps(name, s);         \
    int _r = (call);                \
    if (_r != 0) return _r;         \
  } while (0)

#define GSR_CUDA(x)                         \
  do {                                      \
    cudaError_t _e = (x);                   \
    if (_e != cudaSuccess) return (int)_e;  \
  } while (0)
#define GSR_RC(x)            \
  do {                       \
    int _r = (x);            \
    if (_r != 0) return _r;  \
  } while (0)

extern "C" {

const char* gsr_version(void) { return "gsdf_slam_b200 0.1 (sm_100a)"; }
int64_t gsr_launch_count(void) { return (int64_t)gsr::g_launch_count.load(std::memory_order_relaxed); }

gsr_context* gsr_context_create(void) { return gsr::context_new(); }
void gsr_context_destroy(gsr_context* c) {
  if (!c) return;
  if (c->ev) cudaEventDestroy(c->ev);
  if (c->pinned) cudaFreeHost(c->pinned);
  delete c;
}

int gsr_forward(const gsr_forward_args* a) {
  if (!a || a->P < 0 || a->W <= 0 || a->H <= 0 || !a->out_color || !a->num_rendered) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(a->stream);
  *a->num_rendered = 0;
  if (a->num_buckets) *a->num_buckets = 0;
  const int P = a->P, W = a->W, H = a->H;
  const size_t N = (size_t)W * H;
  if (P == 0) {
    // reference: zero image, empty buffers, nothing launched (src/rasterize_points.cu:74-90)
    GSR_CUDA(cudaMemsetAsync(a->out_color, 0, 3 * N * sizeof(float), s));
    return GSR_OK;
  }
  if (!a->means3D || !a->opacities || !a->radii || !a->viewmatrix || !a->projmatrix || !a->campos ||
      !a->background || !a->geom_resize || !a->binning_resize || !a->image_resize)
    return GSR_ERR_BAD_ARG;
  if (!a->colors_precomp && !a->dc) return GSR_ERR_BAD_ARG;
  if (!a->cov3D_precomp && (!a->scales || !a->rotations)) return GSR_ERR_BAD_ARG;
  if (P >= (1 << 24)) return GSR_ERR_BAD_ARG;  // work-queue entries of the emit stage hold 24-bit list positions
  gsr_context* ctx = a->context ? static_cast<gsr_context*>(a->context) : default_context();
  if (!ctx) return GSR_ERR_ALLOC;
  {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != ctx->device) return GSR_ERR_BAD_ARG;  // context of another device
  }

  const int gx = (W + GSR_TILE_X - 1) / GSR_TILE_X, gy = (H + GSR_TILE_Y - 1) / GSR_TILE_Y;
  const uint32_t T = (uint32_t)gx * gy;

  char* geom_base = a->geom_resize(a->geom_user, geom_bytes((size_t)P));
  char* img_base = a->image_resize(a->image_user, image_bytes(N, T));
  if (!geom_base || !img_base) return GSR_ERR_ALLOC;
  GeomView g = geom_view(geom_base, (size_t)P);
  ImageView im = image_view(img_base, N, T);

  GSR_CUDA(cudaMemsetAsync(g.hdr, 0, sizeof(GeomHeader), s));
  GSR_CUDA(cudaMemsetAsync(im.tile_count, 0, im.zero_bytes, s));  // tile_count | tile_fill | redo_flag | redo_count

  uint32_t* host_R = ctx->pinned;
  cudaEvent_t ev_R = ctx->ev;

  // The per-duplicate buffer is sized OPTIMISTICALLY from the previous call of this context, so
  // the whole forward pass is enqueued without waiting for R (no kernel's grid depends on
  // it; every write is guarded by the capacity).  R is checked once everything is enqueued; if
  // it did not fit — or on the first call — the per-duplicate stages are (re)run with the exact size.
  // The same optimism decides which of the per-tile sort kernels are launched: lists of 512+ (mid) and
  // 4092+ (long) entries get their kernels only if recent frames had such lists.
  // Both hints are sticky (a class seen once keeps its kernel for kHintFrames calls, the capacity decays
  // by 1/64 per call) so that a mapper alternating between keyframes does not pay a re-run on every call.
  constexpr int kHintFrames = 64;
  uint32_t& cap_hint = ctx->cap_hint;
  int& hint_mid = ctx->hint_mid;
  int& hint_long = ctx->hint_long;
  if (ctx->hint_P != P || ctx->hint_W != W || ctx->hint_H != H) {  // a different scene / resolution
    cap_hint = 0;
    hint_mid = hint_long = -1;
    ctx->hint_dense = 0;
    ctx->partial_off = 0;
    ctx->redo_pending = false;
  }
  ctx->hint_P = P; ctx->hint_W = W; ctx->hint_H = H;
  uint32_t cap = cap_hint;

  PreprocessParams pp;
  std::memset(&pp, 0, sizeof(pp));
  pp.P = P; pp.D = a->D; pp.M = a->M; pp.W = W; pp.H = H;
  pp.means3D = a->means3D; pp.scales = a->cov3D_precomp ? nullptr : a->scales;
  pp.rotations = a->cov3D_precomp ? nullptr : a->rotations;
  pp.opacities = a->opacities; pp.dc = a->dc; pp.shs = (a->M > 0) ? a->shs : nullptr;
  pp.cov3D_precomp = a->cov3D_precomp; pp.colors_precomp = a->colors_precomp;
  pp.scale_modifier = a->scale_modifier;
  pp.view = a->viewmatrix; pp.proj = a->projmatrix; pp.campos = a->campos;
  pp.tan_fovx = a->tan_fovx; pp.tan_fovy = a->tan_fovy;
  pp.focal_y = H / (2.0f * a->tan_fovy);
  pp.focal_x = W / (2.0f * a->tan_fovx);
  pp.grid_x = gx; pp.grid_y = gy;
  pp.prefiltered = a->prefiltered;
  pp.activate = a->raw_params != 0;
  const uintptr_t al = (uintptr_t)a->means3D | (uintptr_t)pp.scales | (uintptr_t)pp.rotations |
                       (uintptr_t)a->dc;
  pp.vec_ok = (al & 15) == 0;
  pp.radii = a->radii; pp.depths = g.depths; pp.means2D = g.means2D;
  pp.conic_opacity = g.conic_opacity; pp.rgb = g.rgb; pp.clamped = g.clamped;
  pp.tiles_touched = g.tiles_touched; pp.splat_rec = g.splat_rec; pp.vis_items = g.vis_items; pp.num_visible = &g.hdr->num_visible;
  // the candidate list of the cull lives in the backward pass's moment sums (unused until gsr_backward zeroes them)
  pp.cand = reinterpret_cast<uint32_t*>(g.acc); pp.num_cand = &g.hdr->num_cand;
  pp.large_list = g.large_list; pp.large_count = &g.hdr->num_large; pp.large_capacity = 2u * (uint32_t)P;
  pp.tile_count = im.tile_count;
  if (ctx->hint_dense > 0) pp.cand = nullptr;   // every Gaussian is a candidate (stage 2 repeats the z test anyway)
  else GSR_PROF("precull", launch_precull(pp, s));
  GSR_PROF("preprocess", launch_preprocess(pp, s));
  if (a->radii_ready_event) GSR_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(a->radii_ready_event), s));

  // ranges, R and the per-tile launch order from the exact per-tile counts
  GSR_PROF("tile_meta", launch_tile_meta(im.tile_count, T, im.ranges, g.hdr, im.tile_order, im.max_contrib, s));
  // R and the numbers of mid-length / long per-tile lists (header words 0 .. 5)
  GSR_CUDA(cudaMemcpyAsync(host_R, &g.hdr->num_dup, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  GSR_CUDA(cudaEventRecord(ev_R, s));

  uint32_t R = 0;
  bool have_R = false;
  bool do_mid = hint_mid != 0, do_long = hint_long != 0;
  if (cap == 0) {
    GSR_CUDA(cudaEventSynchronize(ev_R));
    R = host_R[0];
    have_R = true;
    cap = R;
    do_mid = host_R[4] > host_R[5];
    do_long = host_R[5] > 0;
  }
  BinView b;
  bool redo_issued = false;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const bool longer = do_mid || do_long;          // lists of 512+ entries exist (or are expected)
    const bool partial = ctx->partial_off <= 0;     // ... and only their depth-sorted PREFIX is packed
    char* bin_base = a->binning_resize(a->binning_user, bin_bytes((size_t)cap));
    if (!bin_base) return GSR_ERR_ALLOC;
    b = bin_view(bin_base, (size_t)cap);
    if (attempt > 0) {
      // max_contrib may hold values of the aborted attempt (blend_fwd only ever raises it)
      // tile_fill | redo_flag | redo_count (contiguous, behind tile_count)
      GSR_CUDA(cudaMemsetAsync(im.tile_fill, 0, im.zero_bytes - (size_t)(reinterpret_cast<char*>(im.tile_fill) -
                                                                      reinterpret_cast<char*>(im.tile_count)), s));
      GSR_CUDA(cudaMemsetAsync(im.max_contrib, 0, (size_t)T * sizeof(uint32_t), s));
      GSR_CUDA(cudaMemsetAsync(&g.hdr->num_large, 0, sizeof(uint32_t), s));
    }
    if (!have_R || R > 0) {
      GSR_PROF("emit", launch_emit(P, g.vis_items, &g.hdr->num_visible, g.tiles_touched, g.means2D,
                                   g.conic_opacity, a->radii, gx, gy, im.ranges, im.tile_fill, b.items, cap,
                                   &g.hdr->num_large, g.large_list, s));
      GSR_PROF("tile_sort_pack", launch_tile_sort_pack(P, T, im.ranges, im.tile_order, g.hdr, b.hdr, cap, b.items, b.scratch,
                                                       g.splat_rec, b.rec, do_mid && partial, do_long && partial,
                                                       im.sorted_len, im.redo_flag, im.redo_list, im.redo_count, s));
      if (longer && !partial)   // the prefix sort is switched off for this context: every long list in full
        GSR_PROF("tile_sort_full", launch_tile_sort_full(P, T, im.ranges, im.tile_order, g.hdr, cap, b.items, b.scratch,
                                                         g.splat_rec, b.rec, do_mid, do_long, nullptr, nullptr, s));
    }
    GSR_PROF("blend_fwd", launch_blend_fwd(W, H, im.ranges, im.tile_order, b.rec, a->background, a->out_color,
                            im.final_T, im.n_contrib, im.accum_c, im.max_contrib, cap, nullptr,
                            (longer && partial) ? im.sorted_len : nullptr, im.redo_flag, im.redo_list, im.redo_count,
                            false, s));
    if (longer && partial && (!have_R || R > 0)) {
      // the tiles whose pixels were not saturated at the end of their prefix: full sort, and their half tiles
      // continue behind the prefix
      GSR_PROF("tile_sort_redo", launch_tile_sort_full(P, T, im.ranges, im.tile_order, g.hdr, cap, b.items, b.scratch,
                                                       g.splat_rec, b.rec, do_mid, do_long, im.redo_list, im.redo_count, s));
      GSR_PROF("blend_fwd_redo", launch_blend_fwd(W, H, im.ranges, im.tile_order, b.rec, a->background, a->out_color,
                              im.final_T, im.n_contrib, im.accum_c, im.max_contrib, cap, nullptr, im.sorted_len,
                              im.redo_flag, im.redo_list, im.redo_count, true, s));
      GSR_CUDA(cudaMemcpyAsync(&host_R[8], im.redo_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
      redo_issued = true;
    }
    if (!have_R) {
      GSR_CUDA(cudaEventSynchronize(ev_R));  // completed long ago: the GPU is several kernels further
      R = host_R[0];
      have_R = true;
    }
    const bool need_mid = host_R[4] > host_R[5], need_long = host_R[5] > 0;
    const bool sort_ok = (do_mid || !need_mid) && (do_long || !need_long);
    if (R <= cap && sort_ok) break;
    // the optimistic buffer was too small, or a list length class appeared whose sort kernel was not
    // launched: redo the per-duplicate stages with the exact answers
    if (R > cap) cap = R;
    do_mid = need_mid;
    do_long = need_long;
  }
  {
    const uint32_t want = R + R / 8 + 4096, decayed = cap_hint - cap_hint / 64;
    cap_hint = want > decayed ? want : decayed;
  }
  ctx->hint_dense = host_R[2] > (uint32_t)P / 2 ? 1 : 0;   // header word 2: rendered Gaussians
  // the previous call's count of redone tiles has arrived (its copy precedes this call's tile_meta in stream order)
  // (the second pass is a serial tail — a flagged tile's CTA walks the rest of a long list alone, after everybody
  // else — so it pays only while few tiles need it; views whose pixels saturate, cfg1 / cfg4, flag none)
  if (ctx->redo_pending && (uint64_t)host_R[8] * GSR_PARTIAL_REDO_DIV > ctx->redo_long_tiles) ctx->partial_off = kHintFrames;
  else if (ctx->partial_off > 0) ctx->partial_off -= 1;
  ctx->redo_pending = redo_issued;
  ctx->redo_long_tiles = host_R[4];
  hint_mid = host_R[4] > host_R[5] ? kHintFrames : (hint_mid > 0 ? hint_mid - 1 : 0);
  hint_long = host_R[5] > 0 ? kHintFrames : (hint_long > 0 ? hint_long - 1 : 0);
  *a->num_rendered = (int)R;
  if (a->sample_resize) a->sample_resize(a->sample_user, 0);
  if (a->debug) {
    GSR_CUDA(cudaStreamSynchronize(s));
    GSR_CUDA(cudaGetLastError());
  }
  return GSR_OK;
}

int gsr_backward(const gsr_backward_args* a) {
  if (!a || a->P < 0 || a->W <= 0 || a->H <= 0) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(a->stream);
  const int P = a->P, W = a->W, H = a->H;
  if (P == 0) return GSR_OK;
  if (!a->geom_buffer || !a->image_buffer || !a->dL_dpix || !a->radii || !a->means3D ||
      !a->dL_dmean2D || !a->dL_dopacity || !a->dL_dcolor || !a->dL_dmean3D || !a->dL_ddc ||
      !a->dL_dscale || !a->dL_drot)
    return GSR_ERR_BAD_ARG;
  if (a->R > 0 && !a->binning_buffer) return GSR_ERR_BAD_ARG;
  if (a->M > 0 && a->shs && !a->dL_dsh) return GSR_ERR_BAD_ARG;
  if (((uintptr_t)a->dL_drot & 15) || ((uintptr_t)a->rotations & 15) || ((uintptr_t)a->dL_dconic & 15))
    return GSR_ERR_BAD_ARG;
  const size_t N = (size_t)W * H;
  const int gx = (W + GSR_TILE_X - 1) / GSR_TILE_X, gy = (H + GSR_TILE_Y - 1) / GSR_TILE_Y;
  const uint32_t T = (uint32_t)gx * gy;
  GeomView g = geom_view(a->geom_buffer, (size_t)P);
  ImageView im = image_view(a->image_buffer, N, T);

  GSR_CUDA(cudaMemsetAsync(g.acc, 0, (size_t)P * 12 * sizeof(float), s));
  if (a->R > 0) {
    BinView b = bin_view(a->binning_buffer, (size_t)a->R);  // rec does not depend on the capacity
    GSR_PROF("blend_bwd", launch_blend_bwd(W, H, im.ranges, im.tile_order, im.max_contrib, b.rec,
                            a->background, im.final_T, im.n_contrib, im.accum_c, a->dL_dpix, g.acc, s));
  }
  PreBwdParams bp;
  std::memset(&bp, 0, sizeof(bp));
  bp.P = P; bp.D = a->D; bp.M = a->M;
  bp.activate = a->raw_params != 0;
  bp.opacities = a->opacities;
  if (bp.activate && a->cov3D_precomp) return GSR_ERR_BAD_ARG;  // raw scale / rotation needed
  bp.means3D = a->means3D; bp.radii = a->radii; bp.dc = a->dc;
  bp.shs = (a->colors_precomp || a->M <= 0) ? nullptr : a->shs;
  bp.clamped = g.clamped;
  bp.scales = a->cov3D_precomp ? nullptr : a->scales;
  bp.rotations = a->cov3D_precomp ? nullptr : a->rotations;
  bp.cov3D_precomp = a->cov3D_precomp;
  bp.scale_modifier = a->scale_modifier;
  bp.view = a->viewmatrix; bp.proj = a->projmatrix; bp.campos = a->campos;
  bp.focal_y = H / (2.0f * a->tan_fovy);
  bp.focal_x = W / (2.0f * a->tan_fovx);
  bp.tan_fovx = a->tan_fovx; bp.tan_fovy = a->tan_fovy;
  bp.acc = g.acc; bp.conic_opacity = g.conic_opacity; bp.W = W; bp.H = H;
  bp.dL_dmean2D = a->dL_dmean2D; bp.dL_dconic = a->dL_dconic; bp.dL_dopacity = a->dL_dopacity;
  bp.dL_dcolor = a->dL_dcolor; bp.dL_dmean3D = a->dL_dmean3D; bp.dL_dcov3D = a->dL_dcov3D;
  bp.dL_ddc = a->dL_ddc; bp.dL_dsh = a->dL_dsh; bp.dL_dscale = a->dL_dscale; bp.dL_drot = a->dL_drot;
  bp.accumulate = a->accumulate != 0;
  bp.accumulate_rows = a->accumulate ? a->accumulate_rows : nullptr;
  bp.grad_rows = a->grad_rows;
  bp.dense_zeroed = a->rendered_rows_only == 0 && a->accumulate == 0;
  // the list of the Gaussians that received a gradient reuses the emit stage's work-queue storage and counter
  GSR_CUDA(cudaMemsetAsync(&g.hdr->num_large, 0, sizeof(uint32_t), s));
  GSR_PROF("preprocess_bwd", launch_preprocess_bwd(bp, g.vis_items, &g.hdr->num_visible, g.large_list, &g.hdr->num_large,
                                                   a->rendered_rows_only == 0 && a->accumulate == 0, s));
  if (a->debug) {
    GSR_CUDA(cudaStreamSynchronize(s));
    GSR_CUDA(cudaGetLastError());
  }
  return GSR_OK;
}

int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream) {
  (void)projmatrix;
  if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return GSR_ERR_BAD_ARG;
  return launch_mark_visible(P, means3D, viewmatrix, present, static_cast<cudaStream_t>(stream));
}

int gsr_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* stream) {
  if (P < 0 || (P > 0 && (!points || !mean_dist2))) return GSR_ERR_BAD_ARG;
  if (P == 0) return GSR_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* tmp = nullptr;
  GSR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tmp), knn_temp_bytes(P), s));
  int rc = launch_knn(P, points, mean_dist2, tmp, s);
  cudaError_t e = cudaFreeAsync(tmp, s);
  if (rc) return rc;
  return (int)e;
}

int gsr_adam_step(const gsr_adam_group* groups, int n_groups, double beta1, double beta2, double eps,
                  float grad_scale, void* stream) {
  if (n_groups < 0 || (n_groups > 0 && !groups)) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GSR_PROF("adam", launch_adam(groups, n_groups, beta1, beta2, eps, grad_scale, s));
  return GSR_OK;
}

int gsr_adam_step_rows(const gsr_adam_group* groups, int n_groups, double beta1, double beta2, double eps,
                       float grad_scale, int64_t rows, const int32_t* row_mask, int phase, void* stream) {
  if (n_groups < 0 || (n_groups > 0 && !groups) || rows < 0) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GSR_PROF(phase == GSR_ADAM_ROWS_ZERO_GRAD ? "adam_rows_zero_grad"
           : phase == GSR_ADAM_ROWS_WITH_GRAD ? "adam_rows_with_grad" : "adam_rows_all",
           launch_adam_rows(groups, n_groups, beta1, beta2, eps, grad_scale, rows, row_mask, phase, s));
  return GSR_OK;
}

int gsr_export_binning(const char* geom_buffer, const char* binning_buffer, const char* image_buffer,
                       int P, int R, int W, int H, uint64_t* keys_sorted, uint32_t* point_list, void* stream) {
  if (R <= 0) return GSR_OK;
  if (!binning_buffer || !image_buffer || !geom_buffer || P <= 0) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t N = (size_t)W * H;
  const uint32_t T = (uint32_t)(((W + GSR_TILE_X - 1) / GSR_TILE_X) * ((H + GSR_TILE_Y - 1) / GSR_TILE_Y));
  ImageView im = image_view(const_cast<char*>(image_buffer), N, T);
  // the sort arrays sit behind the records, whose extent depends on the carved capacity
  uint64_t cap = 0;
  GSR_CUDA(cudaMemcpyAsync(&cap, binning_buffer, sizeof(cap), cudaMemcpyDeviceToHost, s));
  GSR_CUDA(cudaStreamSynchronize(s));
  if (cap < (uint64_t)R) return GSR_ERR_BAD_ARG;
  BinView b = bin_view(const_cast<char*>(binning_buffer), (size_t)cap);
  // The forward pass orders only a prefix of the lists of 512+ entries (the part the blend kernels read): the
  // export completes the job — the full sort of every such list, from the list as emitted (re-sorting the few
  // lists the forward pass had to sort fully is harmless; the packed records keep the same prefix).
  GeomView g = geom_view(const_cast<char*>(geom_buffer), (size_t)P);
  GSR_RC(launch_tile_sort_full(P, T, im.ranges, im.tile_order, g.hdr, (uint32_t)cap, b.items, b.scratch, g.splat_rec,
                               b.rec, true, true, nullptr, nullptr, s));
  return launch_export_binning(T, im.ranges, b.items, keys_sorted, point_list, s);
}

int gsr_export_redo_count(const char* image_buffer, int W, int H, uint32_t* count_host, void* stream) {
  if (!image_buffer || !count_host || W <= 0 || H <= 0) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t N = (size_t)W * H;
  const size_t T = (size_t)((W + GSR_TILE_X - 1) / GSR_TILE_X) * ((H + GSR_TILE_Y - 1) / GSR_TILE_Y);
  ImageView im = image_view(const_cast<char*>(image_buffer), N, T);
  GSR_CUDA(cudaMemcpyAsync(count_host, im.redo_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  GSR_CUDA(cudaStreamSynchronize(s));
  return GSR_OK;
}

int gsr_export_image(const char* image_buffer, int W, int H, float* final_T, uint32_t* n_contrib,
                     uint32_t* ranges, uint32_t* max_contrib, void* stream) {
  if (!image_buffer) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t N = (size_t)W * H;
  const size_t T = (size_t)((W + GSR_TILE_X - 1) / GSR_TILE_X) * ((H + GSR_TILE_Y - 1) / GSR_TILE_Y);
  ImageView im = image_view(const_cast<char*>(image_buffer), N, T);
  if (final_T) GSR_CUDA(cudaMemcpyAsync(final_T, im.final_T, N * 4, cudaMemcpyDeviceToDevice, s));
  if (n_contrib) GSR_CUDA(cudaMemcpyAsync(n_contrib, im.n_contrib, N * 4, cudaMemcpyDeviceToDevice, s));
  if (ranges) GSR_CUDA(cudaMemcpyAsync(ranges, im.ranges, T * 8, cudaMemcpyDeviceToDevice, s));
  if (max_contrib) GSR_CUDA(cudaMemcpyAsync(max_contrib, im.max_contrib, T * 4, cudaMemcpyDeviceToDevice, s));
  return GSR_OK;
}

int gsr_export_pair_counts(const char* binning_buffer, char* image_buffer, int R, int W, int H,
                           const float* background, float* scratch_color, uint64_t* counts_out, void* stream) {
  // replays blend_fwd on the buffers of a finished gsr_forward with the instrumented kernel; every output it
  // rewrites (final_T, n_contrib, accum_c, max_contrib) receives the value it already holds
  if (!image_buffer || !background || !scratch_color || !counts_out || W <= 0 || H <= 0) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GSR_CUDA(cudaMemsetAsync(counts_out, 0, 3 * sizeof(uint64_t), s));
  if (R <= 0) return GSR_OK;
  if (!binning_buffer) return GSR_ERR_BAD_ARG;
  const size_t N = (size_t)W * H;
  const uint32_t T = (uint32_t)(((W + GSR_TILE_X - 1) / GSR_TILE_X) * ((H + GSR_TILE_Y - 1) / GSR_TILE_Y));
  ImageView im = image_view(image_buffer, N, T);
  BinView b = bin_view(const_cast<char*>(binning_buffer), (size_t)R);  // rec does not depend on the capacity
  return launch_blend_fwd(W, H, im.ranges, im.tile_order, b.rec, background, scratch_color, im.final_T, im.n_contrib,
                          im.accum_c, im.max_contrib, 0xffffffffu, reinterpret_cast<unsigned long long*>(counts_out),
                          im.sorted_len, im.redo_flag, im.redo_list, im.redo_count, false, s);
}

int gsr_export_geometry(const char* geom_buffer, int P, float* depths, float* means2D,
                        float* conic_opacity, float* rgb, uint32_t* tiles_touched, uint8_t* clamped,
                        void* stream) {
  if (P <= 0) return GSR_OK;
  if (!geom_buffer) return GSR_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GeomView g = geom_view(const_cast<char*>(geom_buffer), (size_t)P);
  const size_t n = (size_t)P;
  if (depths) GSR_CUDA(cudaMemcpyAsync(depths, g.depths, n * 4, cudaMemcpyDeviceToDevice, s));
  if (means2D) GSR_CUDA(cudaMemcpyAsync(means2D, g.means2D, n * 8, cudaMemcpyDeviceToDevice, s));
  if (conic_opacity) GSR_CUDA(cudaMemcpyAsync(conic_opacity, g.conic_opacity, n * 16, cudaMemcpyDeviceToDevice, s));
  if (rgb) GSR_CUDA(cudaMemcpyAsync(rgb, g.rgb, n * 12, cudaMemcpyDeviceToDevice, s));
  if (tiles_touched) GSR_CUDA(cudaMemcpyAsync(tiles_touched, g.tiles_touched, n * 4, cudaMemcpyDeviceToDevice, s));
  if (clamped) GSR_CUDA(cudaMemcpyAsync(clamped, g.clamped, n, cudaMemcpyDeviceToDevice, s));
  return GSR_OK;
}

void gsr_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  g_prof.drain();
  g_prof.on = on != 0;
  if (on) {
    for (auto& m : g_prof.ms) m = 0.0;
    for (auto& c : g_prof.count) c = 0;
  }
}

// Writes "name count total_ms\n" lines (NUL terminated) and returns the number of stages.
int gsr_profile_read(char* out, size_t cap) {
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  g_prof.drain();
  size_t off = 0;
  int n = 0;
  for (size_t i = 0; i < g_prof.names.size(); ++i) {
    if (g_prof.count[i] == 0) continue;
    int w = snprintf(out + off, off < cap ? cap - off : 0, "%s %lld %.6f\n", g_prof.names[i].c_str(),
                     g_prof.count[i], g_prof.ms[i]);
    if (w < 0 || off + (size_t)w >= cap) break;
    off += (size_t)w;
    ++n;
  }
  if (cap > 0) out[off < cap ? off : cap - 1] = 0;
  return n;
}

size_t gsr_sort_temp_bytes(int64_t n) { return gsr::sort_temp_bytes(n); }

int gsr_sort_u64(const uint64_t* items_in, uint64_t* items_out, uint64_t* items_alt, int64_t n,
                 int begin_bit, int end_bit, char* tmp, void* stream) {
  if (n < 0 || begin_bit < 0 || end_bit > 64 || end_bit < begin_bit) return GSR_ERR_BAD_ARG;
  if (n == 0) return GSR_OK;
  if (!items_in || !items_out || !tmp) return GSR_ERR_BAD_ARG;
  const int npass = (end_bit - begin_bit + 7) / 8;
  if (npass > 1 && (!items_alt || items_alt == items_out)) return GSR_ERR_BAD_ARG;
  if (items_in == ((npass & 1) ? items_out : items_alt)) return GSR_ERR_BAD_ARG;
  uint64_t* result = nullptr;
  return launch_sort_u64(items_in, items_out, items_alt, n, begin_bit, end_bit, tmp,
                         static_cast<cudaStream_t>(stream), &result);
}

}  // extern "C"
