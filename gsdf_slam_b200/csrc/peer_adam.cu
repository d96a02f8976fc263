// Multi-GPU optimiser step over NVLink peer memory: gradient reduction, Adam and parameter
// broadcast in ONE kernel (the "reduce-scatter -> Adam -> all-gather" of a data-parallel step,
// fused).  Replaces, for N > 1 ranks on one NVSwitch node, the NCCL all-reduce of the 59*P-float
// gradient arena followed by a full-size Adam pass on every rank (gsdf_slam_b200/mapping.py).
//
// Layout: every rank owns a gradient arena and a parameter arena of identical layout in
// SYMMETRIC memory (peer-mapped; torch.distributed._symmetric_memory hands out the N base
// pointers).  Rank r is responsible for the element range [lo, hi) of the arena (a 1/N shard):
//   g      = sum over the ranks that RENDERED the row of grad_peer[k][e]   (remote loads only from those
//                                                        ranks: per-row rank bitmask, gsr_peer_union_mask)
//   m, v   = shard-local Adam moments (only 1/N of the state exists on each rank)
//   p      = Adam update of the local copy
//   store p to the parameter arena of EVERY rank       (skipped when p is unchanged; ONE multicast store that the
//                                                        NVSwitch replicates (multimem.st, NVLS) when the arena has
//                                                        a multicast mapping, else N-1 remote stores + a local one)
// so each rank reads ~ sum_k (fraction rank k rendered) / N * 236 B and writes (N-1)/N * 236 B per
// Gaussian over NVLink, and touches only 1/N of the Adam state in HBM.  The caller brackets the
// kernel with two cross-rank barriers (gradients visible / parameters landed).
//
// Arithmetic per element is adam_elem of adam.cu (LibTorch op order); the gradient sum runs in
// rank order 0..N-1 on every rank, so all replicas stay bit-identical.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>

#include "gsr_capi.h"
#include "gsr_common.cuh"
#include "gsr_params.cuh"

namespace gsr {

constexpr int kPeerThreads = 256;
constexpr int kPeerChunk = kPeerThreads * 4;  // elements per CTA (one float4 per thread)

struct PeerSegDev {
  long long start;       // first arena element of the segment
  long long numel;
  unsigned width;        // elements per row
  unsigned rule;         // GSR_PEER_RULE_*
  float neg_step_size;
  float inv_sqrt_bc2;
};
struct PeerAdamParams {
  float* param[GSR_MAX_PEERS];
  const float* grad[GSR_MAX_PEERS];
  float* param_mc;       // multicast mapping of the parameter arena (NVLS) or null
  int n_peers;
  int n_segs;
  PeerSegDev seg[GSR_PEER_MAX_SEGMENTS];
  long long lo, hi;      // this rank's shard of the arena (multiples of 4 elements)
  long long chunk;       // > 0: the shard is every n_peers-th chunk of `chunk` elements of [0, hi) instead
  int ctas_per_chunk;
  float* exp_avg;        // [hi - lo] shard-local (chunked: the owned chunks back to back)
  float* exp_avg_sq;     // [hi - lo]
  const int* union_mask; // [rows] > 0 <=> some rank rendered the row
  float beta1, one_minus_beta1, beta2, one_minus_beta2, eps, grad_scale;
  int use_scale;
  unsigned long long* traffic;  // kCount only: [0] remote 16-byte gradient loads, [1] changed 16-byte parameter groups
};

__device__ __forceinline__ void peer_adam_elem(float& p, float g, float& m, float& v, const PeerAdamParams& a,
                                               float neg_step_size, float inv_sqrt_bc2) {
  if (a.use_scale) g = __fmul_rn(g, a.grad_scale);
  const float m1 = __fmul_rn(m, a.beta1);
  const float m2 = __fmaf_rn(a.one_minus_beta1, g, m1);
  const float v1 = __fmul_rn(v, a.beta2);
  const float v2 = __fmaf_rn(a.one_minus_beta2, __fmul_rn(g, g), v1);
  // zero operands take the slow paths of the IEEE sqrt / division (see adam.cu: adam_elem); same bits
  float sq = v2;
  if (v2 != 0.f) sq = __fsqrt_rn(v2);
  const float d0 = __fmul_rn(sq, inv_sqrt_bc2);
  const float d = __fadd_rn(d0, a.eps);
  float q = m2;
  if (m2 != 0.f || !(d > 0.f)) q = __fdiv_rn(m2, d);
  p = __fmaf_rn(neg_step_size, q, p);
  m = m2;
  v = v2;
}

// (4 CTAs per SM = at most 64 registers: the statistics rules pushed the kernel to 72 registers and 3 CTAs, which
// cost 15 % — it is bound by instruction issue and latency at this occupancy, not by HBM)
#ifndef GSR_PEER_MINBLOCKS
#define GSR_PEER_MINBLOCKS 4
#endif
// kPeers: n_peers rounded up to 1 / 2 / 4 / 8 — the loops over the peers are unrolled (all remote loads in flight
// together), and the kernel is issue-bound, so two ranks should not execute the predicated code of eight
template <bool kCount, int kPeers>
__global__ void __launch_bounds__(kPeerThreads, GSR_PEER_MINBLOCKS) peer_adam_kernel(const PeerAdamParams a, int my_rank) {
  // element handled by this thread and its slot in the shard-local moment arrays
  long long e, mo;
  if (a.chunk > 0) {
    const long long c_local = blockIdx.x / a.ctas_per_chunk;
    const long long in_chunk = (long long)(blockIdx.x % a.ctas_per_chunk) * kPeerChunk + 4LL * threadIdx.x;
    if (in_chunk >= a.chunk) return;
    e = (c_local * a.n_peers + my_rank) * a.chunk + in_chunk;
    mo = c_local * a.chunk + in_chunk;
  } else {
    e = a.lo + (long long)blockIdx.x * kPeerChunk + 4LL * threadIdx.x;
    mo = e - a.lo;
  }
  if (e >= a.hi) return;
  // local state first: these HBM reads overlap the index arithmetic and the NVLink round trips below (the
  // statistics segments have no moments; their slots of the moment arrays exist and are simply not written)
  const float4 p_old = *reinterpret_cast<const float4*>(a.param[my_rank] + e);
  float4 m = __ldcs(reinterpret_cast<const float4*>(a.exp_avg + mo));
  float4 v = __ldcs(reinterpret_cast<const float4*>(a.exp_avg_sq + mo));
  // segment of this float4 (segments start on 4-element boundaries, so it lies in exactly one; elements between
  // two segments belong to none)
  int si = 0;
#pragma unroll 1
  for (int k = 1; k < a.n_segs; ++k)
    if (e >= a.seg[k].start) si = k;
  const PeerSegDev S = a.seg[si];
  const long long seg_end = S.start + S.numel;
  if (e >= seg_end) return;
  const bool adam = S.rule == GSR_PEER_RULE_ADAM;
  // rows of the four elements: one wide division per thread, then small offsets
  const unsigned long long off = (unsigned long long)(e - S.start);
  unsigned r0, rem;
  if (off <= 0xffffffffull) {  // 32-bit division: the 64-bit one is ~100 instructions per thread
    r0 = (unsigned)off / S.width;
    rem = (unsigned)off - r0 * S.width;
  } else {
    r0 = (unsigned)(off / S.width);
    rem = (unsigned)(off - (unsigned long long)r0 * S.width);
  }
  // per element: which ranks hold a valid gradient row (mask bits 0..7) — or, for the statistics segments, which
  // ranks rendered the row (bits 8..15)
  const unsigned shift = adam ? 0u : 8u;
  unsigned who[4], want = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned rr = rem + j;
    const unsigned row = r0 + (rr >= S.width) + (rr >= 2 * S.width) + (rr >= 3 * S.width);
    who[j] = (e + j < seg_end) ? (((unsigned)__ldg(a.union_mask + row) >> shift) & 0xffu) : 0u;
    want |= who[j];
  }
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (want) {
    // Only the arenas of the ranks with a valid row are read (a view sees a fraction of the map, so most of the
    // N-1 remote loads disappear, and the ranks need not zero-fill the rows they did not write).  All loads are
    // issued before the first sum (NVLink round trips overlap); the additions run in rank order on the one rank
    // that owns the element, so every replica receives the same value.
    float4 x[kPeers];
#pragma unroll
    for (int k = 0; k < kPeers; ++k)
      if (k < a.n_peers && ((want >> k) & 1u)) x[k] = __ldcs(reinterpret_cast<const float4*>(a.grad[k] + e));
    if (kCount) {
      const int remote = __popc(want & ((1u << a.n_peers) - 1u) & ~(1u << my_rank));
      if (remote) atomicAdd(a.traffic, (unsigned long long)remote);
    }
    if (S.rule != GSR_PEER_RULE_MAX) {   // (warp-uniform except where a warp straddles two segments)
#pragma unroll
      for (int k = 0; k < kPeers; ++k) {
        if (k < a.n_peers && ((want >> k) & 1u)) {
          if ((who[0] >> k) & 1u) g.x += x[k].x;
          if ((who[1] >> k) & 1u) g.y += x[k].y;
          if ((who[2] >> k) & 1u) g.z += x[k].z;
          if ((who[3] >> k) & 1u) g.w += x[k].w;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < kPeers; ++k) {
        if (k < a.n_peers && ((want >> k) & 1u)) {
          if ((who[0] >> k) & 1u) g.x = fmaxf(g.x, x[k].x);
          if ((who[1] >> k) & 1u) g.y = fmaxf(g.y, x[k].y);
          if ((who[2] >> k) & 1u) g.z = fmaxf(g.z, x[k].z);
          if ((who[3] >> k) & 1u) g.w = fmaxf(g.w, x[k].w);
        }
      }
    }
  }
  float4 p = p_old;
  if (adam) {
    peer_adam_elem(p.x, g.x, m.x, v.x, a, S.neg_step_size, S.inv_sqrt_bc2);
    peer_adam_elem(p.y, g.y, m.y, v.y, a, S.neg_step_size, S.inv_sqrt_bc2);
    peer_adam_elem(p.z, g.z, m.z, v.z, a, S.neg_step_size, S.inv_sqrt_bc2);
    peer_adam_elem(p.w, g.w, m.w, v.w, a, S.neg_step_size, S.inv_sqrt_bc2);
    __stcs(reinterpret_cast<float4*>(a.exp_avg + mo), m);
    __stcs(reinterpret_cast<float4*>(a.exp_avg_sq + mo), v);
  } else if (S.rule == GSR_PEER_RULE_SUM) {
    // (elements whose row nobody rendered keep their bits: no "+ 0")
    if (who[0]) p.x = p.x + g.x;
    if (who[1]) p.y = p.y + g.y;
    if (who[2]) p.z = p.z + g.z;
    if (who[3]) p.w = p.w + g.w;
  } else {
    if (who[0]) p.x = fmaxf(p.x, g.x);
    if (who[1]) p.y = fmaxf(p.y, g.y);
    if (who[2]) p.z = fmaxf(p.z, g.z);
    if (who[3]) p.w = fmaxf(p.w, g.w);
  }
  // every replica holds p_old already: values that did not change (rows whose moments and gradient are
  // all zero) need not travel
  const bool changed = __float_as_uint(p.x) != __float_as_uint(p_old.x) || __float_as_uint(p.y) != __float_as_uint(p_old.y) ||
                       __float_as_uint(p.z) != __float_as_uint(p_old.z) || __float_as_uint(p.w) != __float_as_uint(p_old.w);
  if (changed) {
    if (kCount) atomicAdd(a.traffic + 1, 1ull);
    if (a.param_mc != nullptr) {
      // one store; the NVSwitch replicates it into the arena of every rank (this one included)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.param_mc + e), "f"(p.x),
                   "f"(p.y), "f"(p.z), "f"(p.w)
                   : "memory");
    } else {
#pragma unroll
      for (int k = 0; k < kPeers; ++k)
        if (k < a.n_peers) *reinterpret_cast<float4*>(a.param[k] + e) = p;
    }
  }
}

// union[row] = bit k set <=> rank k rendered the row (radii_peer[k][row] > 0)
struct PeerMaskParams {
  const int* radii[GSR_MAX_PEERS];
  int n_peers;
};
__global__ void __launch_bounds__(256) peer_union_mask_kernel(const PeerMaskParams a, long long rows,
                                                              int* __restrict__ union_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  int m = 0;
  for (int k = 0; k < a.n_peers; ++k) m |= (a.radii[k][i] > 0) ? (0x101 << k) : 0;
  union_out[i] = m;
}

// the same from one-byte "rendered" flags, four rows per thread (a quarter of the remote traffic)
struct PeerFlagParams {
  const uint8_t* flags[GSR_MAX_PEERS];
  int n_peers;
};
__global__ void __launch_bounds__(256) peer_rank_mask_u8_kernel(const PeerFlagParams a, long long rows,
                                                                int* __restrict__ mask_out) {
  const long long i = 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= rows) return;
  int m[4] = {0, 0, 0, 0};
  if (i + 3 < rows) {
    uint32_t w[GSR_MAX_PEERS];
#pragma unroll
    for (int k = 0; k < GSR_MAX_PEERS; ++k)
      if (k < a.n_peers) w[k] = *reinterpret_cast<const uint32_t*>(a.flags[k] + i);
#pragma unroll
    for (int k = 0; k < GSR_MAX_PEERS; ++k) {
      if (k < a.n_peers) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t b = (w[k] >> (8 * j)) & 0xffu;   // bit 0: rendered, bit 1: gradient row valid
          m[j] |= ((b & 2u) ? (1 << k) : 0) | ((b & 1u) ? (0x100 << k) : 0);
        }
      }
    }
    *reinterpret_cast<int4*>(mask_out + i) = make_int4(m[0], m[1], m[2], m[3]);
  } else {
    for (int j = 0; i + j < rows; ++j) {
      int mm = 0;
      for (int k = 0; k < a.n_peers; ++k) {
        const uint32_t b = a.flags[k][i + j];
        mm |= ((b & 2u) ? (1 << k) : 0) | ((b & 1u) ? (0x100 << k) : 0);
      }
      mask_out[i + j] = mm;
    }
  }
}


// ---- densification statistics across ranks ------------------------------------
// The reference updates, per VIEW (src/gaussian_mapper.cpp:443-447, src/gaussian_model.cpp:667-681), for the
// rows with radii > 0:  max_radii2D = max(max_radii2D, radii);  xyz_gradient_accum += |dL_dmean2D.xy|;
// denom += 1.  The norm is non-linear, so it is taken per keyframe on the rank that rendered it
// (peer_local_stats_kernel, into the rank's per-step delta arrays [accum | denom | max_radii], peer-mapped)
// and the deltas of all ranks are combined afterwards (peer_stats_exchange_kernel): every rank pulls, for
// each row, the deltas of the ranks that rendered it — in rank order, so the replicated statistics stay
// bit-identical on all ranks — and applies sum / sum / max to its own replica.  The same kernel emits the
// per-row rank bitmask gsr_adam_step_peer consumes (it replaces gsr_peer_rank_mask_u8 in that case).
__global__ void __launch_bounds__(256)
peer_local_stats_kernel(long long P, long long stride, const int* __restrict__ radii, const float* __restrict__ dmean2D,
                        float* __restrict__ stats /* [3][stride] or null */, uint8_t* __restrict__ flags,
                        const int* __restrict__ grad_rows /* or null: every rendered row's gradient is valid */) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= P) return;
  const int r = radii[i];
  if (r <= 0) return;
  // bit 0: this rank rendered the row (statistics); bit 1: its gradient row is valid (gsr_backward_args.grad_rows)
  flags[i] = (uint8_t)(1u | ((grad_rows == nullptr || grad_rows[i] > 0) ? 2u : 0u));
  if (stats == nullptr) return;
  const float gx = dmean2D[3 * i], gy = dmean2D[3 * i + 1];
  stats[i] = stats[i] + sqrtf(fmaf(gy, gy, gx * gx));   // same expression as densification_stats_kernel
  stats[stride + i] = stats[stride + i] + 1.0f;
  stats[2 * stride + i] = fmaxf(stats[2 * stride + i], (float)r);
}

struct PeerStatsParams {
  const uint8_t* flags[GSR_MAX_PEERS];
  const float* stats[GSR_MAX_PEERS];
  int n_peers;
};
// Four rows per thread: one 4-byte flag word per peer (a quarter of the NVLink requests of a byte per row) and
// 16-byte statistics loads from the ranks that rendered any of the four rows.
__global__ void __launch_bounds__(256)
peer_stats_exchange_kernel(const PeerStatsParams a, long long P, long long stride, int* __restrict__ mask_out,
                           float* __restrict__ max_radii2D, float* __restrict__ accum, float* __restrict__ denom) {
  const long long i = 4 * ((long long)blockIdx.x * 256 + threadIdx.x);
  if (i >= P) return;
  const int nrow = (int)min(4LL, P - i);
  uint32_t w[GSR_MAX_PEERS];
  unsigned any = 0;   // bit k: rank k rendered one of the rows
#pragma unroll
  for (int k = 0; k < GSR_MAX_PEERS; ++k) {
    if (k < a.n_peers) {
      w[k] = *reinterpret_cast<const uint32_t*>(a.flags[k] + i);   // the flag buffers are padded by 16 bytes
      if (nrow < 4) w[k] &= (1u << (8 * nrow)) - 1u;
      any |= w[k] ? (1u << k) : 0u;
    }
  }
  int m[4] = {0, 0, 0, 0};    // ranks that RENDERED the row (flag bit 0): whose statistics count
  int mg[4] = {0, 0, 0, 0};   // ranks whose GRADIENT row is valid (flag bit 1): what gsr_adam_step_peer pulls
#pragma unroll
  for (int k = 0; k < GSR_MAX_PEERS; ++k) {
    if (k < a.n_peers) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        m[j] |= ((w[k] >> (8 * j)) & 0xffu) ? (1 << k) : 0;
        mg[j] |= ((w[k] >> (8 * j)) & 0x02u) ? (1 << k) : 0;
      }
    }
  }
  if (nrow == 4) {
    *reinterpret_cast<int4*>(mask_out + i) = make_int4(mg[0] | (m[0] << 8), mg[1] | (m[1] << 8), mg[2] | (m[2] << 8),
                                                       mg[3] | (m[3] << 8));
  } else {
    for (int j = 0; j < nrow; ++j) mask_out[i + j] = mg[j] | (m[j] << 8);
  }
  if (any == 0) return;
  // all remote loads first (NVLink round trips overlap), then the combination in rank order
  float4 da[GSR_MAX_PEERS], dd[GSR_MAX_PEERS], dr[GSR_MAX_PEERS];
#pragma unroll
  for (int k = 0; k < GSR_MAX_PEERS; ++k) {
    if (k < a.n_peers && ((any >> k) & 1u)) {
      // stride is a multiple of 4 and the peer arrays are 16-byte aligned: the 4-row group never straddles
      da[k] = *reinterpret_cast<const float4*>(a.stats[k] + i);
      dd[k] = *reinterpret_cast<const float4*>(a.stats[k] + stride + i);
      dr[k] = *reinterpret_cast<const float4*>(a.stats[k] + 2 * stride + i);
    }
  }
  float sa[4] = {0.f, 0.f, 0.f, 0.f}, sd[4] = {0.f, 0.f, 0.f, 0.f}, mr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < GSR_MAX_PEERS; ++k) {
    if (k < a.n_peers && ((any >> k) & 1u)) {
      const float xa[4] = {da[k].x, da[k].y, da[k].z, da[k].w};
      const float xd[4] = {dd[k].x, dd[k].y, dd[k].z, dd[k].w};
      const float xr[4] = {dr[k].x, dr[k].y, dr[k].z, dr[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((m[j] >> k) & 1) {     // only the ranks that rendered THIS row (others hold 0 there anyway)
          sa[j] += xa[j];
          sd[j] += xd[j];
          mr[j] = fmaxf(mr[j], xr[j]);
        }
      }
    }
  }
  for (int j = 0; j < nrow; ++j) {
    if (m[j] == 0) continue;
    accum[i + j] = accum[i + j] + sa[j];
    denom[i + j] = denom[i + j] + sd[j];
    max_radii2D[i + j] = fmaxf(max_radii2D[i + j], mr[j]);
  }
}

}  // namespace gsr

using namespace gsr;

extern "C" {

int gsr_peer_union_mask(const int32_t* const* radii_peers, int n_peers, int64_t rows, int32_t* union_out,
                        void* stream) {
  if (n_peers < 1 || n_peers > GSR_MAX_PEERS || rows < 0 || !radii_peers || (rows > 0 && !union_out))
    return GSR_ERR_BAD_ARG;
  if (rows == 0) return GSR_OK;
  PeerMaskParams a;
  a.n_peers = n_peers;
  for (int k = 0; k < n_peers; ++k) {
    if (!radii_peers[k]) return GSR_ERR_BAD_ARG;
    a.radii[k] = radii_peers[k];
  }
  peer_union_mask_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, rows, union_out);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

int gsr_peer_rank_mask_u8(const uint8_t* const* flags_peers, int n_peers, int64_t rows, int32_t* mask_out,
                          void* stream) {
  if (n_peers < 1 || n_peers > GSR_MAX_PEERS || rows < 0 || !flags_peers || (rows > 0 && !mask_out))
    return GSR_ERR_BAD_ARG;
  if (rows == 0) return GSR_OK;
  PeerFlagParams a;
  a.n_peers = n_peers;
  for (int k = 0; k < n_peers; ++k) {
    if (!flags_peers[k] || ((uintptr_t)flags_peers[k] & 3)) return GSR_ERR_BAD_ARG;
    a.flags[k] = flags_peers[k];
  }
  if ((uintptr_t)mask_out & 15) return GSR_ERR_BAD_ARG;
  const long long threads = (rows + 3) / 4;
  peer_rank_mask_u8_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, rows, mask_out);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

int gsr_peer_local_stats(int64_t P, const int32_t* radii, const float* dL_dmeans2D, float* stats_local,
                         int64_t stats_stride, uint8_t* flags, const int32_t* grad_rows, void* stream) {
  if (P < 0) return GSR_ERR_BAD_ARG;
  if (P == 0) return GSR_OK;
  if (!radii || !flags || (stats_local && (!dL_dmeans2D || stats_stride < P))) return GSR_ERR_BAD_ARG;
  peer_local_stats_kernel<<<(unsigned)((P + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      P, stats_stride, radii, dL_dmeans2D, stats_local, flags, grad_rows);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

int gsr_peer_stats_exchange(const uint8_t* const* flags_peers, const float* const* stats_peers, int64_t stats_stride,
                            int n_peers, int64_t P, int32_t* mask_out, float* max_radii2D, float* xyz_gradient_accum,
                            float* denom, void* stream) {
  if (n_peers < 1 || n_peers > GSR_MAX_PEERS || P < 0 || !flags_peers || !stats_peers) return GSR_ERR_BAD_ARG;
  if (P == 0) return GSR_OK;
  if (!mask_out || !max_radii2D || !xyz_gradient_accum || !denom || stats_stride < P || (stats_stride & 3) ||
      ((uintptr_t)mask_out & 15))
    return GSR_ERR_BAD_ARG;
  PeerStatsParams a;
  a.n_peers = n_peers;
  for (int k = 0; k < n_peers; ++k) {
    if (!flags_peers[k] || !stats_peers[k] || ((uintptr_t)flags_peers[k] & 3) || ((uintptr_t)stats_peers[k] & 15))
      return GSR_ERR_BAD_ARG;
    a.flags[k] = flags_peers[k];
    a.stats[k] = stats_peers[k];
  }
  const long long threads = (P + 3) / 4;
  peer_stats_exchange_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, P, stats_stride, mask_out, max_radii2D, xyz_gradient_accum, denom);
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

int gsr_adam_step_peer(const gsr_peer_adam_args* in) {
  if (!in || in->n_peers < 1 || in->n_peers > GSR_MAX_PEERS || in->rank < 0 || in->rank >= in->n_peers ||
      in->n_segments < 1 || in->n_segments > GSR_PEER_MAX_SEGMENTS || in->shard_lo < 0 ||
      in->shard_hi < in->shard_lo || (in->shard_lo & 3) || !in->union_mask || in->step < 1)
    return GSR_ERR_BAD_ARG;
  if (in->shard_hi == in->shard_lo) return GSR_OK;
  if (!in->exp_avg || !in->exp_avg_sq) return GSR_ERR_BAD_ARG;
  PeerAdamParams a;
  a.n_peers = in->n_peers;
  for (int k = 0; k < in->n_peers; ++k) {
    if (!in->param_peers[k] || !in->grad_peers[k]) return GSR_ERR_BAD_ARG;
    if (((uintptr_t)in->param_peers[k] | (uintptr_t)in->grad_peers[k]) & 15) return GSR_ERR_BAD_ARG;
    a.param[k] = in->param_peers[k];
    a.grad[k] = in->grad_peers[k];
  }
  if (((uintptr_t)in->exp_avg | (uintptr_t)in->exp_avg_sq | (uintptr_t)in->param_multicast) & 15) return GSR_ERR_BAD_ARG;
  a.param_mc = in->param_multicast;
  const double bc1 = 1.0 - std::pow(in->beta1, (double)in->step);
  const double bc2 = 1.0 - std::pow(in->beta2, (double)in->step);
  a.n_segs = in->n_segments;
  long long prev_end = 0;
  for (int k = 0; k < in->n_segments; ++k) {
    const gsr_peer_segment& s = in->segments[k];
    if (s.start < prev_end || (s.start & 3) || s.numel < 0 || s.width < 1 || s.width > 4096) return GSR_ERR_BAD_ARG;
    a.seg[k].start = s.start;
    a.seg[k].numel = s.numel;
    a.seg[k].width = (unsigned)s.width;
    if (s.rule < GSR_PEER_RULE_ADAM || s.rule > GSR_PEER_RULE_MAX) return GSR_ERR_BAD_ARG;
    a.seg[k].rule = (unsigned)s.rule;
    a.seg[k].neg_step_size = (float)(-(s.lr / bc1));
    a.seg[k].inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
    prev_end = s.start + s.numel;
  }
  a.lo = in->shard_lo;
  a.hi = in->shard_hi;
  a.exp_avg = in->exp_avg;
  a.exp_avg_sq = in->exp_avg_sq;
  a.union_mask = in->union_mask;
  a.beta1 = (float)in->beta1;
  a.beta2 = (float)in->beta2;
  a.one_minus_beta1 = (float)(1.0 - in->beta1);
  a.one_minus_beta2 = (float)(1.0 - in->beta2);
  a.eps = (float)in->eps;
  a.grad_scale = in->grad_scale;
  a.use_scale = in->grad_scale != 1.0f;
  a.chunk = in->shard_chunk;
  a.ctas_per_chunk = 1;
  long long ctas = (a.hi - a.lo + kPeerChunk - 1) / kPeerChunk;
  if (a.chunk > 0) {
    // interleaved sharding: rank r owns the chunks c with c % n_peers == r of [0, shard_hi)
    if ((a.chunk & 3) || in->shard_lo != 0) return GSR_ERR_BAD_ARG;
    a.ctas_per_chunk = (int)((a.chunk + kPeerChunk - 1) / kPeerChunk);
    const long long n_chunks = (a.hi + a.chunk - 1) / a.chunk;
    const long long owned = n_chunks > in->rank ? (n_chunks - in->rank + in->n_peers - 1) / in->n_peers : 0;
    ctas = owned * a.ctas_per_chunk;
    if (ctas == 0) return GSR_OK;
  }
  a.traffic = in->traffic_counters;
  cudaStream_t st = static_cast<cudaStream_t>(in->stream);
#define GSR_PEER_LAUNCH(N)                                                                              \
  do {                                                                                                  \
    if (a.traffic != nullptr) peer_adam_kernel<true, N><<<(unsigned)ctas, kPeerThreads, 0, st>>>(a, in->rank);  \
    else peer_adam_kernel<false, N><<<(unsigned)ctas, kPeerThreads, 0, st>>>(a, in->rank);               \
  } while (0)
  if (in->n_peers <= 1) GSR_PEER_LAUNCH(1);
  else if (in->n_peers <= 2) GSR_PEER_LAUNCH(2);
  else if (in->n_peers <= 4) GSR_PEER_LAUNCH(4);
  else GSR_PEER_LAUNCH(8);
#undef GSR_PEER_LAUNCH
  GSR_LAUNCHED();
  return (int)cudaGetLastError();
}

}  // extern "C"
