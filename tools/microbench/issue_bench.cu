// Micro-benchmark (developer tool): issue-slot / pipe throughput of FFMA vs the packed FFMA2
// (fma.rn.f32x2, sm_100+) alone and mixed with ALU / MUFU work.  Motivation: the blend kernels are
// issue-bound (ncu r1: 76-82 % issue-active), so the question is whether packing two pixels' FP32
// operations into one instruction frees issue slots for the rest of the instruction mix.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_bench issue_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2(float a) {
  float d;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a));
  return d;
}
__device__ __forceinline__ float fmnmx(float a, float b) {
  float d;
  asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float fmxmx(float a, float b) {
  float d;
  asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) {
  unsigned d;
  asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

constexpr int CH = 8;  // independent chains per thread

// mode 0: FFMA x CH ; 1: FFMA2 x CH ; 2: FFMA x CH + ALU(min) x CH ; 3: FFMA2 x CH + ALU x CH
// 4: MUFU x CH ; 5: FFMA x CH + MUFU x (CH/4) ; 6: FFMA2 x CH + MUFU x (CH/4) ; 7: FFMA2 x (CH/2) + ALU x CH
// 8: ALU(min) x CH only ; 9: FFMA x CH + LOP x CH ; 10: FFMA2 x CH + ALU x CH + MUFU x (CH/4)
template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float a, float b) {
  float x[CH], m[CH];
  unsigned long long y[CH];
  unsigned u[CH];
  float e[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    x[k] = threadIdx.x * 1e-3f + k;
    y[k] = pk(x[k], x[k] + 0.5f);
    m[k] = x[k] * 0.25f;
    u[k] = threadIdx.x + k;
    e[k] = -1.0f - k * 0.1f;
  }
  const unsigned long long a2 = pk(a, a), b2 = pk(b, b);
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        if (MODE == 0 || MODE == 2 || MODE == 5 || MODE == 9) x[k] = ffma(x[k], a, b);
        if (MODE == 1 || MODE == 3 || MODE == 6 || MODE == 10) y[k] = ffma2(y[k], a2, b2);
        if (MODE == 7 && (k & 1) == 0) y[k] = ffma2(y[k], a2, b2);
        if (MODE == 2 || MODE == 3 || MODE == 7 || MODE == 8 || MODE == 10) m[k] = (r & 1) ? fmxmx(m[k], x[(k + 1) % CH]) : fmnmx(m[k], x[(k + 1) % CH]);
        if (MODE == 9) u[k] = lop(u[k], u[(k + 1) % CH]);
        if (MODE == 4) e[k] = ex2(e[k]);
        if ((MODE == 5 || MODE == 6 || MODE == 10) && (k & 3) == 0) e[k] = ex2(e[k]);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    s += x[k] + m[k] + e[k] + __uint_as_float((unsigned)y[k]) + __uint_as_float((unsigned)(y[k] >> 32)) + (float)u[k];
  }
  if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, double inst_per_iter_per_thread, float* d_out, int ctas_per_sm) {
  const int iters = 4096;
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  bench<MODE><<<grid, 256>>>(d_out, 64, 0.999f, 0.001f);
  CK(cudaDeviceSynchronize());
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    bench<MODE><<<grid, 256>>>(d_out, iters, 0.999f, 0.001f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  // warp-instructions issued per SMSP
  const double warps_per_smsp = ctas_per_sm * 8 / 4.0;
  const double winst = inst_per_iter_per_thread * iters * warps_per_smsp;
  printf("%-44s ctas/SM=%d  %.3f ms  -> %.3f ns per warp-inst per SMSP (at 1.965 GHz: %.2f clk/inst)\n", name, ctas_per_sm,
         best, best * 1e6 / winst, best * 1e6 / winst * 1.965);
}

int main() {
  float* d_out; CK(cudaMalloc(&d_out, 1024));
  for (int c : {1, 2, 4}) {
    run<0>("FFMA x8", 32, d_out, c);
    run<1>("FFMA2 x8", 32, d_out, c);
    run<8>("FMNMX x8", 32, d_out, c);
    run<2>("FFMA x8 + FMNMX x8", 64, d_out, c);
    run<3>("FFMA2 x8 + FMNMX x8", 64, d_out, c);
    run<7>("FFMA2 x4 + FMNMX x8", 48, d_out, c);
    run<9>("FFMA x8 + LOP x8", 64, d_out, c);
    run<4>("MUFU.EX2 x8", 32, d_out, c);
    run<5>("FFMA x8 + MUFU x2", 40, d_out, c);
    run<6>("FFMA2 x8 + MUFU x2", 40, d_out, c);
    run<10>("FFMA2 x8 + FMNMX x8 + MUFU x2", 72, d_out, c);
  }
  return 0;
}
