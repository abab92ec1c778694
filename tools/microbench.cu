// Pipe micro-benchmarks for sm_100a (B200): issue/throughput of FFMA, packed FFMA2 / FADD2 / FMUL2,
// MUFU, and the Gram kernel's FFMA2 : MUFU : LDS mix.  Build: make -C tools ; run on the GPU box.
// Prints one line per probe: name, Gop/s per SM-clock-independent totals and lane-ops / clk / SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int kThreads = 256;
constexpr int kChains = 16;

__device__ __forceinline__ float ex2a(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqa(float x) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcpa(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

template <int MODE>
__global__ void __launch_bounds__(kThreads) probe(int iters, float seed, float *sink) {
  float2 a[kChains];
#pragma unroll
  for (int k = 0; k < kChains; ++k) a[k] = make_float2(seed + threadIdx.x + k, seed - k);
  const float2 m = make_float2(0.999f + seed, 1.001f + seed), c = make_float2(1e-3f, 2e-3f);
  const float ms = 0.999f + seed;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // scalar FFMA, 32 independent chains
#pragma unroll
      for (int k = 0; k < kChains; ++k) { a[k].x = fmaf(a[k].x, m.x, c.x); a[k].y = fmaf(a[k].y, m.y, c.y); }
    } else if (MODE == 1) {   // FFMA2, 16 independent chains
#pragma unroll
      for (int k = 0; k < kChains; ++k) a[k] = __ffma2_rn(a[k], m, c);
    } else if (MODE == 2) {   // FFMA2 with a scalar-broadcast multiplier (the Gram kernel's form)
#pragma unroll
      for (int k = 0; k < kChains; ++k) a[k] = __ffma2_rn(make_float2(ms, ms), a[k], c);
    } else if (MODE == 3) {   // FADD2
#pragma unroll
      for (int k = 0; k < kChains; ++k) a[k] = __fadd2_rn(a[k], c);
    } else if (MODE == 4) {   // FMUL2
#pragma unroll
      for (int k = 0; k < kChains; ++k) a[k] = __fmul2_rn(a[k], m);
    } else if (MODE == 5) {   // MUFU.EX2
#pragma unroll
      for (int k = 0; k < kChains; ++k) { a[k].x = ex2a(a[k].x); a[k].y = ex2a(a[k].y); }
    } else if (MODE == 6) {   // MUFU.SQRT
#pragma unroll
      for (int k = 0; k < kChains; ++k) { a[k].x = sqa(a[k].x); a[k].y = sqa(a[k].y); }
    } else if (MODE == 7) {   // Gram mix per cell-keypoint: 12 FFMA2 + 2 packed/2 + 2 MUFU + 1 FMNMX
#pragma unroll
      for (int k = 0; k < 12; ++k) a[k] = __ffma2_rn(make_float2(a[15].x, a[15].x), a[k], c);
      a[12] = __fadd2_rn(a[12], c);
      a[13] = __ffma2_rn(a[12], a[12], a[13]);
      a[15].x = fmaxf(ex2a(-sqa(a[13].x)), 0.25f);
    } else if (MODE == 8) {   // FADD.RM (round-down add) scalar
#pragma unroll
      for (int k = 0; k < kChains; ++k) { a[k].x = __fadd_rd(a[k].x, c.x); a[k].y = __fadd_rd(a[k].y, c.y); }
    } else if (MODE == 9) {   // MUFU.RCP
#pragma unroll
      for (int k = 0; k < kChains; ++k) { a[k].x = rcpa(a[k].x); a[k].y = rcpa(a[k].y); }
    } else if (MODE == 10) {  // DFMA
      double *d = reinterpret_cast<double *>(a);
#pragma unroll
      for (int k = 0; k < kChains; ++k) d[k] = fma(d[k], 0.999, 1e-3);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kChains; ++k) s += a[k].x + a[k].y;
  if (s == 123456.789f) *sink = s;
}

template <int MODE>
static void run(const char *name, double lane_ops_per_iter, int sms, double clk_ghz, float *sink) {
  const int iters = 1 << 14, blocks = sms * 8;
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  probe<MODE><<<blocks, kThreads>>>(iters, 0.f, sink);
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CHECK(cudaEventRecord(e0));
    probe<MODE><<<blocks, kThreads>>>(iters, 0.f, sink);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double total = lane_ops_per_iter * iters * (double)blocks * kThreads;
  const double per_s = total / (best * 1e-3);
  printf("%-28s %8.3f ms  %9.2f Tlane-op/s  %7.1f lane-ops/clk/SM (at %.3f GHz)\n", name, best, per_s / 1e12,
         per_s / (sms * clk_ghz * 1e9), clk_ghz);
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  int clk_khz = 0;
  CHECK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double ghz = clk_khz * 1e-6;
  printf("%s: %d SMs, max clock %.3f GHz\n", prop.name, prop.multiProcessorCount, ghz);
  float *sink;
  CHECK(cudaMalloc(&sink, 4));
  const int sms = prop.multiProcessorCount;
  run<0>("FFMA (scalar, fp32 results)", 32, sms, ghz, sink);
  run<1>("FFMA2 (fp32 results)", 32, sms, ghz, sink);
  run<2>("FFMA2 bcast (fp32 results)", 32, sms, ghz, sink);
  run<3>("FADD2 (fp32 results)", 32, sms, ghz, sink);
  run<4>("FMUL2 (fp32 results)", 32, sms, ghz, sink);
  run<5>("MUFU.EX2", 32, sms, ghz, sink);
  run<6>("MUFU.SQRT", 32, sms, ghz, sink);
  run<9>("MUFU.RCP", 32, sms, ghz, sink);
  run<8>("FADD.RM (scalar)", 32, sms, ghz, sink);
  run<10>("DFMA", 16, sms, ghz, sink);
  run<7>("gram mix (24 useful FMA)", 24, sms, ghz, sink);
  return 0;
}
