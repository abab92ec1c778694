// Pipe micro-benchmarks for sm_100a (B200): issue/throughput of FFMA, packed FFMA2 / FADD2 / FMUL2,
// MUFU, and the Gram kernel's FFMA2 : MUFU : LDS mix.  Build: make -C tools ; run on the GPU box.
// Prints one line per probe: name, Gop/s per SM-clock-independent totals and lane-ops / clk / SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

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
    } else if (MODE == 11) {  // cvt.rna.tf32.f32
#pragma unroll
      for (int k = 0; k < kChains; ++k) {
        uint32_t u, v;
        asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(a[k].x));
        asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(v) : "f"(a[k].y));
        a[k].x = __uint_as_float(u) + 1e-3f; a[k].y = __uint_as_float(v) + 1e-3f;
      }
    } else if (MODE == 12) {  // MUFU.SQRT + MUFU.EX2 + FMNMX chain (the weight)
#pragma unroll
      for (int k = 0; k < kChains; ++k) { a[k].x = fmaxf(ex2a(-sqa(a[k].x)), 0.25f) + 1.f; a[k].y = fmaxf(ex2a(-sqa(a[k].y)), 0.25f) + 1.f; }
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

// The Gram kernel's inner loop with its real operand pattern (48 accumulator pairs, 4 broadcast
// weights, 12 keypoint-term pairs), built up piece by piece: FMA only (LEVEL 0), + the 7 broadcast
// LDS.128 per keypoint row (1), + the weight arithmetic without MUFU (2), + MUFU.SQRT/EX2 (3).
template <int LEVEL>
__global__ void __launch_bounds__(32) gram_mix(int iters, float seed, float *sink) {
  __shared__ __align__(16) float rows[64 * 28];
  for (int i = threadIdx.x; i < 64 * 28; i += 32) rows[i] = seed + (i % 13) * 0.01f;
  __syncwarp();
  float2 acc[4][12];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int t = 0; t < 12; ++t) acc[r][t] = make_float2(seed + r, seed - t);
  float ax[4], ay[4], w2[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { ax[r] = seed + threadIdx.x * 0.01f + r; ay[r] = seed - r; w2[r] = 0.999f + seed + r * 1e-4f; }
  float4 p[6];
#pragma unroll
  for (int v = 0; v < 6; ++v) p[v] = make_float4(seed + v, seed + 0.1f * v, seed - v, seed + 2 * v);
  const float4 *rv = reinterpret_cast<const float4 *>(rows);
  for (int it = 0; it < iters; ++it) {
    const float4 *row = rv + (it & 63) * 7;
    if (LEVEL >= 2) {
      const float4 q = (LEVEL >= 1) ? row[6] : p[0];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float dx = ax[r] - q.x, dy = ay[r] - q.z;
        const float d2 = fmaf(dy, dy, dx * dx);
        w2[r] = (LEVEL >= 3) ? fmaxf(ex2a(-sqa(d2)), 0.25f) : fmaxf(d2, 0.25f);
      }
    }
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      const float4 pv = (LEVEL >= 1) ? row[v] : p[v];
      const float2 plo = make_float2(pv.x, pv.y), phi = make_float2(pv.z, pv.w);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float2 ww = make_float2(w2[r], w2[r]);
        acc[r][2 * v + 0] = __ffma2_rn(ww, plo, acc[r][2 * v + 0]);
        acc[r][2 * v + 1] = __ffma2_rn(ww, phi, acc[r][2 * v + 1]);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int t = 0; t < 12; ++t) s += acc[r][t].x + acc[r][t].y;
  if (s == 123456.789f) *sink = s;
}

// FMA-only operand-pattern study for the register tile 4 cells x 24 terms (no memory, no MUFU).
//   PAT 0: FFMA2, pairs along terms, keypoint-vector outer / cell inner   (w scalar broadcast)
//   PAT 1: FFMA2, pairs along terms, cell outer / vector inner
//   PAT 2: FFMA2, pairs along cells (w pair, p scalar broadcast), term outer
//   PAT 3: scalar FFMA, term outer / cell inner
//   PAT 4: scalar FFMA, cell outer / term inner
//   PAT 5: FFMA2 pairs along terms, 2 cells x 24 terms tile (R = 2)
template <int PAT>
__global__ void __launch_bounds__(32) tile_pat(int iters, float seed, float *sink) {
  float2 acc[4][12];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int t = 0; t < 12; ++t) acc[r][t] = make_float2(seed + r, seed - t);
  float w[4];
  float2 p[12];
#pragma unroll
  for (int r = 0; r < 4; ++r) w[r] = 0.999f + seed + r * 1e-4f + threadIdx.x * 1e-6f;
#pragma unroll
  for (int t = 0; t < 12; ++t) p[t] = make_float2(seed + t * 1e-3f, seed - t * 1e-3f);
  for (int it = 0; it < iters; ++it) {
    if (PAT == 0) {
#pragma unroll
      for (int t = 0; t < 12; ++t)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][t] = __ffma2_rn(make_float2(w[r], w[r]), p[t], acc[r][t]);
    } else if (PAT == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int t = 0; t < 12; ++t) acc[r][t] = __ffma2_rn(make_float2(w[r], w[r]), p[t], acc[r][t]);
    } else if (PAT == 2) {
      // acc[h][t] here = (cell 2h, cell 2h+1) for term t (24 terms: both halves of p[t])
      float2 (*a2)[24] = reinterpret_cast<float2 (*)[24]>(&acc[0][0]);
#pragma unroll
      for (int t = 0; t < 12; ++t) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 ww = make_float2(w[2 * h], w[2 * h + 1]);
          a2[h][2 * t] = __ffma2_rn(ww, make_float2(p[t].x, p[t].x), a2[h][2 * t]);
          a2[h][2 * t + 1] = __ffma2_rn(ww, make_float2(p[t].y, p[t].y), a2[h][2 * t + 1]);
        }
      }
    } else if (PAT == 3) {
#pragma unroll
      for (int t = 0; t < 12; ++t)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[r][t].x = fmaf(w[r], p[t].x, acc[r][t].x);
          acc[r][t].y = fmaf(w[r], p[t].y, acc[r][t].y);
        }
    } else if (PAT == 4) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int t = 0; t < 12; ++t) {
          acc[r][t].x = fmaf(w[r], p[t].x, acc[r][t].x);
          acc[r][t].y = fmaf(w[r], p[t].y, acc[r][t].y);
        }
    } else if (PAT == 5) {
#pragma unroll
      for (int rep = 0; rep < 2; ++rep)
#pragma unroll
        for (int t = 0; t < 12; ++t)
#pragma unroll
          for (int r = 0; r < 2; ++r) acc[r][t] = __ffma2_rn(make_float2(w[r], w[r]), p[t], acc[r][t]);
    }
    // keep the operands loop-variant so nothing is hoisted (cheap: 2 instructions per iteration)
    w[it & 3] += 1e-7f;
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int t = 0; t < 12; ++t) s += acc[r][t].x + acc[r][t].y;
  if (s == 123456.789f) *sink = s;
}

template <int PAT>
static void run_pat(const char *name, int warps_per_sm, int sms, double clk_ghz, float *sink) {
  const int iters = 1 << 14, blocks = sms * warps_per_sm;
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  tile_pat<PAT><<<blocks, 32>>>(iters, 0.f, sink);
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CHECK(cudaEventRecord(e0));
    tile_pat<PAT><<<blocks, 32>>>(iters, 0.f, sink);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double total = 96.0 * iters * (double)blocks * 32;
  printf("%-44s %2d warps/SM %8.3f ms  %7.1f FMA lanes/clk/SM\n", name, warps_per_sm, best,
         total / (best * 1e-3) / (sms * clk_ghz * 1e9));
}

template <int LEVEL>
static void run_mix(const char *name, int warps_per_sm, int sms, double clk_ghz, float *sink) {
  const int iters = 1 << 14, blocks = sms * warps_per_sm;
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  gram_mix<LEVEL><<<blocks, 32>>>(iters, 0.f, sink);
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CHECK(cudaEventRecord(e0));
    gram_mix<LEVEL><<<blocks, 32>>>(iters, 0.f, sink);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  const double total = 96.0 * iters * (double)blocks * 32;     // useful FMA lane-ops
  const double per_s = total / (best * 1e-3);
  printf("%-34s %2d warps/SM %8.3f ms  %7.1f useful FMA lanes/clk/SM\n", name, warps_per_sm, best,
         per_s / (sms * clk_ghz * 1e9));
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
  run<11>("cvt.rna.tf32.f32 (+FADD)", 32, sms, ghz, sink);
  run<12>("weight: SQRT+EX2+FMNMX+FADD", 32, sms, ghz, sink);
  run<7>("gram mix (24 useful FMA)", 24, sms, ghz, sink);
  for (int w : {4, 12}) {
    run_pat<0>("tile: FFMA2 term-pairs, vector outer", w, sms, ghz, sink);
    run_pat<1>("tile: FFMA2 term-pairs, cell outer", w, sms, ghz, sink);
    run_pat<2>("tile: FFMA2 cell-pairs, p broadcast", w, sms, ghz, sink);
    run_pat<3>("tile: scalar FFMA, term outer", w, sms, ghz, sink);
    run_pat<4>("tile: scalar FFMA, cell outer", w, sms, ghz, sink);
    run_pat<5>("tile: FFMA2 term-pairs, R=2 tile x2", w, sms, ghz, sink);
  }
  for (int w : {12}) {
    run_mix<0>("loop: FFMA2 only", w, sms, ghz, sink);
    run_mix<1>("loop: + LDS.128 rows", w, sms, ghz, sink);
    run_mix<2>("loop: + weight arithmetic", w, sms, ghz, sink);
    run_mix<3>("loop: + MUFU.SQRT/EX2", w, sms, ghz, sink);
  }
  return 0;
}
