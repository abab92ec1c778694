// Design-space lab for the Gram kernel (K1): times kernel variants on C2-shaped synthetic data.
// Not part of the library; the winning variant is what csrc/gram.cu implements.
//   build: make -C tools gram_lab ; run on the GPU box: tools/gram_lab [cells] [n_kp]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../cvx_proj_b200/csrc/common.cuh"

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
using namespace apap;

constexpr int kRowF = 28;

// ---------------------------------------------------------------------------------------------
// Variant A: one warp per CTA, R cells per thread, optional software-pipelined weights.
template <int R, int ROWS, bool PIPE, int ORDER = 0>
__global__ void __launch_bounds__(32) gram_a(const float *__restrict__ kp_table, const float *__restrict__ anchors,
                                             int cells, int cells_padded, int n_chunks, int chunks_per_split,
                                             float gamma_sq, float *__restrict__ partials) {
  constexpr uint32_t kBytes = ROWS * kRowF * 4;
  __shared__ __align__(128) float stage[2][ROWS * kRowF];
  __shared__ __align__(8) uint64_t full_bar[2];
  const int lane = threadIdx.x;
  const int split = blockIdx.y;
  const int c_begin = split * chunks_per_split;
  const int c_end = min(n_chunks, c_begin + chunks_per_split);
  const int n_local = (c_end - c_begin) * (128 / ROWS);
  const int cell0 = blockIdx.x * (32 * R) + lane;
  float ax[R], ay[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float2 v = reinterpret_cast<const float2 *>(anchors)[min(cell0 + r * 32, cells - 1)];
    ax[r] = v.x; ay[r] = v.y;
  }
  float2 acc[R][12];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int t = 0; t < 12; ++t) acc[r][t] = make_float2(0.f, 0.f);
  if (lane == 0) { mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1); mbar_fence_init(); }
  __syncwarp();
  const char *gsrc = reinterpret_cast<const char *>(kp_table) + (size_t)c_begin * (128 * kRowF * 4);
  if (lane == 0 && n_local > 0) { mbar_arrive_expect_tx(&full_bar[0], kBytes); bulk_g2s(stage[0], gsrc, kBytes, &full_bar[0]); }
  for (int lc = 0; lc < n_local; ++lc) {
    const int s = lc & 1;
    if (lane == 0 && lc + 1 < n_local) {
      mbar_arrive_expect_tx(&full_bar[s ^ 1], kBytes);
      bulk_g2s(stage[s ^ 1], gsrc + (size_t)(lc + 1) * kBytes, kBytes, &full_bar[s ^ 1]);
    }
    mbar_wait(&full_bar[s], (lc >> 1) & 1);
    const float4 *rows = reinterpret_cast<const float4 *>(stage[s]);
    auto weights = [&](const float4 q, float (&w2)[R]) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float dx = ax[r] - q.x, dy = ay[r] - q.z;
        w2[r] = fmaxf(ex2_approx(-sqrt_approx(fmaf(dy, dy, dx * dx))), gamma_sq);
      }
    };
    float w2c[R];
    if (PIPE) weights(rows[6], w2c);
#pragma unroll 2
    for (int i = 0; i < ROWS; ++i) {
      const float4 *row = rows + i * 7;
      float w2[R];
      if (PIPE) {
#pragma unroll
        for (int r = 0; r < R; ++r) w2[r] = w2c[r];
        weights(rows[min(i + 1, ROWS - 1) * 7 + 6], w2c);
      } else {
        weights(row[6], w2);
      }
      if (ORDER == 0) {
#pragma unroll
        for (int v = 0; v < 6; ++v) {
          const float4 p = row[v];
          const float2 plo = make_float2(p.x, p.y), phi = make_float2(p.z, p.w);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float2 ww = make_float2(w2[r], w2[r]);
            acc[r][2 * v + 0] = __ffma2_rn(ww, plo, acc[r][2 * v + 0]);
            acc[r][2 * v + 1] = __ffma2_rn(ww, phi, acc[r][2 * v + 1]);
          }
        }
      } else {
        float4 p[6];
#pragma unroll
        for (int v = 0; v < 6; ++v) p[v] = row[v];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 ww = make_float2(w2[r], w2[r]);
#pragma unroll
          for (int v = 0; v < 6; ++v) {
            acc[r][2 * v + 0] = __ffma2_rn(ww, make_float2(p[v].x, p[v].y), acc[r][2 * v + 0]);
            acc[r][2 * v + 1] = __ffma2_rn(ww, make_float2(p[v].z, p[v].w), acc[r][2 * v + 1]);
          }
        }
      }
    }
    __syncwarp();
  }
  float *dst = partials + (size_t)split * 24 * cells_padded;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = cell0 + r * 32;
    if (c < cells) {
#pragma unroll
      for (int t = 0; t < 12; ++t) {
        dst[(size_t)(2 * t + 0) * cells_padded + c] = acc[r][t].x;
        dst[(size_t)(2 * t + 1) * cells_padded + c] = acc[r][t].y;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Variant B: 4 warps per CTA (one per SM sub-partition) sharing one TMA ring of NST stages; each
// warp owns a contiguous run of 32-cell groups (3 or 4 of them) so the grid can be cut into exactly
// 3 CTAs per SM with near-equal work per sub-partition.
template <int NST, int ROWS>
struct RingB {
  float stage[NST][ROWS * kRowF];
  uint64_t full_bar[NST];
  uint64_t empty_bar[NST];
};

template <int R, int ROWS>
__device__ __forceinline__ void gram_b_body(const float4 *rows, const float (&ax)[4], const float (&ay)[4], float gamma_sq,
                                            float2 (&acc)[4][12]) {
  auto weights = [&](const float4 q, float (&w2)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float dx = ax[r] - q.x, dy = ay[r] - q.z;
      w2[r] = fmaxf(ex2_approx(-sqrt_approx(fmaf(dy, dy, dx * dx))), gamma_sq);
    }
  };
  float w2c[R];
  weights(rows[6], w2c);
#pragma unroll 2
  for (int i = 0; i < ROWS; ++i) {
    const float4 *row = rows + i * 7;
    float w2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) w2[r] = w2c[r];
    weights(rows[min(i + 1, ROWS - 1) * 7 + 6], w2c);
#pragma unroll
    for (int v = 0; v < 6; ++v) {
      const float4 p = row[v];
      const float2 plo = make_float2(p.x, p.y), phi = make_float2(p.z, p.w);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float2 ww = make_float2(w2[r], w2[r]);
        acc[r][2 * v + 0] = __ffma2_rn(ww, plo, acc[r][2 * v + 0]);
        acc[r][2 * v + 1] = __ffma2_rn(ww, phi, acc[r][2 * v + 1]);
      }
    }
  }
}

template <int NST, int ROWS>
__global__ void __launch_bounds__(128) gram_b(const float *__restrict__ kp_table, const float *__restrict__ anchors,
                                              int cells, int cells_padded, int n_chunks, int chunks_per_split,
                                              int ctas_per_split, int n_groups, float gamma_sq,
                                              float *__restrict__ partials) {
  constexpr uint32_t kBytes = ROWS * kRowF * 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RingB<NST, ROWS> &ring = *reinterpret_cast<RingB<NST, ROWS> *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int split = blockIdx.x / ctas_per_split;
  const int cta = blockIdx.x - split * ctas_per_split;
  const int c_begin = split * chunks_per_split;
  const int c_end = min(n_chunks, c_begin + chunks_per_split);
  const int n_local = (c_end - c_begin) * (128 / ROWS);
  // groups of this CTA, then of this warp (rotated so the bigger shares move around the sub-partitions)
  const int g0 = (int)((long long)cta * n_groups / ctas_per_split);
  const int g1 = (int)((long long)(cta + 1) * n_groups / ctas_per_split);
  const int ng = g1 - g0;
  const int slot = (warp + cta) & 3;
  const int w0 = g0 + slot * ng / 4, w1 = g0 + (slot + 1) * ng / 4;
  const int my = w1 - w0;                                   // 0..4 groups
  float ax[4], ay[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = min((w0 + min(r, max(my - 1, 0))) * 32 + lane, cells - 1);
    const float2 v = reinterpret_cast<const float2 *>(anchors)[c];
    ax[r] = v.x; ay[r] = v.y;
  }
  float2 acc[4][12];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int t = 0; t < 12; ++t) acc[r][t] = make_float2(0.f, 0.f);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) { mbar_init(&ring.full_bar[s], 1); mbar_init(&ring.empty_bar[s], 4); }
    mbar_fence_init();
  }
  __syncthreads();
  const char *gsrc = reinterpret_cast<const char *>(kp_table) + (size_t)c_begin * (128 * kRowF * 4);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST && s < n_local; ++s) {
      mbar_arrive_expect_tx(&ring.full_bar[s], kBytes);
      bulk_g2s(ring.stage[s], gsrc + (size_t)s * kBytes, kBytes, &ring.full_bar[s]);
    }
  }
  for (int lc = 0; lc < n_local; ++lc) {
    const int s = lc % NST;
    const uint32_t ph = (lc / NST) & 1;
    mbar_wait(&ring.full_bar[s], ph);
    const float4 *rows = reinterpret_cast<const float4 *>(ring.stage[s]);
    if (my == 4) gram_b_body<4, ROWS>(rows, ax, ay, gamma_sq, acc);
    else if (my == 3) gram_b_body<3, ROWS>(rows, ax, ay, gamma_sq, acc);
    else if (my == 2) gram_b_body<2, ROWS>(rows, ax, ay, gamma_sq, acc);
    else if (my == 1) gram_b_body<1, ROWS>(rows, ax, ay, gamma_sq, acc);
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ring.empty_bar[s])) : "memory");
    // producer: refill stage s with chunk lc + NST once all four warps have released it
    if (threadIdx.x == 0 && lc + NST < n_local) {
      mbar_wait(&ring.empty_bar[s], ph);
      mbar_arrive_expect_tx(&ring.full_bar[s], kBytes);
      bulk_g2s(ring.stage[s], gsrc + (size_t)(lc + NST) * kBytes, kBytes, &ring.full_bar[s]);
    }
  }
  float *dst = partials + (size_t)split * 24 * cells_padded;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int c = (w0 + r) * 32 + lane;
    if (r < my && c < cells) {
#pragma unroll
      for (int t = 0; t < 12; ++t) {
        dst[(size_t)(2 * t + 0) * cells_padded + c] = acc[r][t].x;
        dst[(size_t)(2 * t + 1) * cells_padded + c] = acc[r][t].y;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
static float time_it(void (*launch)(), int reps = 5) {
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  launch(); CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CHECK(cudaEventRecord(e0)); launch(); CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    best = fminf(best, ms);
  }
  return best;
}

static float *d_table, *d_anchors, *d_partials;
static int g_cells, g_cells_padded, g_n_chunks, g_cps, g_splits;
static std::vector<double> g_ref;

static double checksum() {
  std::vector<float> h((size_t)g_splits * 24 * g_cells_padded);
  CHECK(cudaMemcpy(h.data(), d_partials, h.size() * 4, cudaMemcpyDeviceToHost));
  double s = 0;
  for (int sp = 0; sp < g_splits; ++sp)
    for (int t = 0; t < 24; ++t)
      for (int c = 0; c < g_cells; ++c) s += (double)h[((size_t)sp * 24 + t) * g_cells_padded + c] * (1 + (t % 5));
  return s;
}

template <int R, int ROWS, bool PIPE, int ORDER = 0>
static void launch_a() {
  dim3 grid((g_cells + 32 * R - 1) / (32 * R), g_splits);
  gram_a<R, ROWS, PIPE, ORDER><<<grid, 32>>>(d_table, d_anchors, g_cells, g_cells_padded, g_n_chunks, g_cps, 0.25f, d_partials);
}
static int g_ctas_per_split;
template <int NST, int ROWS>
static void launch_b() {
  const int n_groups = (g_cells + 31) / 32;
  const size_t smem = sizeof(RingB<NST, ROWS>);
  static bool once = false;
  if (!once) { CHECK(cudaFuncSetAttribute(gram_b<NST, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); once = true; }
  gram_b<NST, ROWS><<<g_ctas_per_split * g_splits, 128, smem>>>(d_table, d_anchors, g_cells, g_cells_padded, g_n_chunks,
                                                                g_cps, g_ctas_per_split, n_groups, 0.25f, d_partials);
}

static void report(const char *name, float ms) {
  CHECK(cudaGetLastError());
  const double flops = 2.0 * 24 * (double)g_n_chunks * 128 * g_cells;
  printf("%-44s %8.3f ms  %6.2f TFLOP/s  checksum %.9e\n", name, ms, flops / (ms * 1e-3) / 1e12, checksum());
}

int main(int argc, char **argv) {
  g_cells = argc > 1 ? atoi(argv[1]) : 40000;
  const int n_kp = argc > 2 ? atoi(argv[2]) : 5000;
  const int n_pad = (n_kp + 127) / 128 * 128;
  g_n_chunks = n_pad / 128;
  g_cps = (g_n_chunks + 3) / 4; if (g_cps > 8) g_cps = 8;
  g_splits = (g_n_chunks + g_cps - 1) / g_cps;
  g_cells_padded = (g_cells + 511) / 512 * 512;
  printf("cells %d  n_kp %d (padded %d)  chunks %d  splits %d\n", g_cells, n_kp, n_pad, g_n_chunks, g_splits);
  std::vector<float> table((size_t)n_pad * kRowF, 0.f), anchors((size_t)g_cells * 2);
  srand(1);
  const float s = 2.0f * 1.4426950f / 1e4f;
  for (int i = 0; i < n_kp; ++i) {
    for (int t = 0; t < 24; ++t) table[(size_t)i * kRowF + t] = (rand() % 2001 - 1000) * 1e-3f;
    const float kx = (rand() % 3840) * s, ky = (rand() % 2160) * s;
    table[(size_t)i * kRowF + 24] = kx; table[(size_t)i * kRowF + 25] = kx;
    table[(size_t)i * kRowF + 26] = ky; table[(size_t)i * kRowF + 27] = ky;
  }
  const int side = (int)ceil(sqrt((double)g_cells));
  for (int c = 0; c < g_cells; ++c) { anchors[2 * c] = (c % side) * 4448.f / side * s; anchors[2 * c + 1] = (c / side) * 2332.f / side * s; }
  CHECK(cudaMalloc(&d_table, table.size() * 4)); CHECK(cudaMalloc(&d_anchors, anchors.size() * 4));
  CHECK(cudaMalloc(&d_partials, (size_t)g_splits * 24 * g_cells_padded * 4));
  CHECK(cudaMemcpy(d_table, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(d_anchors, anchors.data(), anchors.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemset(d_partials, 0, (size_t)g_splits * 24 * g_cells_padded * 4));

  report("A R=4 rows=64 plain", time_it(launch_a<4, 64, false>));
  report("A R=4 rows=64 pipelined weights", time_it(launch_a<4, 64, true>));
  report("A R=4 rows=64 pipelined, cell-outer", time_it(launch_a<4, 64, true, 1>));
  report("A R=4 rows=64 plain, cell-outer", time_it(launch_a<4, 64, false, 1>));
  report("A R=3 rows=64 pipelined", time_it(launch_a<3, 64, true>));
  report("A R=2 rows=32 plain", time_it(launch_a<2, 32, false>));
  report("A R=2 rows=32 pipelined", time_it(launch_a<2, 32, true>));
  int sms = 0; CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int per_sm = 2; per_sm <= 4; ++per_sm) {
    g_ctas_per_split = sms * per_sm / g_splits;
    char name[96];
    snprintf(name, sizeof name, "B 4-warp CTA, %d CTA/SM (%d per split), 3 st x 64", per_sm, g_ctas_per_split);
    CHECK(cudaMemset(d_partials, 0, (size_t)g_splits * 24 * g_cells_padded * 4));
    report(name, time_it(launch_b<3, 64>));
  }
  return 0;
}
