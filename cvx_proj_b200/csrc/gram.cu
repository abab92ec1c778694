// K1 -- moving-DLT weights + Gram contraction (reference pyviz/apap.py:150-152,159).
//
// For every cell c and keypoint i:   w2 = max(exp(-2 |v_c - x_i| / sigma^2), gamma^2)
//                                    S_t(c) += w2 * P[i][t],  t = 0..23
// i.e. the contraction [cells x N of w^2] . [N x 24] with the weights generated on the fly and
// never written to memory.  FP32 SIMT: one thread owns R cells (register tile R x 24
// accumulators), the keypoint rows (24 terms + kx, ky) are staged in shared memory by TMA bulk
// copies (cp.async.bulk + mbarrier, double buffered) and read back as warp-wide broadcast
// LDS.128, so one 7-instruction row fetch feeds R x 24 FFMAs.
//
// Accuracy: each FP32 accumulator sums at most kMaxChainChunks*128 = 1024 keypoints (one
// "split"); the splits are written as partial sums and combined in float64 by K2.  The split
// boundaries depend only on the keypoint count, so a cell's result does not depend on how the
// grid is sharded across GPUs.
#include "common.cuh"

namespace apap {

constexpr int kGramThreads = 128;
constexpr int kStages = 2;
constexpr uint32_t kStageBytes = kChunk * kRowFloats * sizeof(float);   // 14336

static int g_sm_count = 0;
int sm_count_cached() {
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_sm_count = n;
    else
      return 148;   // B200; only reached without a device (planning on the host)
  }
  return g_sm_count;
}

GramPlan make_gram_plan(int cells, int n_kp_padded, int sm_count) {
  GramPlan p;
  const int n_chunks = n_kp_padded / kChunk;
  int cps = (n_chunks + 3) / 4;                 // a function of N only (see header comment)
  if (cps < 1) cps = 1;
  if (cps > kMaxChainChunks) cps = kMaxChainChunks;
  p.chunks_per_split = cps;
  p.k_splits = (n_chunks + cps - 1) / cps;
  if (p.k_splits < 1) p.k_splits = 1;
  // register tile: the largest R that still gives every SM two CTAs' worth of work
  int r = 4;
  while (r > 1) {
    long tiles = (cells + kGramThreads * r - 1) / (kGramThreads * r);
    if (tiles * p.k_splits >= 2L * sm_count) break;
    r >>= 1;
  }
  p.cells_per_thread = r;
  // the partial buffer is padded for the widest tile so its size does not depend on R
  const int pad = kGramThreads * 4;
  p.cells_padded = (cells + pad - 1) / pad * pad;
  p.cell_tiles = (cells + kGramThreads * r - 1) / (kGramThreads * r);
  return p;
}

template <int R>
__global__ void __launch_bounds__(kGramThreads) k_gram(const float *__restrict__ kp_table,
                                                        const float *__restrict__ anchors, int cells,
                                                        int cells_padded, int n_chunks,
                                                        int chunks_per_split, int k_splits, float k2,
                                                        float gamma_sq, float *__restrict__ partials) {
  __shared__ __align__(128) float stage[kStages][kChunk * kRowFloats];
  __shared__ __align__(8) uint64_t full_bar[kStages];

  const int tid = threadIdx.x;
  const int split = blockIdx.y;
  const int scene = blockIdx.z;
  const int c_begin = split * chunks_per_split;
  const int c_end = min(n_chunks, c_begin + chunks_per_split);
  const int n_local = c_end - c_begin;

  kp_table += (size_t)scene * n_chunks * (kChunk * kRowFloats);
  anchors += (size_t)scene * cells * 2;
  partials += (size_t)scene * k_splits * kTerms * cells_padded;

  const int cell0 = blockIdx.x * (kGramThreads * R) + tid;
  float vx[R], vy[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = min(cell0 + r * kGramThreads, cells - 1);
    const float2 v = reinterpret_cast<const float2 *>(anchors)[c];
    vx[r] = v.x;
    vy[r] = v.y;
  }

  float acc[R][kTerms];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int t = 0; t < kTerms; ++t) acc[r][t] = 0.f;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const char *gsrc = reinterpret_cast<const char *>(kp_table) + (size_t)c_begin * kStageBytes;
  if (tid == 0 && n_local > 0) {
    mbar_arrive_expect_tx(&full_bar[0], kStageBytes);
    bulk_g2s(stage[0], gsrc, kStageBytes, &full_bar[0]);
  }

  for (int lc = 0; lc < n_local; ++lc) {
    const int s = lc & 1;
    // prefetch the next chunk into the other stage (freed by the barrier that ended lc-1)
    if (tid == 0 && lc + 1 < n_local) {
      mbar_arrive_expect_tx(&full_bar[s ^ 1], kStageBytes);
      bulk_g2s(stage[s ^ 1], gsrc + (size_t)(lc + 1) * kStageBytes, kStageBytes, &full_bar[s ^ 1]);
    }
    mbar_wait(&full_bar[s], (lc >> 1) & 1);

    const float4 *rows = reinterpret_cast<const float4 *>(stage[s]);
#pragma unroll 2
    for (int i = 0; i < kChunk; ++i) {
      const float4 *row = rows + i * 7;
      const float4 q = row[6];            // kx, ky, pad, pad
      float w2[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float dx = vx[r] - q.x;
        const float dy = vy[r] - q.y;
        const float d2 = fmaf(dy, dy, dx * dx);
        const float e = ex2_approx(sqrt_approx(d2) * k2);
        w2[r] = fmaxf(e, gamma_sq);
      }
#pragma unroll
      for (int v = 0; v < 6; ++v) {
        const float4 p = row[v];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[r][4 * v + 0] = fmaf(w2[r], p.x, acc[r][4 * v + 0]);
          acc[r][4 * v + 1] = fmaf(w2[r], p.y, acc[r][4 * v + 1]);
          acc[r][4 * v + 2] = fmaf(w2[r], p.z, acc[r][4 * v + 2]);
          acc[r][4 * v + 3] = fmaf(w2[r], p.w, acc[r][4 * v + 3]);
        }
      }
    }
    __syncthreads();   // every warp is done with stage s before it is refilled
  }

  // partials[split][t][cell]: consecutive threads write consecutive cells (coalesced)
  float *dst = partials + (size_t)split * kTerms * cells_padded;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int c = cell0 + r * kGramThreads;
    if (c < cells) {
#pragma unroll
      for (int t = 0; t < kTerms; ++t) dst[(size_t)t * cells_padded + c] = acc[r][t];
    }
  }
}

int launch_gram(const float *kp_table, const float *anchors, int batch, int cells, int n_kp_padded, float k2,
                float gamma_sq, float *partials, cudaStream_t st) {
  const GramPlan p = make_gram_plan(cells, n_kp_padded, sm_count_cached());
  const int n_chunks = n_kp_padded / kChunk;
  dim3 grid(p.cell_tiles, p.k_splits, batch);
  if (p.k_splits > 65535 || batch > 65535) return fail(APAP_E_TOOBIG, "gram: grid.y/z exceeds 65535");
  switch (p.cells_per_thread) {
    case 4:
      k_gram<4><<<grid, kGramThreads, 0, st>>>(kp_table, anchors, cells, p.cells_padded, n_chunks,
                                               p.chunks_per_split, p.k_splits, k2, gamma_sq, partials);
      break;
    case 2:
      k_gram<2><<<grid, kGramThreads, 0, st>>>(kp_table, anchors, cells, p.cells_padded, n_chunks,
                                               p.chunks_per_split, p.k_splits, k2, gamma_sq, partials);
      break;
    default:
      k_gram<1><<<grid, kGramThreads, 0, st>>>(kp_table, anchors, cells, p.cells_padded, n_chunks,
                                               p.chunks_per_split, p.k_splits, k2, gamma_sq, partials);
      break;
  }
  return check_cuda(cudaGetLastError(), "k_gram launch");
}

}  // namespace apap
