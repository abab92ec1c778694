// K2 -- batched 9x9 symmetric eigensolve + de-normalisation
// (reference pyviz/apap.py:160-161 cv.SVDecomp + V[-1], and :164-168).
//
// One thread per cell, everything in registers (every index is a compile-time constant after
// unrolling).  The k_splits FP32 partial sums of K1 are combined in float64 in a fixed order and
// scaled by 1/sum(w^2).  The wanted vector -- the right singular vector of the weighted DLT
// matrix for its smallest singular value -- is the eigenvector of the smallest eigenvalue of the
// 9x9 Gram matrix.  That eigenvalue is separated from the next one by 4-6 orders of magnitude
// on real data, so the fast path is float64 inverse iteration on an LDL^T
// factorisation (~0.5 kflop per cell, settles in 2-3 steps).  A cell whose iteration has not
// settled after kMaxInvIter steps (tiny spectral gap: degenerate keypoint sets) falls back to
// the full cyclic Jacobi diagonalisation in FP32 with the relative rotation threshold
// |g_pq| <= eps * sqrt(g_pp g_qq).  The eigenvector is de-normalised in float64 and stored as
// float32.
#include "common.cuh"

namespace apap {

constexpr int kEigThreads = 128;
constexpr int kMaxSweeps = 12;
constexpr int kMaxInvIter = 8;

__host__ __device__ constexpr int tri(int i, int j) {   // index into the packed upper triangle
  return i <= j ? (i * (19 - i)) / 2 + (j - i) : (j * (19 - j)) / 2 + (i - j);
}
// index of (a, b), a,b in 0..2, into the packed symmetric 3x3 block [xx xy x yy y 1]
__host__ __device__ constexpr int sym3(int a, int b) {
  return a <= b ? (a == 0 ? b : (a == 1 ? 2 + b : 5)) : sym3(b, a);
}

// ------------------------------------------------------------------------- shared pieces
// Combine the k_splits FP32 partial sums of one cell in float64 (fixed order -> deterministic).
__device__ __forceinline__ void combine_partials(const float *__restrict__ partials, int cells_padded, int k_splits,
                                                 int cell, double (&sum)[kTerms]) {
#pragma unroll
  for (int t = 0; t < kTerms; ++t) sum[t] = 0.0;
  for (int s = 0; s < k_splits; ++s) {
    const float *src = partials + (size_t)s * kTerms * cells_padded + cell;
#pragma unroll
    for (int t = 0; t < kTerms; ++t) sum[t] += (double)__ldcg(src + (size_t)t * cells_padded);   // L2: K1 may still be running
  }
}

// Packed 9x9 Gram matrix [[S,0,-Sx],[0,S,-Sy],[-Sx,-Sy,Sr]] / sum(w^2) from the 24 sums.
template <typename T>
__device__ __forceinline__ void expand_gram(const double (&sum)[kTerms], T (&g)[45]) {
  const double inv_w = 1.0 / sum[5];   // sum of w^2 > 0 (gamma > 0 or any finite weight)
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int k = sym3(a, b);
      if (a <= b) {
        g[tri(a, b)] = (T)(sum[k] * inv_w);
        g[tri(3 + a, 3 + b)] = (T)(sum[k] * inv_w);
        g[tri(6 + a, 6 + b)] = (T)(sum[18 + k] * inv_w);
      }
      g[tri(a, 3 + b)] = (T)0;
      g[tri(a, 6 + b)] = (T)(-sum[6 + k] * inv_w);
      g[tri(3 + a, 6 + b)] = (T)(-sum[12 + k] * inv_w);
    }
  }
}

// H = T2inv * reshape(h, 3, 3) * T1, divided by H[2][2]  (float64, stored float32; pyviz/apap.py:164-168)
__device__ __forceinline__ void denorm_store(const double (&h)[9], const double *__restrict__ tmats,
                                             float *__restrict__ dst) {
  double t2[9], t1[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    t2[i] = tmats[i];
    t1[i] = tmats[9 + i];
  }
  double m[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      m[r * 3 + c] = t2[r * 3 + 0] * h[0 * 3 + c] + t2[r * 3 + 1] * h[1 * 3 + c] + t2[r * 3 + 2] * h[2 * 3 + c];
  double o[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      o[r * 3 + c] = m[r * 3 + 0] * t1[0 * 3 + c] + m[r * 3 + 1] * t1[1 * 3 + c] + m[r * 3 + 2] * t1[2 * 3 + c];
  const double inv22 = 1.0 / o[8];
#pragma unroll
  for (int i = 0; i < 9; ++i) dst[i] = (float)(o[i] * inv22);
}

// ------------------------------------------------------------------ fallback: cyclic Jacobi
template <int P, int Q>
__device__ __forceinline__ bool rotate(float (&g)[45], float (&v)[81]) {
  const float apq = g[tri(P, Q)];
  const float app = g[tri(P, P)];
  const float aqq = g[tri(Q, Q)];
  const bool rot = fabsf(apq) > 5.9604645e-8f * sqrtf(fabsf(app * aqq));
  const float theta = (aqq - app) / (2.f * apq);
  float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(fmaf(theta, theta, 1.f)));
  float c = 1.f / sqrtf(fmaf(t, t, 1.f));
  float s = t * c;
  if (!rot) {
    t = 0.f;
    c = 1.f;
    s = 0.f;
  }
  g[tri(P, P)] = fmaf(-t, apq, app);
  g[tri(Q, Q)] = fmaf(t, apq, aqq);
  g[tri(P, Q)] = rot ? 0.f : apq;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    if (k != P && k != Q) {
      const float gkp = g[tri(k, P)];
      const float gkq = g[tri(k, Q)];
      g[tri(k, P)] = fmaf(c, gkp, -s * gkq);
      g[tri(k, Q)] = fmaf(s, gkp, c * gkq);
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float vkp = v[k * 9 + P];
    const float vkq = v[k * 9 + Q];
    v[k * 9 + P] = fmaf(c, vkp, -s * vkq);
    v[k * 9 + Q] = fmaf(s, vkp, c * vkq);
  }
  return rot;
}

template <int P, int Q>
struct Sweep {
  __device__ __forceinline__ static bool run(float (&g)[45], float (&v)[81]) {
    const bool a = rotate<P, Q>(g, v);
    const bool b = Sweep<(Q == 8 ? P + 1 : P), (Q == 8 ? P + 2 : Q + 1)>::run(g, v);
    return a || b;
  }
};
template <>
struct Sweep<8, 9> {
  __device__ __forceinline__ static bool run(float (&)[45], float (&)[81]) { return false; }
};

// One cell by full diagonalisation, then de-normalise and store.  Self-contained (re-reads the
// partials) so the fast path of k_eig keeps its matrix in registers.
__device__ __noinline__ void eig_cell_jacobi(const float *__restrict__ partials, const double *__restrict__ tmats,
                                             int cells_padded, int k_splits, int cell, float *__restrict__ dst,
                                             int *__restrict__ sweeps_out) {
  float g[45];
  {
    double sum[kTerms];
    combine_partials(partials, cells_padded, k_splits, cell, sum);
    expand_gram(sum, g);
  }
  float v[81];
#pragma unroll
  for (int i = 0; i < 81; ++i) v[i] = (i / 9 == i % 9) ? 1.f : 0.f;
  int sweeps = 0;
  for (; sweeps < kMaxSweeps; ++sweeps) {
    if (!Sweep<0, 1>::run(g, v)) break;
  }
  int kmin = 0;
  float lmin = g[tri(0, 0)];
#pragma unroll
  for (int k = 1; k < 9; ++k) {
    const float l = g[tri(k, k)];
    if (l < lmin) {
      lmin = l;
      kmin = k;
    }
  }
  double h[9];
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    float x = v[r * 9 + 0];
#pragma unroll
    for (int k = 1; k < 9; ++k) x = (kmin == k) ? v[r * 9 + k] : x;
    h[r] = (double)x;
  }
  denorm_store(h, tmats, dst);
  if (sweeps_out) *sweeps_out = sweeps;
}

// ------------------------------------------------------- fast path: LDL^T inverse iteration
// In-place LDL^T of the packed symmetric matrix: afterwards a[tri(j,j)] = d_j, inv_d[j] = 1/d_j
// and, for i > j, a[tri(j,i)] = l_ij.  Pivots are floored at floor_d so a singular Gram matrix
// (exact-homography data) still yields a usable factorisation for inverse iteration.
__device__ __forceinline__ void ldlt9(double (&a)[45], double (&inv_d)[9], double floor_d) {
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    double t[9];                                     // t_k = l_jk d_k
    double d = a[tri(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) {
      t[k] = a[tri(k, j)] * a[tri(k, k)];
      d = fma(-a[tri(k, j)], t[k], d);
    }
    d = fmax(d, floor_d);
    a[tri(j, j)] = d;
    const double r = 1.0 / d;
    inv_d[j] = r;
#pragma unroll
    for (int i = j + 1; i < 9; ++i) {
      double s = a[tri(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fma(-a[tri(k, i)], t[k], s);
      a[tri(j, i)] = s * r;
    }
  }
}

// x <- (L D L^T)^-1 x
__device__ __forceinline__ void ldlt9_solve(const double (&a)[45], const double (&inv_d)[9], double (&x)[9]) {
#pragma unroll
  for (int i = 1; i < 9; ++i)
#pragma unroll
    for (int k = 0; k < i; ++k) x[i] = fma(-a[tri(k, i)], x[k], x[i]);
#pragma unroll
  for (int i = 0; i < 9; ++i) x[i] *= inv_d[i];
#pragma unroll
  for (int i = 7; i >= 0; --i)
#pragma unroll
    for (int k = i + 1; k < 9; ++k) x[i] = fma(-a[tri(i, k)], x[k], x[i]);
}

// 3 CTAs per SM (168 registers): a CTA fits beside two resident CTAs of K1, so in the overlapped launch K2
// fills the SMs as K1's last CTAs retire.  tile_done (overlapped launch only): K1 counts the finished splits of
// every 128-cell tile there; a CTA = one tile waits for its count, then clears it for the next call.
__global__ void __launch_bounds__(kEigThreads, 3) k_eig(const float *partials, const double *__restrict__ tmats,
                                                         int cells, int cells_padded, int k_splits, int force_jacobi,
                                                         float *__restrict__ out_h, int *__restrict__ out_sweeps,
                                                         int *tile_done) {
  const int cell = blockIdx.x * kEigThreads + threadIdx.x;
  const int scene = blockIdx.y;
  if (tile_done) {
    if (threadIdx.x == 0) {
      int *cnt = tile_done + (size_t)scene * gridDim.x + blockIdx.x;
      int seen;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
        if (seen < k_splits) __nanosleep(200);
      } while (seen < k_splits);
      *cnt = 0;                                        // zero on exit: ready for the next call
    }
    __syncthreads();
  }
  if (cell >= cells) return;
  partials += (size_t)scene * k_splits * kTerms * cells_padded;
  tmats += (size_t)scene * 18;
  float *dst = out_h + ((size_t)scene * cells + cell) * 9;
  int *sw = out_sweeps ? out_sweeps + (size_t)scene * cells + cell : nullptr;

  double f[45];
  {
    double sum[kTerms];
    combine_partials(partials, cells_padded, k_splits, cell, sum);
    expand_gram(sum, f);
  }
  double trace = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) trace += f[tri(i, i)];
  double inv_d[9];
  ldlt9(f, inv_d, trace * 1e-30 + 1e-300);
  // fixed start vector with components of mixed size and sign (not orthogonal to anything special)
  double h[9] = {0.31, -0.17, 0.43, 0.29, 0.37, -0.23, 0.41, 0.19, 0.47};
  int iters = 0;
  bool settled = false;
  double prev_diff = 1.0;                              // the change of the previous step ~ the error before it
  for (; iters < kMaxInvIter && !settled; ++iters) {
    double y[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) y[i] = h[i];
    ldlt9_solve(f, inv_d, y);
    double nrm = 0.0, big = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      nrm = fma(y[i], y[i], nrm);
      big = (fabs(y[i]) > fabs(big)) ? y[i] : big;
    }
    if (!(nrm > 0.0 && nrm < 1e300)) break;            // overflow / NaN (rank-deficient matrix): not settled
    const double sc = copysign(rsqrt(nrm), big);       // unit norm, largest component positive
    double diff = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double yi = y[i] * sc;
      diff = fmax(diff, fabs(yi - h[i]));
      h[i] = yi;
    }
    // Inverse iteration contracts the error by rho = lambda_0 / lambda_1 per step and `diff` is the error of the
    // PREVIOUS iterate, so the new one is within ~ diff * rho: stop when that is below 1e-11 (the output is float32),
    // with rho estimated from the last two changes (never trusted below 1e-3 per step, and not before step 2)
    const double rho = fmin(fmax(diff / prev_diff, 1e-3), 1.0);
    settled = diff <= 1e-10 || (iters >= 1 && diff * rho <= 1e-11);
    prev_diff = fmax(diff, 1e-300);
  }
  if (!settled || force_jacobi) {
    eig_cell_jacobi(partials, tmats, cells_padded, k_splits, cell, dst, sw);
    return;
  }
  denorm_store(h, tmats, dst);
  if (sw) *sw = -iters;                                 // negative: inverse-iteration steps used
}

int launch_eig(const float *partials, const double *tmats, int batch, int cells, int k_splits, float *out_h,
               int *out_sweeps, int force_jacobi, int *tile_done, cudaStream_t st) {
  const int cells_padded = make_gram_plan(cells, kChunk, APAP_GRAM_FFMA2).cells_padded;   // depends on cells only
  dim3 grid((cells + kEigThreads - 1) / kEigThreads, batch);
  if (batch > 65535) return fail(APAP_E_TOOBIG, "eig: batch exceeds 65535");
  if (!tile_done) {
    k_eig<<<grid, kEigThreads, 0, st>>>(partials, tmats, cells, cells_padded, k_splits, force_jacobi, out_h, out_sweeps,
                                        nullptr);
    return check_cuda(cudaGetLastError(), "k_eig launch");
  }
  // overlapped with K1 (the previous kernel in the stream) by programmatic dependent launch: K2 may start once
  // every CTA of K1 is running; it waits per tile on tile_done instead of on the completion of the grid
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kEigThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return check_cuda(cudaLaunchKernelEx(&cfg, k_eig, partials, tmats, cells, cells_padded, k_splits, force_jacobi, out_h,
                                       out_sweeps, tile_done),
                    "k_eig launch (programmatic dependent)");
}

}  // namespace apap
