// K2 -- batched 9x9 symmetric eigensolve + de-normalisation
// (reference pyviz/apap.py:160-161 cv.SVDecomp + V[-1], and :164-168).
//
// One thread per cell: the 45 upper-triangle entries of the Gram matrix and the 81 entries of
// the eigenvector matrix live in registers (every index is a compile-time constant after
// unrolling), so a warp solves 32 cells with no shuffles or shared memory.  The k_splits FP32
// partial sums of K1 are combined in float64 in a fixed order, scaled by 1/sum(w^2) and
// diagonalised by cyclic Jacobi in FP32 with the relative rotation threshold
// |g_pq| <= eps * sqrt(g_pp g_qq) (so small eigenvalues keep their relative accuracy).  The
// eigenvector of the smallest eigenvalue is de-normalised in float64 and stored as float32.
#include "common.cuh"

namespace apap {

constexpr int kEigThreads = 128;
constexpr int kMaxSweeps = 12;

__host__ __device__ constexpr int tri(int i, int j) {   // index into the packed upper triangle
  return i <= j ? (i * (19 - i)) / 2 + (j - i) : (j * (19 - j)) / 2 + (i - j);
}
// index of (a, b), a,b in 0..2, into the packed symmetric 3x3 block [xx xy x yy y 1]
__host__ __device__ constexpr int sym3(int a, int b) {
  return a <= b ? (a == 0 ? b : (a == 1 ? 2 + b : 5)) : sym3(b, a);
}

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

__global__ void __launch_bounds__(kEigThreads) k_eig(const float *__restrict__ partials,
                                                      const double *__restrict__ tmats, int cells,
                                                      int cells_padded, int k_splits,
                                                      float *__restrict__ out_h, int *__restrict__ out_sweeps) {
  const int cell = blockIdx.x * kEigThreads + threadIdx.x;
  const int scene = blockIdx.y;
  if (cell >= cells) return;
  partials += (size_t)scene * k_splits * kTerms * cells_padded;
  tmats += (size_t)scene * 18;

  // ---- combine the split partials in float64 (fixed order -> deterministic) -----------------
  double sum[kTerms];
#pragma unroll
  for (int t = 0; t < kTerms; ++t) sum[t] = 0.0;
  for (int s = 0; s < k_splits; ++s) {
    const float *src = partials + (size_t)s * kTerms * cells_padded + cell;
#pragma unroll
    for (int t = 0; t < kTerms; ++t) sum[t] += (double)__ldg(src + (size_t)t * cells_padded);
  }
  const double inv_w = 1.0 / sum[5];   // sum of w^2 > 0 (gamma > 0 or any finite weight)

  // ---- expand to the packed 9x9:  [[S,0,-Sx],[0,S,-Sy],[-Sx,-Sy,Sr]] ------------------------
  float g[45];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int k = sym3(a, b);
      if (a <= b) {
        g[tri(a, b)] = (float)(sum[k] * inv_w);
        g[tri(3 + a, 3 + b)] = (float)(sum[k] * inv_w);
        g[tri(6 + a, 6 + b)] = (float)(sum[18 + k] * inv_w);
      }
      g[tri(a, 3 + b)] = 0.f;
      g[tri(a, 6 + b)] = (float)(-sum[6 + k] * inv_w);
      g[tri(3 + a, 6 + b)] = (float)(-sum[12 + k] * inv_w);
    }
  }
  float v[81];
#pragma unroll
  for (int i = 0; i < 81; ++i) v[i] = (i / 9 == i % 9) ? 1.f : 0.f;

  // ---- cyclic Jacobi ------------------------------------------------------------------------
  int sweeps = 0;
  for (; sweeps < kMaxSweeps; ++sweeps) {
    if (!Sweep<0, 1>::run(g, v)) break;
  }

  // ---- eigenvector of the smallest eigenvalue -------------------------------------------------
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

  // ---- H = T2inv * reshape(h, 3, 3) * T1, divided by H[2][2]  (float64, stored float32) -------
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
  float *dst = out_h + ((size_t)scene * cells + cell) * 9;
#pragma unroll
  for (int i = 0; i < 9; ++i) dst[i] = (float)(o[i] * inv22);
  if (out_sweeps) out_sweeps[(size_t)scene * cells + cell] = sweeps;
}

int launch_eig(const float *partials, const double *tmats, int batch, int cells, int n_kp_padded, float *out_h,
               int *out_sweeps, cudaStream_t st) {
  const GramPlan p = make_gram_plan(cells, n_kp_padded, sm_count_cached());
  dim3 grid((cells + kEigThreads - 1) / kEigThreads, batch);
  if (batch > 65535) return fail(APAP_E_TOOBIG, "eig: batch exceeds 65535");
  k_eig<<<grid, kEigThreads, 0, st>>>(partials, tmats, cells, p.cells_padded, p.k_splits, out_h, out_sweeps);
  return check_cuda(cudaGetLastError(), "k_eig launch");
}

}  // namespace apap
