// Spectral match weighting (second half) -- the reference's calculate_M,
// pyviz/spectral_method.py:96-125: the N x N affinity matrix over the coarse matches and its leading
// singular vector.
//   k_affinity   M[i][j] = max(4.5 - ((|s_i - s_j|^2 - |d_i - d_j|^2)^2) rcp, 0) for i != j in float32 exactly as numpy
//                evaluates :112-118 (separate multiplies and adds, no FMA), widened to float64; M[i][i] = diag[i]
//                (descriptor similarity + epipolar term, O(N D) on the host, :104-111)
//   k_matvec     y = M x in float64, one warp per row, and |y|^2 accumulated for the normalisation
//   k_rescale    x <- y / |y|, max |x_new - x_old| for the convergence test
// The reference takes |U[:, 0]| of a full SVD (np.linalg.svd, O(N^3)); M is symmetric with non-negative entries and a
// positive diagonal, so that vector is its Perron vector and power iteration from a positive start converges to it.
// HBM-bound: a matvec streams the 8 N^2 bytes of M once.
#include "common.cuh"

namespace apap {

__global__ void __launch_bounds__(256) k_affinity(const float2 *__restrict__ src, const float2 *__restrict__ dst,
                                                   const double *__restrict__ diag, int n, float rcp_value,
                                                   double *__restrict__ m) {
  const int j = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
  if (j >= n) return;
  double v;
  if (i == j) {
    v = diag[i];
  } else {
    const float2 si = src[i], sj = src[j], di = dst[i], dj = dst[j];
    const float sx = __fsub_rn(si.x, sj.x), sy = __fsub_rn(si.y, sj.y);
    const float dx = __fsub_rn(di.x, dj.x), dy = __fsub_rn(di.y, dj.y);
    const float sm = __fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy));
    const float dm = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float e = __fsub_rn(sm, dm);
    const float dist = __fmul_rn(__fmul_rn(e, e), rcp_value);
    v = (double)fmaxf(__fsub_rn(4.5f, dist), 0.f);
  }
  m[(size_t)i * n + j] = v;
}

constexpr int kMvWarps = 8;
__global__ void __launch_bounds__(kMvWarps * 32) k_matvec(const double *__restrict__ m, const double *__restrict__ x,
                                                           int n, double *__restrict__ y, double *__restrict__ norm_sq) {
  const int row = blockIdx.x * kMvWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const double *r = m + (size_t)row * n;
  double acc0 = 0.0, acc1 = 0.0;
  int j = lane;
  for (; j + 32 < n; j += 64) {
    acc0 = fma(__ldg(r + j), x[j], acc0);
    acc1 = fma(__ldg(r + j + 32), x[j + 32], acc1);
  }
  if (j < n) acc0 = fma(__ldg(r + j), x[j], acc0);
  double acc = acc0 + acc1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    y[row] = acc;
    atomicAdd(norm_sq, acc * acc);
  }
}

__global__ void __launch_bounds__(256) k_rescale(const double *__restrict__ y, const double *__restrict__ norm_sq, int n,
                                                  double *__restrict__ x, unsigned long long *__restrict__ max_diff_bits) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  double d = 0.0;
  if (i < n) {
    const double v = y[i] / sqrt(*norm_sq);
    d = fabs(v - x[i]);
    x[i] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
  if ((threadIdx.x & 31) == 0 && d > 0.0) atomicMax(max_diff_bits, (unsigned long long)__double_as_longlong(d));
}

}  // namespace apap

using namespace apap;

extern "C" {

int apap_affinity_matrix(const float *src_pts, const float *dst_pts, const double *diag, int n, float rcp_value, double *m,
                         void *stream) {
  if (!src_pts || !dst_pts || !diag || !m || n <= 0) return fail(APAP_E_BADARG, "affinity_matrix: bad arguments");
  if (n > 65535) return fail(APAP_E_TOOBIG, "affinity_matrix: more than 65535 matches");
  if ((reinterpret_cast<uintptr_t>(src_pts) | reinterpret_cast<uintptr_t>(dst_pts)) & 7u)
    return fail(APAP_E_ALIGN, "affinity_matrix: point arrays must be 8-byte aligned");
  k_affinity<<<dim3((n + 255) / 256, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2 *>(src_pts), reinterpret_cast<const float2 *>(dst_pts), diag, n, rcp_value, m);
  return check_cuda(cudaGetLastError(), "k_affinity launch");
}

int apap_power_step(const double *m, int n, double *x, double *y, double *norm_sq, unsigned long long *max_diff_bits,
                    void *stream) {
  if (!m || !x || !y || !norm_sq || !max_diff_bits || n <= 0) return fail(APAP_E_BADARG, "power_step: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cuda(cudaMemsetAsync(norm_sq, 0, sizeof(double), st), "power_step: memset");
  if (rc) return rc;
  k_matvec<<<(n + kMvWarps - 1) / kMvWarps, kMvWarps * 32, 0, st>>>(m, x, n, y, norm_sq);
  rc = check_cuda(cudaGetLastError(), "k_matvec launch");
  if (rc) return rc;
  k_rescale<<<(n + 255) / 256, 256, 0, st>>>(y, norm_sq, n, x, max_diff_bits);
  return check_cuda(cudaGetLastError(), "k_rescale launch");
}

}  // extern "C"
