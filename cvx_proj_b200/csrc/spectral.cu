// Spectral match weighting (second half) -- the reference's calculate_M,
// pyviz/spectral_method.py:96-125: the N x N affinity matrix over the coarse matches and its leading
// singular vector.
//   k_affinity   M[i][j] = max(4.5 - ((|s_i - s_j|^2 - |d_i - d_j|^2)^2) rcp, 0) for i != j in float32 exactly as numpy
//                evaluates :112-118 (separate multiplies and adds, no FMA), widened to float64; M[i][i] = diag[i]
//                (descriptor similarity + epipolar term, O(N D) on the host, :104-111)
//   k_power_step y_next = M (y_prev / |y_prev|) in float64, four warps per row, |y_next|^2 accumulated for the next step
//   k_power_diff x = y / |y| and max |x_new - x_old| for the convergence test (once per group of steps)
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

constexpr int kMvWarps = 8;          // warps per CTA
constexpr int kRowWarps = 4;         // warps sharing one row of M (enough loads in flight to stream M from L2 / HBM)
constexpr int kMvRows = kMvWarps / kRowWarps;
// One power step: y_next = M (y_prev / |y_prev|), |y_next|^2 accumulated into *norm_next.  The normalisation of the
// previous iterate is applied to the finished row sum, so a step is ONE kernel; the three norm accumulators form a
// ring (read k % 3, accumulate into (k + 1) % 3, clear (k + 2) % 3 for the step after).  A row is summed by
// kRowWarps warps over interleaved 32-element segments, four segments in flight per lane, partial sums combined in
// a fixed order.
__global__ void __launch_bounds__(kMvWarps * 32) k_power_step(const double *__restrict__ m, const double *__restrict__ y_prev,
                                                               const double *__restrict__ norm_prev, int n,
                                                               double *__restrict__ y_next, double *__restrict__ norm_next,
                                                               double *__restrict__ norm_clear) {
  __shared__ double part[kMvWarps];
  if (blockIdx.x == 0 && threadIdx.x == 0) *norm_clear = 0.0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kMvRows + warp / kRowWarps;
  double acc = 0.0;
  if (row < n) {
    const double *r = m + (size_t)row * n;
    constexpr int kStride = kRowWarps * 32;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int j = (warp % kRowWarps) * 32 + lane;
    for (; j + 3 * kStride < n; j += 4 * kStride) {
      const double m0 = __ldg(r + j), m1 = __ldg(r + j + kStride), m2 = __ldg(r + j + 2 * kStride),
                   m3 = __ldg(r + j + 3 * kStride);
      acc0 = fma(m0, y_prev[j], acc0);
      acc1 = fma(m1, y_prev[j + kStride], acc1);
      acc2 = fma(m2, y_prev[j + 2 * kStride], acc2);
      acc3 = fma(m3, y_prev[j + 3 * kStride], acc3);
    }
    for (; j < n; j += kStride) acc0 = fma(__ldg(r + j), y_prev[j], acc0);
    acc = (acc0 + acc1) + (acc2 + acc3);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x < kMvRows) {
    const int out_row = blockIdx.x * kMvRows + threadIdx.x;
    if (out_row < n) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) v += part[threadIdx.x * kRowWarps + w];
      v /= sqrt(*norm_prev);
      y_next[out_row] = v;
      atomicAdd(norm_next, v * v);
    }
  }
}

// x = y_next / |y_next| and max_i |x[i] - y_prev[i] / |y_prev|| (bits of a non-negative double, atomic max).
__global__ void __launch_bounds__(256) k_power_diff(const double *__restrict__ y_prev, const double *__restrict__ norm_prev,
                                                     const double *__restrict__ y_next, const double *__restrict__ norm_next,
                                                     int n, double *__restrict__ x, unsigned long long *__restrict__ max_diff_bits) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  double d = 0.0;
  if (i < n) {
    const double v = y_next[i] / sqrt(*norm_next);
    d = fabs(v - y_prev[i] / sqrt(*norm_prev));
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

int apap_power_iterate(const double *m, int n, double *y, double *norms, int first_step, int steps, double *x,
                       unsigned long long *max_diff_bits, void *stream) {
  if (!m || !y || !norms || !x || !max_diff_bits || n <= 0 || first_step < 0 || steps <= 0)
    return fail(APAP_E_BADARG, "power_iterate: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_cuda(cudaMemsetAsync(max_diff_bits, 0, sizeof(unsigned long long), st), "power_iterate: memset");
  if (rc) return rc;
  int k = first_step;
  for (int s = 0; s < steps; ++s, ++k)
    k_power_step<<<(n + kMvRows - 1) / kMvRows, kMvWarps * 32, 0, st>>>(
        m, y + (size_t)(k & 1) * n, norms + k % 3, n, y + (size_t)((k + 1) & 1) * n, norms + (k + 1) % 3, norms + (k + 2) % 3);
  rc = check_cuda(cudaGetLastError(), "k_power_step launch");
  if (rc) return rc;
  // k = index of the iterate the last step produced; k - 1 the one before
  k_power_diff<<<(n + 255) / 256, 256, 0, st>>>(y + (size_t)((k - 1) & 1) * n, norms + (k - 1) % 3, y + (size_t)(k & 1) * n,
                                                norms + k % 3, n, x, max_diff_bits);
  return check_cuda(cudaGetLastError(), "k_power_diff launch");
}

}  // extern "C"
