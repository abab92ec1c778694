// Exact 1-nearest-neighbour descriptor matching -- the matcher step of the reference's keypoint-pair producer
// (pyviz/utils.py:149-150: cv.FlannBasedMatcher().match(feats_cp, feats_op)), SURVEY 8f row N3.
//
// FLANN's randomised kd-trees are approximate and not reproducible from run to run (tools/n3_host_cost.py: 92 % of the
// matches at c2 are the true nearest neighbour, 83 % identical over three runs), so the parity target is the EXACT
// nearest neighbour, cv.BFMatcher(cv.NORM_L2).match on the same descriptors: train index of the smallest
// sum_k (q_k - t_k)^2 in float32 (the lowest index on a tie), distance = sqrt of that sum.  SIFT descriptors are
// integers 0..255 stored as float32, so every partial sum is an exact integer < 2^24 in any summation order and the
// result is bit-identical to OpenCV's; for general float descriptors the index can differ only where the two smallest
// distances agree to float32 rounding.
//
// CTA = 64 queries x a slice of the train set, 128 threads, thread = 4 queries x 4 train descriptors per tile of 32
// (4 x 8 per tile of 64 is 11 % faster at 20 000 x 20 000 and 16 % slower at the 4 000 x 4 000 of c2: fewer, larger CTAs),
// descriptor pairs of components on the packed FP32x2 pipe (FADD2 + FFMA2 per two components of a pair).  Query and
// train tiles are staged in shared memory as [component pair][descriptor] float2 (two broadcast LDS.128 per operand and
// pair of components).  The per-query minimum goes to global memory as atomicMin on (distance bits << 32 | train index): the
// distance is a non-negative float, its bits order like the value, ties fall to the lowest index.
#include "common.cuh"

namespace apap {

constexpr int kMatchQ = 64;        // queries per CTA
constexpr int kMatchT = 32;        // train descriptors per tile
constexpr int kMatchThreads = 128;
constexpr int kMatchMaxDim = 256;   // both tiles of all components in shared memory: 100 KB at 256
constexpr int kMatchPad = 2;       // float2 columns of padding per shared-memory row (rows stay 16-byte aligned)
constexpr int kQPitch = kMatchQ + kMatchPad, kTPitch = kMatchT + kMatchPad;

__global__ void __launch_bounds__(kMatchThreads) k_match_nn(const float *__restrict__ query, const float *__restrict__ train,
                                                            int nq, int nt, int dim, int t_per_cta,
                                                            unsigned long long *__restrict__ best) {
  extern __shared__ __align__(16) unsigned char match_smem[];
  const int pairs = dim / 2;                                        // dim is even (checked by the launcher)
  float2 *qs = reinterpret_cast<float2 *>(match_smem);              // [pairs][kMatchQ + kMatchPad]
  float2 *ts = qs + (size_t)pairs * kQPitch;                        // [pairs][kMatchT + kMatchPad]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kMatchQ;
  const int t_begin = blockIdx.y * t_per_cta, t_end = min(nt, t_begin + t_per_cta);
  // a warp moves one descriptor at a time: coalesced 8-byte reads along the components, transposed into
  // [component pair][descriptor] (the two padding columns keep those stores at two per bank); rows past the end are zero
  for (int q = warp; q < kMatchQ; q += kMatchThreads / 32) {
    const bool ok = q0 + q < nq;
    const float2 *row = reinterpret_cast<const float2 *>(query + (size_t)(ok ? q0 + q : 0) * dim);
    for (int k = lane; k < pairs; k += 32) qs[k * kQPitch + q] = ok ? row[k] : make_float2(0.f, 0.f);
  }
  const int qg = tid & 15, tg = tid >> 4;                           // 16 query groups of 4, 8 train groups of 4
  float best_d[4];
  int best_i[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) { best_d[a] = INFINITY; best_i[a] = 0x7fffffff; }
  for (int t0 = t_begin; t0 < t_end; t0 += kMatchT) {
    __syncthreads();                                                // the previous tile has been consumed (and qs is written)
    for (int t = warp; t < kMatchT; t += kMatchThreads / 32) {
      const bool ok = t0 + t < t_end;
      const float2 *row = reinterpret_cast<const float2 *>(train + (size_t)(ok ? t0 + t : 0) * dim);
      for (int k = lane; k < pairs; k += 32) ts[k * kTPitch + t] = ok ? row[k] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2 acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < pairs; ++k) {
      float2 qv[4], tv[4];
      // the thread's queries are 2 qg, 2 qg + 1, 32 + 2 qg, 33 + 2 qg: the 16-byte pieces of a quarter-warp are contiguous
      const float4 *qp = reinterpret_cast<const float4 *>(qs + k * kQPitch) + qg;
      const float4 *tp = reinterpret_cast<const float4 *>(ts + k * kTPitch + tg * 4);
      const float4 qa = qp[0], qb = qp[16], ta = tp[0], tb = tp[1];
      qv[0] = make_float2(qa.x, qa.y); qv[1] = make_float2(qa.z, qa.w); qv[2] = make_float2(qb.x, qb.y); qv[3] = make_float2(qb.z, qb.w);
      tv[0] = make_float2(-ta.x, -ta.y); tv[1] = make_float2(-ta.z, -ta.w); tv[2] = make_float2(-tb.x, -tb.y); tv[3] = make_float2(-tb.z, -tb.w);
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float2 d = __fadd2_rn(qv[a], tv[b]);
          acc[a][b] = __ffma2_rn(d, d, acc[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {                                 // ascending train index: strict < keeps the first minimum
        const int t = t0 + tg * 4 + b;
        const float d = acc[a][b].x + acc[a][b].y;
        if (t < t_end && d < best_d[a]) { best_d[a] = d; best_i[a] = t; }
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int q = q0 + (a >> 1) * 32 + 2 * qg + (a & 1);
    if (q < nq && best_i[a] != 0x7fffffff) {
      // NaN distances never win (d < best is false); -0 cannot occur (sum of squares)
      const unsigned long long key = ((unsigned long long)__float_as_uint(best_d[a]) << 32) | (unsigned)best_i[a];
      atomicMin(best + q, key);
    }
  }
}

__global__ void __launch_bounds__(256) k_match_finish(const unsigned long long *__restrict__ best, int nq, int *__restrict__ idx,
                                                      float *__restrict__ dist) {
  const int q = blockIdx.x * 256 + threadIdx.x;
  if (q >= nq) return;
  const unsigned long long key = best[q];
  const bool none = key == ~0ull;
  idx[q] = none ? -1 : (int)(unsigned)(key & 0xffffffffu);
  dist[q] = none ? INFINITY : __fsqrt_rn(__uint_as_float((unsigned)(key >> 32)));
}

int launch_match_nn(const float *query, const float *train, int nq, int nt, int dim, unsigned long long *scratch, int *idx,
                    float *dist, cudaStream_t st) {
  if (nq == 0) return 0;
  int rc = check_cuda(cudaMemsetAsync(scratch, 0xff, (size_t)nq * sizeof(unsigned long long), st), "match: memset");
  if (rc) return rc;
  if (nt > 0) {
    const int q_ctas = (nq + kMatchQ - 1) / kMatchQ;
    // about four CTAs per SM (their shared memory allows it at 128 components), each with at least four train tiles
    int slices = (4 * sm_count_cached() + q_ctas - 1) / q_ctas;
    const int max_slices = (nt + 4 * kMatchT - 1) / (4 * kMatchT);
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    int t_per_cta = (nt + slices - 1) / slices;
    t_per_cta = (t_per_cta + kMatchT - 1) / kMatchT * kMatchT;
    slices = (nt + t_per_cta - 1) / t_per_cta;
    const size_t smem = (size_t)(dim / 2) * (kQPitch + kTPitch) * sizeof(float2);
    static bool attr_set = false;
    if (!attr_set) {
      rc = check_cuda(cudaFuncSetAttribute(k_match_nn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (kMatchMaxDim / 2) * (kQPitch + kTPitch) * (int)sizeof(float2)),
                      "match: cudaFuncSetAttribute");
      if (rc) return rc;
      attr_set = true;
    }
    k_match_nn<<<dim3(q_ctas, slices), kMatchThreads, smem, st>>>(query, train, nq, nt, dim, t_per_cta, scratch);
    rc = check_cuda(cudaGetLastError(), "k_match_nn launch");
    if (rc) return rc;
  }
  k_match_finish<<<(nq + 255) / 256, 256, 0, st>>>(scratch, nq, idx, dist);
  return check_cuda(cudaGetLastError(), "k_match_finish launch");
}

}  // namespace apap

using namespace apap;

extern "C" int apap_match_nn(const float *query, const float *train, int nq, int nt, int dim, unsigned long long *scratch,
                             int *idx, float *dist, void *stream) {
  if (nq < 0 || nt < 0) return fail(APAP_E_BADARG, "match_nn: negative size");
  if (nq == 0) return 0;
  if (!query || (!train && nt) || !scratch || !idx || !dist) return fail(APAP_E_BADARG, "match_nn: null pointer");
  if (dim <= 0 || dim % 2 || dim > kMatchMaxDim) return fail(APAP_E_BADARG, "match_nn: dim must be even and <= 256");
  if ((reinterpret_cast<uintptr_t>(query) | reinterpret_cast<uintptr_t>(train)) & 7u)
    return fail(APAP_E_ALIGN, "match_nn: descriptors must be 8-byte aligned");
  if (reinterpret_cast<uintptr_t>(scratch) & 7u) return fail(APAP_E_ALIGN, "match_nn: scratch must be 8-byte aligned");
  return launch_match_nn(query, train, nq, nt, dim, scratch, idx, dist, static_cast<cudaStream_t>(stream));
}
