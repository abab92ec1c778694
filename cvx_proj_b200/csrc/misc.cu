// local_weight (second output of APAP.local_homography, reference pyviz/apap.py:150-153) and the
// FP32 FMA-pipe probe that measures the denominator of the Gram kernel's roofline fraction.
#include "common.cuh"

namespace apap {

// out[c][i] = max(exp(-sqrt(dx^2 + dy^2) * inv_sigma_sq), gamma) in float64, dx = anchor - keypoint
// in float64 (the caller promotes float32 keypoints exactly, as numpy promotes float64 - float32).  Pure streaming
// writes (8 B / element).
__global__ void __launch_bounds__(256) k_weight(const double *__restrict__ anchors, const double *__restrict__ kp_xy,
                                                 int n_kp, double inv_sigma_sq, double gamma,
                                                 double *__restrict__ out) {
  const int cell = blockIdx.y;
  const double vx = anchors[2 * cell], vy = anchors[2 * cell + 1];
  double *dst = out + (size_t)cell * n_kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_kp; i += gridDim.x * blockDim.x) {
    const double2 k = reinterpret_cast<const double2 *>(kp_xy)[i];
    const double dx = vx - k.x, dy = vy - k.y;
    const double w = exp(-(sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))) * inv_sigma_sq));
    dst[i] = w < gamma ? gamma : w;
  }
}

int launch_weight(const double *anchors, const double *kp_xy, int cells, int n_kp, double inv_sigma_sq, double gamma,
                  double *out, cudaStream_t st) {
  if (cells == 0 || n_kp == 0) return 0;
  if (cells > 65535) return fail(APAP_E_TOOBIG, "local_weight: at most 65535 cells per call (slice the grid)");
  int bx = (n_kp + 255) / 256;
  if (bx > 64) bx = 64;
  k_weight<<<dim3(bx, cells), 256, 0, st>>>(anchors, kp_xy, n_kp, inv_sigma_sq, gamma, out);
  return check_cuda(cudaGetLastError(), "k_weight launch");
}

// 16 independent FFMA chains per thread, 8 warps per CTA, 8 CTAs per SM: enough ILP and TLP to
// saturate the FMA pipe.  flops = 2 * 16 * iters * threads.
constexpr int kProbeThreads = 256;
__global__ void __launch_bounds__(kProbeThreads) k_probe(int iters, float seed, float *sink) {
  float a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = seed + (float)(threadIdx.x + k);
  const float m = 0.999f + seed, c = 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fmaf(a[k], m, c);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 123456.789f) *sink = s;   // never true; keeps the chains alive
}

// 16 independent MUFU.EX2 chains per thread: the XU (special-function) pipe at saturation.
__global__ void __launch_bounds__(kProbeThreads) k_probe_mufu(int iters, float seed, float *sink) {
  float a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = seed - 1e-3f * (float)(threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = ex2_approx(a[k]);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 123456.789f) *sink = s;   // never true; keeps the chains alive
}

// Broadcast of a contiguous buffer into every replica of an NVLS multicast mapping: coalesced 16-byte loads of the
// local copy, 16-byte multimem stores (each warp writes 512 contiguous bytes: full-size NVLink packets, which the
// 96-byte row fragments of the warp kernel's own multicast mode are not -- measured 2.5x faster at 2 GPUs).
__global__ void __launch_bounds__(256) k_multicast_copy(const uint4 *__restrict__ src, uint4 *mc_dst, size_t n_vec) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * 256) {
    const uint4 v = __ldcs(src + i);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i),
                 "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
  }
}

int launch_multicast_copy(const void *src, void *mc_dst, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return 0;
  const size_t n_vec = bytes / 16;
  size_t blocks = (n_vec + 255) / 256;
  const size_t cap = (size_t)sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  k_multicast_copy<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(mc_dst),
                                                      n_vec);
  return check_cuda(cudaGetLastError(), "k_multicast_copy launch");
}

// The same broadcast by unicast stores: each 16 bytes of the local band are written into every listed peer's buffer
// (peer-mapped device addresses, e.g. torch symmetric memory's buffer_ptrs + offset).  A rank sends its band once per
// peer, but no GPU receives its own band back from the switch, as it does through a multicast mapping that includes
// it: (N-1)/N of the panorama comes in per GPU instead of all of it.
struct PeerList {
  uint4 *dst[APAP_MAX_PEERS];
  int n;
};

__global__ void __launch_bounds__(256) k_peer_copy(const uint4 *__restrict__ src, PeerList pl, size_t n_vec) {
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_vec; i += 2 * stride) {
    const bool two = i + stride < n_vec;
    const uint4 a = __ldcs(src + i);
    uint4 b = a;
    if (two) b = __ldcs(src + i + stride);
#pragma unroll
    for (int k = 0; k < APAP_MAX_PEERS; ++k) {           // unrolled: the list stays in the parameter bank
      if (k < pl.n) {
        pl.dst[k][i] = a;
        if (two) pl.dst[k][i + stride] = b;
      }
    }
  }
}

int launch_peer_copy(const void *src, void *const *peers, int n_peers, size_t bytes, cudaStream_t st) {
  if (bytes == 0 || n_peers == 0) return 0;
  PeerList pl;
  pl.n = n_peers;
  for (int k = 0; k < n_peers; ++k) pl.dst[k] = reinterpret_cast<uint4 *>(peers[k]);
  const size_t n_vec = bytes / 16;
  size_t blocks = (n_vec + 511) / 512;
  const size_t cap = (size_t)sm_count_cached() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_peer_copy<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint4 *>(src), pl, n_vec);
  return check_cuda(cudaGetLastError(), "k_peer_copy launch");
}

int launch_probe(int kind, int iters, float *sink, double *ops, cudaStream_t st) {
  const int blocks = sm_count_cached() * 8;
  if (kind == APAP_PROBE_FFMA) {
    k_probe<<<blocks, kProbeThreads, 0, st>>>(iters, 0.f, sink);
    if (ops) *ops = 2.0 * 16.0 * (double)iters * (double)blocks * kProbeThreads;
  } else if (kind == APAP_PROBE_MUFU) {
    k_probe_mufu<<<blocks, kProbeThreads, 0, st>>>(iters, 0.f, sink);
    if (ops) *ops = 16.0 * (double)iters * (double)blocks * kProbeThreads;
  } else {
    return fail(APAP_E_BADARG, "probe: unknown kind");
  }
  return check_cuda(cudaGetLastError(), "k_probe launch");
}

}  // namespace apap
