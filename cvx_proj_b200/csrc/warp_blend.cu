// K3 -- mesh warp (reference APAP.local_warp pixel loop, pyviz/apap.py:206-215), optionally
// fused with K4, and K4 -- uniform_blend (pyviz/apap_utils.py:75-88).
//
// HBM-bound byte work: every canvas pixel is written once (3 B), every source pixel is read at
// most once from HBM (gathers of neighbouring canvas pixels hit neighbouring source pixels, so
// L1/L2 absorb the reuse).  One thread owns 4 consecutive canvas pixels = 12 contiguous output
// bytes = three aligned 32-bit stores, a warp writes 384 contiguous bytes.
//
// Pixel selection must equal the reference's float64 arithmetic (float32 H^-1 promoted to
// float64, divide, strict bounds, truncate).  A float32 fast path evaluates the coordinates with
// a residual-corrected division; the host supplies, per cell, a rigorous bound (eps_x, eps_y) on
// the fast path's absolute coordinate error inside that cell.  A coordinate farther than eps from
// every integer has the same floor and the same bounds decision in both arithmetics; the others
// (about 1 % of pixels at 8K) are recomputed in float64 exactly as the reference does.
#include "common.cuh"

namespace apap {

constexpr int kWarpThreads = 256;
constexpr int kPxPerThread = 4;

struct WarpParams {
  const uint8_t *src;
  const float *hinv;
  const uint16_t *col_cell;
  const uint16_t *row_cell;
  const uint8_t *centre;
  uint8_t *out;
  long long n_px;        // pixels in the band
  long long px0;         // absolute index of the band's first pixel
  int src_h, src_w;
  int grid_cols;
  int canvas_w;
  int off_x, off_y;
  int centre_h, centre_w;
  int force_exact;
};

constexpr float kMagic = 12582912.f;          // 1.5 * 2^23: x + kMagic (round down) = floor(x) in the mantissa
constexpr int kMagicBits = 0x4B400000;

// Source byte offset of the pixel the reference would copy, or -1 when it leaves the canvas
// pixel black.  (x, y) are the canvas coordinates minus the offsets.
__device__ __forceinline__ int exact_lookup(const float (&h)[9], int x, int y, int src_w, int src_h) {
  const double xd = (double)x, yd = (double)y;
  const double t0 = (double)h[0] * xd + (double)h[1] * yd + (double)h[2];
  const double t1 = (double)h[3] * xd + (double)h[4] * yd + (double)h[5];
  const double t2 = (double)h[6] * xd + (double)h[7] * yd + (double)h[8];
  const double tx = t0 / t2;
  const double ty = t1 / t2;
  if (0.0 < tx && tx < (double)src_w && 0.0 < ty && ty < (double)src_h) return ((int)ty * src_w + (int)tx) * 3;
  return -1;
}

__device__ __forceinline__ int fast_lookup(const float (&h)[9], float eps_x, float eps_y, int x, int y, int src_w,
                                           int src_h, bool &flagged) {
  const float xf = (float)x, yf = (float)y;
  const float t0 = fmaf(h[0], xf, fmaf(h[1], yf, h[2]));
  const float t1 = fmaf(h[3], xf, fmaf(h[4], yf, h[5]));
  const float t2 = fmaf(h[6], xf, fmaf(h[7], yf, h[8]));
  const float r = rcp_approx(t2);
  float qx = t0 * r;
  float qy = t1 * r;
  qx = fmaf(fmaf(-qx, t2, t0), r, qx);       // residual correction: ~1 ulp of the true quotient
  qy = fmaf(fmaf(-qy, t2, t1), r, qy);
  qx = fminf(fmaxf(qx, -0.5f), 4194303.5f);   // NaN -> -0.5 (out of bounds)
  qy = fminf(fmaxf(qy, -0.5f), 4194303.5f);
  const float mx = add_rd(qx, kMagic);
  const float my = add_rd(qy, kMagic);
  const float fx = qx - (mx - kMagic);        // exact fractional parts in [0, 1)
  const float fy = qy - (my - kMagic);
  flagged = (fabsf(fx - 0.5f) > 0.5f - eps_x) || (fabsf(fy - 0.5f) > 0.5f - eps_y);
  const int ix = __float_as_int(mx) - kMagicBits;
  const int iy = __float_as_int(my) - kMagicBits;
  if ((unsigned)ix < (unsigned)src_w && (unsigned)iy < (unsigned)src_h) return (iy * src_w + ix) * 3;
  return -1;
}

template <bool kBlend>
__global__ void __launch_bounds__(kWarpThreads) k_warp(const WarpParams p) {
  const long long group = (long long)blockIdx.x * kWarpThreads + threadIdx.x;
  const long long rel0 = group * kPxPerThread;
  if (rel0 >= p.n_px) return;
  const long long abs0 = p.px0 + rel0;
  int i = (int)(abs0 / p.canvas_w);
  int j = (int)(abs0 - (long long)i * p.canvas_w);
  int row_c = __ldg(p.row_cell + i);

  float h[9];
  float eps_x = 1.f, eps_y = 1.f;
  int cur_cell = -1;
  uint32_t px[kPxPerThread];     // 0x00RRGGBB-style packed 3-byte pixels (byte 0 = channel 0)

#pragma unroll
  for (int k = 0; k < kPxPerThread; ++k) {
    uint32_t val = 0;
    if (rel0 + k < p.n_px) {
      const int cell = row_c * p.grid_cols + (int)__ldg(p.col_cell + j);
      if (cell != cur_cell) {
        const float4 *hp = reinterpret_cast<const float4 *>(p.hinv + (size_t)cell * kHinvRow);
        const float4 a = __ldg(hp), b = __ldg(hp + 1), c = __ldg(hp + 2);
        h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w;
        h[4] = b.x; h[5] = b.y; h[6] = b.z; h[7] = b.w;
        h[8] = c.x; eps_x = c.y; eps_y = c.z;
        cur_cell = cell;
      }
      const int x = j - p.off_x, y = i - p.off_y;
      bool flagged;
      int off = fast_lookup(h, eps_x, eps_y, x, y, p.src_w, p.src_h, flagged);
      if (flagged || p.force_exact) off = exact_lookup(h, x, y, p.src_w, p.src_h);
      if (off >= 0) {
        const uint8_t *s = p.src + off;
        val = (uint32_t)__ldg(s) | ((uint32_t)__ldg(s + 1) << 8) | ((uint32_t)__ldg(s + 2) << 16);
      }
      if (kBlend) {
        // centre image pasted at (off_x, off_y) (pyviz/apap.py:259-260), then uniform_blend
        const int cy = i - p.off_y, cx = j - p.off_x;
        if ((unsigned)cy < (unsigned)p.centre_h && (unsigned)cx < (unsigned)p.centre_w) {
          const uint8_t *c = p.centre + ((size_t)cy * p.centre_w + cx) * 3;
          const uint32_t cv = (uint32_t)__ldg(c) | ((uint32_t)__ldg(c + 1) << 8) | ((uint32_t)__ldg(c + 2) << 16);
          if (cv != 0) val = (val != 0) ? __vhaddu4(val, cv) : cv;
        }
      }
      if (++j == p.canvas_w) {
        j = 0;
        ++i;
        if (rel0 + k + 1 < p.n_px) row_c = __ldg(p.row_cell + i);
      }
    }
    px[k] = val;
  }

  uint8_t *dst = p.out + rel0 * 3;
  if (rel0 + kPxPerThread <= p.n_px) {
    uint32_t *d = reinterpret_cast<uint32_t *>(dst);
    d[0] = px[0] | (px[1] << 24);
    d[1] = (px[1] >> 8) | (px[2] << 16);
    d[2] = (px[2] >> 16) | (px[3] << 8);
  } else {
    for (int k = 0; k < kPxPerThread && rel0 + k < p.n_px; ++k) {
      dst[3 * k + 0] = (uint8_t)(px[k]);
      dst[3 * k + 1] = (uint8_t)(px[k] >> 8);
      dst[3 * k + 2] = (uint8_t)(px[k] >> 16);
    }
  }
}

// ------------------------------------------------------------------------------------ K4 blend
// 16 pixels = 48 B = three 16-byte vectors per thread and per image.
constexpr int kBlendThreads = 256;

__device__ __forceinline__ uint32_t get_byte(const uint32_t (&w)[12], int o) { return (w[o >> 2] >> ((o & 3) * 8)) & 0xffu; }

__global__ void __launch_bounds__(kBlendThreads) k_blend(const uint4 *__restrict__ a, const uint4 *__restrict__ b,
                                                          uint4 *__restrict__ out, long long n_vec3,
                                                          const uint8_t *__restrict__ a8, const uint8_t *__restrict__ b8,
                                                          uint8_t *__restrict__ out8, long long n_px) {
  const long long gid = (long long)blockIdx.x * kBlendThreads + threadIdx.x;
  if (gid < n_vec3) {
    uint32_t wa[12], wb[12], wo[12];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const uint4 x = __ldcs(a + gid * 3 + v);
      const uint4 y = __ldcs(b + gid * 3 + v);
      wa[4 * v + 0] = x.x; wa[4 * v + 1] = x.y; wa[4 * v + 2] = x.z; wa[4 * v + 3] = x.w;
      wb[4 * v + 0] = y.x; wb[4 * v + 1] = y.y; wb[4 * v + 2] = y.z; wb[4 * v + 3] = y.w;
    }
    // per-byte mask: 0xff where the byte's pixel is non-black in BOTH images
    uint32_t both[12];
#pragma unroll
    for (int w = 0; w < 12; ++w) both[w] = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t na = get_byte(wa, 3 * k) | get_byte(wa, 3 * k + 1) | get_byte(wa, 3 * k + 2);
      const uint32_t nb = get_byte(wb, 3 * k) | get_byte(wb, 3 * k + 1) | get_byte(wb, 3 * k + 2);
      if (na != 0 && nb != 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) both[(3 * k + c) >> 2] |= 0xffu << (((3 * k + c) & 3) * 8);
      }
    }
#pragma unroll
    for (int w = 0; w < 12; ++w) {
      const uint32_t avg = __vhaddu4(wa[w], wb[w]);     // per-byte (a + b) >> 1
      const uint32_t sum = __vadd4(wa[w], wb[w]);       // per-byte a + b (one of them is 0)
      wo[w] = (avg & both[w]) | (sum & ~both[w]);
    }
#pragma unroll
    for (int v = 0; v < 3; ++v)
      __stcs(out + gid * 3 + v, make_uint4(wo[4 * v], wo[4 * v + 1], wo[4 * v + 2], wo[4 * v + 3]));
  } else if (gid == n_vec3) {
    // tail: fewer than 16 pixels
    for (long long px = n_vec3 * 16; px < n_px; ++px) {
      const uint8_t *pa = a8 + px * 3, *pb = b8 + px * 3;
      const bool both = (pa[0] | pa[1] | pa[2]) && (pb[0] | pb[1] | pb[2]);
      for (int c = 0; c < 3; ++c) {
        const unsigned s = (unsigned)pa[c] + pb[c];
        out8[px * 3 + c] = (uint8_t)(both ? (s >> 1) : s);
      }
    }
  }
}

int launch_warp(const uint8_t *src, int src_h, int src_w, const float *hinv, const uint16_t *col_cell,
                const uint16_t *row_cell, int grid_cols, int canvas_w, int canvas_h, int off_x, int off_y, int row0,
                int row1, const uint8_t *centre, int centre_h, int centre_w, uint8_t *out_band, int force_exact,
                cudaStream_t st) {
  if (row0 < 0 || row1 > canvas_h || row0 > row1) return fail(APAP_E_BADARG, "warp: bad row band");
  if (src_w > 4194302 || src_h > 4194302 || (long long)src_w * src_h * 3 > 2147483647LL)
    return fail(APAP_E_TOOBIG, "warp: source image too large for 32-bit byte offsets");
  if ((reinterpret_cast<uintptr_t>(out_band) & 3u) || (reinterpret_cast<uintptr_t>(hinv) & 15u))
    return fail(APAP_E_ALIGN, "warp: out_band must be 4-byte and hinv 16-byte aligned");
  WarpParams p;
  p.src = src; p.hinv = hinv; p.col_cell = col_cell; p.row_cell = row_cell; p.centre = centre; p.out = out_band;
  p.n_px = (long long)(row1 - row0) * canvas_w;
  p.px0 = (long long)row0 * canvas_w;
  p.src_h = src_h; p.src_w = src_w; p.grid_cols = grid_cols; p.canvas_w = canvas_w;
  p.off_x = off_x; p.off_y = off_y; p.centre_h = centre_h; p.centre_w = centre_w; p.force_exact = force_exact;
  if (p.n_px == 0) return 0;
  const long long groups = (p.n_px + kPxPerThread - 1) / kPxPerThread;
  const long long blocks = (groups + kWarpThreads - 1) / kWarpThreads;
  if (blocks > 2147483647LL) return fail(APAP_E_TOOBIG, "warp: canvas too large");
  if (centre)
    k_warp<true><<<(unsigned)blocks, kWarpThreads, 0, st>>>(p);
  else
    k_warp<false><<<(unsigned)blocks, kWarpThreads, 0, st>>>(p);
  return check_cuda(cudaGetLastError(), "k_warp launch");
}

int launch_blend(const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n_px, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15u)
    return fail(APAP_E_ALIGN, "blend: pointers must be 16-byte aligned");
  if (n_px == 0) return 0;
  const long long n_vec3 = (long long)(n_px / 16);
  const long long threads = n_vec3 + 1;   // +1 thread for the tail
  const long long blocks = (threads + kBlendThreads - 1) / kBlendThreads;
  if (blocks > 2147483647LL) return fail(APAP_E_TOOBIG, "blend: image too large");
  k_blend<<<(unsigned)blocks, kBlendThreads, 0, st>>>(reinterpret_cast<const uint4 *>(a),
                                                      reinterpret_cast<const uint4 *>(b),
                                                      reinterpret_cast<uint4 *>(out), n_vec3, a, b, out,
                                                      (long long)n_px);
  return check_cuda(cudaGetLastError(), "k_blend launch");
}

}  // namespace apap
