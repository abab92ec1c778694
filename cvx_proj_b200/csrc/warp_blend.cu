// K3 -- mesh warp (reference APAP.local_warp pixel loop, pyviz/apap.py:206-215), optionally
// fused with K4, and K4 -- uniform_blend (pyviz/apap_utils.py:75-88).
//
// HBM-bound byte work: every canvas pixel is written once (3 B), every source pixel is read at
// most once from HBM.  Thread layout: a warp owns 32 consecutive canvas columns x P consecutive
// rows; lane l works on column j0 + l, so in every row the 32 lanes gather ~96 contiguous source
// bytes (1-2 L1 wavefronts per byte plane instead of one line per lane) and their 32 pixels are
// 96 contiguous output bytes, which the warp re-packs with two shuffles + one byte-permute into
// 24 aligned 32-bit stores.  A thread keeps its cell column (and the cell record) across its P
// rows and reloads only when the cell row changes.
//
// Pixel selection must equal the reference's float64 arithmetic (float32 H^-1 promoted to
// float64, divide, strict bounds, truncate).  The float32 fast path works in cell-relative
// coordinates: the host rewrites each cell's H^-1 relative to the cell's first pixel (dx, dy
// small) and to an integer base (qbx, qby) near the cell's source position, and normalises the
// denominator to ~1, so numerators and quotient are O(cell size) instead of O(image size) and the
// float32 error is ~1e-5 px:   src_x = qbx + floor((A0 dx + B0 dy + C0) / (A2 dx + B2 dy + C2)).
// The host also supplies, per cell, a rigorous bound eps on that error.  A quotient farther than
// eps from every integer has the same floor and the same bounds decision in both arithmetics;
// the others (~1e-4 of the pixels) are recomputed in float64 exactly as the reference does.
#include "common.cuh"

namespace apap {

constexpr int kWarpThreads = 256;
constexpr int kWarpsPerCta = kWarpThreads / 32;
constexpr int kRowsPerThread = 8;

struct WarpParams {
  const uint8_t *src;
  const float4 *cell_fast;     // [cells][3] float4: A0 B0 C0 A1 | B1 C1 A2 B2 | C2 qbx' qby' (int bits) 0.5-eps
  const float *cell_hinv;      // [cells][9]: the reference's inverted grid (float64 path only)
  const uint2 *col_lut;        // [canvas_w]: {cell column, float bits of x - cell's first x}
  const uint2 *row_lut;        // [canvas_h]: {cell row,    float bits of y - cell's first y}
  const uint8_t *centre;
  uint8_t *out;                // first byte of canvas row `row0`
  int row0, row1;              // band of canvas rows
  int chunks_per_row;          // ceil(canvas_w / 32)
  int n_warps;                 // chunks_per_row * ceil((row1 - row0) / P)
  int src_h, src_w;
  int grid_cols;
  int canvas_w;
  int off_x, off_y;
  int centre_h, centre_w;
  int force_exact;
  int word_stores;             // canvas_w % 4 == 0 and out 4-byte aligned: packed 32-bit stores
};

constexpr float kMagic = 12582912.f;          // 1.5 * 2^23: x + kMagic (round down) = floor(x) in the mantissa

// float64 path = the reference's arithmetic (pyviz/apap.py:182-183,211-215): float32 H^-1 promoted
// to float64, IEEE divide, strict bounds, truncation.  Returns the source pixel index or -1.
__device__ __noinline__ int exact_lookup(const float *__restrict__ h, int x, int y, int src_w, int src_h) {
  const double xd = (double)x, yd = (double)y;
  const double t0 = (double)h[0] * xd + (double)h[1] * yd + (double)h[2];
  const double t1 = (double)h[3] * xd + (double)h[4] * yd + (double)h[5];
  const double t2 = (double)h[6] * xd + (double)h[7] * yd + (double)h[8];
  const double tx = t0 / t2;
  const double ty = t1 / t2;
  if (0.0 < tx && tx < (double)src_w && 0.0 < ty && ty < (double)src_h) return (int)ty * src_w + (int)tx;
  return -1;
}

__device__ __forceinline__ uint32_t load_px(const uint8_t *__restrict__ p) {
  const uint32_t b0 = __ldg(p), b1 = __ldg(p + 1), b2 = __ldg(p + 2);
  return __byte_perm(__byte_perm(b0, b1, 0x1140), b2, 0x3410);     // zero-extended bytes -> b0 | b1<<8 | b2<<16
}

// One canvas row of one lane.  kWords: the band's rows are 4-byte aligned (canvas_w % 4 == 0), the
// warp's 32 pixels leave as 24 packed 32-bit stores; otherwise three byte stores per lane.
template <bool kBlend, bool kWords>
__device__ __forceinline__ void warp_row(const WarpParams &p, const uint8_t *__restrict__ src, int i, int x, bool col_ok,
                                         const uint2 cl, int &cur_row_cell, int &cell, float dxf, float &b0,
                                         float &b1, float &b2, float &m0, float &m1, float &m2, float &hme,
                                         int &qbx, int &qby, int lane_a, int lane_b, uint32_t sel, bool store_ok,
                                         uint8_t *dst) {
  const uint2 rl = __ldg(p.row_lut + i);           // same address in every lane
  if ((int)rl.x != cur_row_cell) {                 // warp-uniform: new cell row -> new record
    cur_row_cell = (int)rl.x;
    cell = cur_row_cell * p.grid_cols + (int)cl.x;
    const float4 *rec = p.cell_fast + (size_t)cell * 3;
    const float4 u = __ldg(rec), v = __ldg(rec + 1), w = __ldg(rec + 2);
    m0 = fmaf(u.x, dxf, u.z); b0 = u.y;
    m1 = fmaf(u.w, dxf, v.y); b1 = v.x;
    m2 = fmaf(v.z, dxf, w.x); b2 = v.w;
    qbx = __float_as_int(w.y);
    qby = __float_as_int(w.z);
    hme = p.force_exact ? -1.f : w.w;
  }
  const float dyf = __uint_as_float(rl.y);
  const float n0 = fmaf(b0, dyf, m0);
  const float n1 = fmaf(b1, dyf, m1);
  const float d = fmaf(b2, dyf, m2);
  const float r = rcp_approx(d);
  const float qx = n0 * r;
  const float qy = n1 * r;
  const float tx = add_rd(qx, kMagic);
  const float ty = add_rd(qy, kMagic);
  const float fx = qx - (tx - kMagic);             // exact fractional parts in [0, 1)
  const float fy = qy - (ty - kMagic);
  // hme = 0.5 - eps; a record with hme < 0 (degenerate cell, forced) never passes; NaN never passes
  const bool clear = fmaxf(fabsf(fx - 0.5f), fabsf(fy - 0.5f)) <= hme;
  const int ix = __float_as_int(tx) + qbx;         // qbx already holds (base - bits(kMagic))
  const int iy = __float_as_int(ty) + qby;
  bool inb = ((unsigned)ix < (unsigned)p.src_w) & ((unsigned)iy < (unsigned)p.src_h);
  int idx = iy * p.src_w + ix;
  if (!clear) {
    idx = exact_lookup(p.cell_hinv + (size_t)cell * 9, x, i - p.off_y, p.src_w, p.src_h);
    inb = idx >= 0;
  }
  uint32_t val = 0;
  if (inb) val = load_px(src + (size_t)(unsigned)idx * 3);
  if (kBlend) {
    // centre image pasted at (off_x, off_y) (pyviz/apap.py:259-260), then uniform_blend
    const int cy = i - p.off_y;
    if ((unsigned)cy < (unsigned)p.centre_h && (unsigned)x < (unsigned)p.centre_w && col_ok) {
      const uint32_t cv = load_px(p.centre + ((size_t)cy * p.centre_w + x) * 3);
      if (cv != 0) val = (val != 0) ? __vhaddu4(val, cv) : cv;
    }
  }
  if (kWords) {
    const uint32_t va = __shfl_sync(0xffffffffu, val, lane_a);
    const uint32_t vb = __shfl_sync(0xffffffffu, val, lane_b);
    if (store_ok) *reinterpret_cast<uint32_t *>(dst) = __byte_perm(va, vb, sel);
  } else if (store_ok) {
    dst[0] = (uint8_t)val;
    dst[1] = (uint8_t)(val >> 8);
    dst[2] = (uint8_t)(val >> 16);
  }
}

template <bool kBlend, bool kWords>
__global__ void __launch_bounds__(kWarpThreads) k_warp(const WarpParams p) {
  constexpr int P = kRowsPerThread;
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (wid >= p.n_warps) return;                    // warp-uniform
  const int rg = wid / p.chunks_per_row;
  const int j0 = (wid - rg * p.chunks_per_row) * 32;
  const int i0 = p.row0 + rg * P;
  const int n_rows = min(P, p.row1 - i0);
  const int j = j0 + lane;
  const bool col_ok = j < p.canvas_w;
  const uint2 cl = __ldg(p.col_lut + (col_ok ? j : p.canvas_w - 1));
  const float dxf = __uint_as_float(cl.y);
  const int x = j - p.off_x;
  const uint8_t *__restrict__ src = p.src;

  // store side.  kWords: output word w = lane (< 24) takes its 4 bytes from the pixels of lanes
  // lane_a = (4w)/3 and lane_a + 1, starting at byte (4w) % 3 of the first
  const int lane_a = (lane + lane / 3) & 31, lane_b = (lane_a + 1) & 31;
  const uint32_t sel = (lane % 3 == 0) ? 0x4210u : (lane % 3 == 1) ? 0x5421u : 0x6542u;
  const bool store_ok = kWords ? lane < min(24, (3 * (p.canvas_w - j0)) >> 2) : col_ok;
  const size_t pitch = (size_t)p.canvas_w * 3;
  uint8_t *dst = p.out + ((size_t)(i0 - p.row0) * p.canvas_w + j0) * 3 + (kWords ? 4 : 3) * lane;

  float b0 = 0.f, b1 = 0.f, b2 = 0.f, m0 = 0.f, m1 = 0.f, m2 = 1.f, hme = -1.f;
  int qbx = 0, qby = 0;
  int cur_row_cell = -1, cell = 0;

  if (n_rows == P) {
#pragma unroll
    for (int k = 0; k < P; ++k)
      warp_row<kBlend, kWords>(p, src, i0 + k, x, col_ok, cl, cur_row_cell, cell, dxf, b0, b1, b2, m0, m1, m2, hme,
                               qbx, qby, lane_a, lane_b, sel, store_ok, dst + k * pitch);
  } else {
    for (int k = 0; k < n_rows; ++k)
      warp_row<kBlend, kWords>(p, src, i0 + k, x, col_ok, cl, cur_row_cell, cell, dxf, b0, b1, b2, m0, m1, m2, hme,
                               qbx, qby, lane_a, lane_b, sel, store_ok, dst + k * pitch);
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

int launch_warp(const uint8_t *src, int src_h, int src_w, const float *cell_fast, const float *cell_hinv,
                const uint32_t *col_lut, const uint32_t *row_lut, int grid_cols, int canvas_w, int canvas_h, int off_x,
                int off_y, int row0, int row1, const uint8_t *centre, int centre_h, int centre_w, uint8_t *out_band,
                int force_exact, cudaStream_t st) {
  if (row0 < 0 || row1 > canvas_h || row0 > row1) return fail(APAP_E_BADARG, "warp: bad row band");
  if ((long long)src_w * src_h > 2147483647LL)
    return fail(APAP_E_TOOBIG, "warp: source image has more than 2^31-1 pixels");
  if ((reinterpret_cast<uintptr_t>(cell_fast) & 15u) || (reinterpret_cast<uintptr_t>(col_lut) & 7u) ||
      (reinterpret_cast<uintptr_t>(row_lut) & 7u))
    return fail(APAP_E_ALIGN, "warp: cell_fast must be 16-byte, col_lut / row_lut 8-byte aligned");
  if (row0 == row1) return 0;
  WarpParams p;
  p.src = src; p.cell_fast = reinterpret_cast<const float4 *>(cell_fast); p.cell_hinv = cell_hinv;
  p.col_lut = reinterpret_cast<const uint2 *>(col_lut); p.row_lut = reinterpret_cast<const uint2 *>(row_lut);
  p.centre = centre; p.out = out_band;
  p.row0 = row0; p.row1 = row1;
  p.chunks_per_row = (canvas_w + 31) / 32;
  const long long n_warps = (long long)p.chunks_per_row * ((row1 - row0 + kRowsPerThread - 1) / kRowsPerThread);
  if (n_warps > 2147483647LL) return fail(APAP_E_TOOBIG, "warp: canvas band too large");
  p.n_warps = (int)n_warps;
  p.src_h = src_h; p.src_w = src_w; p.grid_cols = grid_cols; p.canvas_w = canvas_w;
  p.off_x = off_x; p.off_y = off_y; p.centre_h = centre_h; p.centre_w = centre_w; p.force_exact = force_exact;
  p.word_stores = (canvas_w % 4 == 0) && !(reinterpret_cast<uintptr_t>(out_band) & 3u);
  const unsigned blocks = (unsigned)((n_warps + kWarpsPerCta - 1) / kWarpsPerCta);
  if (centre) {
    if (p.word_stores) k_warp<true, true><<<blocks, kWarpThreads, 0, st>>>(p);
    else k_warp<true, false><<<blocks, kWarpThreads, 0, st>>>(p);
  } else {
    if (p.word_stores) k_warp<false, true><<<blocks, kWarpThreads, 0, st>>>(p);
    else k_warp<false, false><<<blocks, kWarpThreads, 0, st>>>(p);
  }
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
