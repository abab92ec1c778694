// Pieces shared by the two mesh-warp engines (csrc/warp_tile.cu, the tile engine, and the legacy
// strip kernel in csrc/warp_blend.cu): launch parameters, the float64 reference lookup, the per-lane
// cell state of the float32 fast path.
#pragma once

#include "common.cuh"

namespace apap {

constexpr int kBlockRows = APAP_WARP_BLOCK_ROWS;      // rows per row block (processed as 2 pairs)
static_assert(kBlockRows == 4, "the kernel processes a row block as two row pairs");

struct WarpParams {
  const uint8_t *src;
  const float4 *cell_fast;     // [cells][3] float4: A0 B0 C0 A1 | B1 C1 A2 B2 | C2 qbx' qby' (int bits) g
  const float *cell_hinv;      // [cells][9]: the reference's inverted grid (float64 path only)
  const uint2 *col_lut;        // [canvas_w]: {cell column, float bits of x - cell's first x}
  const uint2 *row_blocks;     // [n_blocks]: {first canvas row | rows << 28, cell row | dy of the first row << 16}
  const uint8_t *centre;
  uint8_t *out;                // first byte of canvas row `row0`
  int row0;                    // first canvas row of the band `out` holds
  int band_rows;               // canvas rows `out` holds
  int n_blocks;
  int chunks_per_row;          // ceil(canvas_w / 32)
  int src_h, src_w;
  int grid_cols;
  int canvas_w;
  int off_x, off_y;
  int centre_h, centre_w;
  int force_exact;
  int multicast;               // out is an NVLS multicast address: every store goes to all GPUs of the group
};

// One 32-bit store to every replica of a multicast (NVLS) mapping: the NVSwitch fans it out.
__device__ __forceinline__ void multimem_st_v4(void *mc_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(a)),
               "f"(__uint_as_float(b)), "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}

constexpr float kMagic = 12582912.f;          // 1.5 * 2^23: x + kMagic (round down) = floor(x) in the mantissa

// float64 path = the reference's arithmetic (pyviz/apap.py:182-183,211-215): float32 H^-1 promoted
// to float64, IEEE divide, strict bounds, truncation.  Returns the source pixel index or -1.
static __device__ __noinline__ int exact_lookup(const float *__restrict__ h, int x, int y, int src_w, int src_h) {
  const double xd = (double)x, yd = (double)y;
  const double t0 = (double)h[0] * xd + (double)h[1] * yd + (double)h[2];
  const double t1 = (double)h[3] * xd + (double)h[4] * yd + (double)h[5];
  const double t2 = (double)h[6] * xd + (double)h[7] * yd + (double)h[8];
  const double tx = t0 / t2;
  const double ty = t1 / t2;
  if (0.0 < tx && tx < (double)src_w && 0.0 < ty && ty < (double)src_h) return (int)ty * src_w + (int)tx;
  return -1;
}

// Per-lane state of the cell the strip is currently in.
struct CellState {
  float b0, b1, b2, m0, m1, m2, hme;
  int qbx, qby;                // integer base - bits(kMagic)
  int cell_row, cell;
  bool outside;                // every pixel of the cell maps outside the source
};

// Enter the cell row of a block: (re)load the lane's cell record when it changes (warp-uniform).
__device__ __forceinline__ void enter_cell_row(const WarpParams &p, const uint2 cl, float dxf, int cell_row,
                                               CellState &c) {
  if (cell_row != c.cell_row) {
    c.cell_row = cell_row;
    c.cell = cell_row * p.grid_cols + (int)cl.x;
    const float4 *rec = p.cell_fast + (size_t)c.cell * 3;
    const float4 u = __ldg(rec), v = __ldg(rec + 1), w = __ldg(rec + 2);
    c.m0 = fmaf(u.x, dxf, u.z); c.b0 = u.y;
    c.m1 = fmaf(u.w, dxf, v.y); c.b1 = v.x;
    c.m2 = fmaf(v.z, dxf, w.x); c.b2 = v.w;
    c.qbx = __float_as_int(w.y);
    c.qby = __float_as_int(w.z);
    // g = 0.5 - eps; a record with g < 0 (degenerate cell, forced) never passes; NaN never passes
    c.hme = p.force_exact ? -1.f : w.w;
    c.outside = w.w > 1.f && !p.force_exact;
  }
}

int launch_warp_tile(const WarpParams &w, bool words, const void *tiles, cudaStream_t st);
int launch_warp_tiles(const WarpParams &w, const int *col_ext, void *tiles, size_t tiles_bytes, cudaStream_t st);
bool warp_tile_usable(const WarpParams &w);
size_t warp_tiles_bytes(int canvas_w, int n_blocks);
int launch_warp_bilinear(const WarpParams &p, cudaStream_t st);

}  // namespace apap
