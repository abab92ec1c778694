// K3, tile engine -- mesh warp through shared-memory tiles (reference APAP.local_warp pixel loop,
// pyviz/apap.py:206-215), optionally fused with K4 (pyviz/apap_utils.py:75-88, pyviz/apap.py:259-261).
//
// Persistent, warp-specialised CTAs (8 worker warps + 1 producer warp) walk the canvas tiles
// (128 columns x 8 consecutive row blocks, <= 32 rows) t = blockIdx.x, blockIdx.x + gridDim.x, ...
//
//   producer warp, one tile ahead of the workers (two-stage ring, full / empty mbarriers):
//     footprint: lane = (cell column, row block) maps the four corners of that pixel rectangle of the
//     tile through the cell's H^-1 (float32 is enough: the result is widened by half a pixel).  Inside a
//     fast-path cell the maps are ratios of affine functions with a denominator of constant sign, so
//     every source pixel the tile can pick lies inside the bounding box of the corner images.  ONE 2-D
//     tensor-map TMA copy (cp.async.bulk.tensor.2d, SASS UTMALDG) stages that box in shared memory; the
//     part of the box outside the image arrives as zeros, which is exactly the reference's "leave
//     black", so no pixel needs a bounds test.  Three box shapes (wide, square-ish, tall) are encoded
//     per launch; the producer takes the first that holds the tile's box.  The tile's mode goes with it:
//       black   every cell of the tile maps outside the source: the workers only clear the output tile
//       staged  gather from the staged box
//       global  no box (it fits none of the shapes, source rows not 16-byte aligned, column LUT not
//               monotone, forced float64): same arithmetic with a bounds test, gathers from global memory
//   worker warps: 32 columns x 4 row blocks each; per row block the float32 fast path with the guard
//     band of csrc/warp_blend.cu (identical decisions), LDS.U8 gathers from the staged box, STS.U8 into
//     the output tile; guard-band pixels are re-decided in float64 and fetched from global memory
//     exactly like the reference;
//   the output tile (double buffered) leaves as whole row segments: TMA bulk stores (UBLKCP.G.S, four
//     rows per warp), 16-byte multimem.st for an NVLS multicast panorama, or plain word / byte stores for
//     unaligned canvases.
#include <cuda.h>
#include <limits.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "warp_common.cuh"

namespace apap {

#ifndef APAP_TILE_CTAS
#define APAP_TILE_CTAS 3
#endif
#ifndef APAP_TILE_BOX_KB
#define APAP_TILE_BOX_KB 20
#endif
constexpr int kWorkerWarps = 8;
constexpr int kWorkerThreads = kWorkerWarps * 32;
constexpr int kTileThreads = kWorkerThreads + 32;         // + the producer warp
constexpr int kTileCols = 128;                            // 4 chunks of 32 columns
constexpr int kTileBlocks = 8;                            // row blocks per tile
constexpr int kWarpBlocks = 4;                            // row blocks per worker warp (two warps per column chunk)
constexpr int kTileRows = kTileBlocks * kBlockRows;       // 32
constexpr int kOutPitch = kTileCols * 3;                  // 384 B per tile row
constexpr int kOutBytes = kTileRows * kOutPitch;          // 12 KB
constexpr int kBoxBytes = APAP_TILE_BOX_KB * 1024;        // staged source box
constexpr uint32_t kFlagOff = kBoxBytes + 4;              // staged mode, "guard band" marker: reads the zero word behind the box
constexpr uint32_t kNoPixel = 0xffffffffu;                // global mode: "leave black"
constexpr uint32_t kFlagPixel = 0xfffffffeu;              // global mode: guard band
constexpr int kMaxCellCols = 32;                          // footprint: at most this many cell columns per tile

enum TileMode : int { kBlack = 0, kStagedMode = 1, kGlobal = 3 };

// box shapes of the staged source: {bytes per row (multiple of 16, <= 1024), rows (<= 256)}, each <= kBoxBytes
constexpr int kBoxShapes = 3;
__host__ __device__ constexpr int box_w(int m) { return m == 0 ? 448 : m == 1 ? 256 : 128; }
__host__ __device__ constexpr int box_h(int m) { return kBoxBytes / box_w(m) > 256 ? 256 : kBoxBytes / box_w(m); }

constexpr int kRecRuns = 4;                               // cell rows of a tile whose records are staged ...
constexpr int kRecCols = 16;                              // ... for at most this many cell columns (else global loads)

struct TileInfo {                // what the producer tells the workers about a tile
  int x_lo, y_lo;                // first source pixel column / row of the staged box (kGlobal: 0, 0)
  int pitch;                     // bytes per staged row (kGlobal: pixels per image row)
  int shift;                     // byte offset of pixel x_lo inside a staged row
  int mode;
  int r_first, n_rows;           // canvas rows of the tile
  int c_lo;                      // first cell column of the tile
  int rec_ok;                    // the cell records of the tile are staged in Stage::rec
  int pad[3];
  uint2 blk[kTileBlocks];        // the tile's row block entries
  int slot[kTileBlocks];         // cell-row run of every block = first index into Stage::rec
};

struct Stage {
  alignas(128) uint8_t box[kBoxBytes];
  alignas(16) uint8_t zero[16];
  alignas(16) float4 rec[kRecRuns][kRecCols][3];          // fast-path records of the tile's cells
};

struct TileSmem {
  Stage st[2];
  alignas(128) uint8_t out[2][kOutBytes];
  alignas(16) TileInfo info[2];
  alignas(8) uint64_t full[2];
  uint64_t empty[2];
};

static_assert(offsetof(Stage, zero) == kBoxBytes, "the zero word sits right behind the staged box");
static_assert(box_w(0) * box_h(0) <= kBoxBytes && box_w(1) * box_h(1) <= kBoxBytes && box_w(2) * box_h(2) <= kBoxBytes, "box shapes");

struct TileParams {
  CUtensorMap maps[kBoxShapes];  // the source image as uint32 [src_h][src_w * 3 / 4], one map per box shape
  WarpParams w;
  const int2 *col_ext;         // [grid_cols] {first, last} canvas column of the cell column
  int src_tma_ok;              // source rows 16-byte aligned and the maps encoded: boxes can be staged
  int store_mode;              // 0 bytes, 1 words, 2 TMA bulk rows, 3 multimem
  int tiles_x, n_tiles;
  int step_x, step_y;          // gridDim.x = step_y * tiles_x + step_x: how a CTA's tile coordinates advance
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void worker_barrier() {          // the 8 worker warps only (named barrier 1)
  asm volatile("bar.sync 1, %0;" ::"n"(kWorkerThreads) : "memory");
}
// 2-D tensor-map TMA copy global -> shared (SASS: UTMALDG); coordinates in elements of the map, out-of-range parts
// of the box arrive as zeros
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_but_one() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// uniform_blend of one pixel packed as b0 | b1 << 8 | b2 << 16 with the centre image's pixel (0 = none)
__device__ __forceinline__ uint32_t blend_px(uint32_t val, uint32_t cv) {
  if (cv != 0) val = (val != 0) ? __vhaddu4(val, cv) : cv;
  return val;
}

__device__ __forceinline__ uint32_t centre_px(const WarpParams &p, int cx, int cy, bool col_ok) {
  if ((unsigned)cy < (unsigned)p.centre_h && (unsigned)cx < (unsigned)p.centre_w && col_ok) {
    const uint8_t *q = p.centre + ((size_t)cy * p.centre_w + cx) * 3;
    return (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
  }
  return 0u;
}

// One row block of one lane: gather offsets on the float32 fast path (same arithmetic and decisions as
// block_issue of csrc/warp_blend.cu), gathers, blend, bytes into the output tile; then the float64 path
// for the rows whose quotient lies inside the guard band.
//   kStaged: the shared-memory address of the pixel comes straight from the float bits of the floors
//   (address = ty_bits * pitch + tx_bits * 3 + kbase), no bounds test: whatever the tile can pick is inside
//   the staged box and the box is zero outside the image;  else: test against the image, gather from global memory.
template <bool kStaged, bool kFull, bool kBlend>
__device__ __forceinline__ void tile_block(const WarpParams &p, const CellState &c, int pitch, int kbase,
                                           uint32_t guard_addr, uint8_t *__restrict__ orow, float dy0, int n_rows, int x,
                                           int y0, bool col_ok, int cell) {
  constexpr int P = kBlockRows;
  uint32_t so[P];                              // kStaged: shared-memory address; else pixel index / kNoPixel / kFlagPixel
  const uint32_t guard = kStaged ? guard_addr : kFlagPixel;
  if (__all_sync(0xffffffffu, c.outside)) {    // every pixel of the warp's 32 x 4 block stays black
#pragma unroll
    for (int k = 0; k < (kFull ? P : n_rows); ++k) {
      uint32_t val = 0;
      if (kBlend) val = centre_px(p, x, y0 + k, col_ok);
      uint8_t *d = orow + k * kOutPitch;
      d[0] = (uint8_t)val; d[1] = (uint8_t)(val >> 8); d[2] = (uint8_t)(val >> 16);
    }
    return;
  }
  {
    // global mode, a lane whose cell maps outside the source (mixed warp): an x index beyond the image
    const int qx = (!kStaged && c.outside) ? 0x40000000 : c.qbx;
    const float hme = c.hme;                   // outside cells carry 2: their guard test always passes
    const float2 km = make_float2(kMagic, kMagic), nkm = make_float2(-kMagic, -kMagic), nh = make_float2(-0.5f, -0.5f);
#pragma unroll
    for (int k = 0; k < P; k += 2) {
      const float2 dy = __fadd2_rn(make_float2(dy0, dy0), make_float2((float)k, (float)(k + 1)));
      const float2 n0 = __ffma2_rn(make_float2(c.b0, c.b0), dy, make_float2(c.m0, c.m0));
      const float2 n1 = __ffma2_rn(make_float2(c.b1, c.b1), dy, make_float2(c.m1, c.m1));
      const float2 d = __ffma2_rn(make_float2(c.b2, c.b2), dy, make_float2(c.m2, c.m2));
      const float2 r = make_float2(rcp_approx(d.x), rcp_approx(d.y));
      const float2 qxq = __fmul2_rn(n0, r), qyq = __fmul2_rn(n1, r);
      const float2 tx = __fadd2_rd(qxq, km), ty = __fadd2_rd(qyq, km);            // floor + kMagic
      const float2 gx = __fadd2_rn(tx, nkm), gy = __fadd2_rn(ty, nkm);            // floor
      const float2 fx = __fadd2_rn(qxq, make_float2(-gx.x, -gx.y));               // exact fraction in [0, 1)
      const float2 fy = __fadd2_rn(qyq, make_float2(-gy.x, -gy.y));
      const float2 hx = __fadd2_rn(fx, nh), hy = __fadd2_rn(fy, nh);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float txe = e ? tx.y : tx.x, tye = e ? ty.y : ty.x;
        const float hxe = e ? hx.y : hx.x, hye = e ? hy.y : hy.x;
        const bool clear = fmaxf(fabsf(hxe), fabsf(hye)) <= hme;
        uint32_t off;
        if (kStaged) {
          off = (uint32_t)(__float_as_int(tye) * pitch + (__float_as_int(txe) * 3 + kbase));
        } else {
          const int ix = __float_as_int(txe) + qx;
          const int iy = __float_as_int(tye) + c.qby;
          const uint32_t in_off = (uint32_t)(iy * p.src_w + ix);
          asm("{\n\t.reg .pred p;\n\t"
              "setp.lt.u32 p, %1, %2;\n\t"
              "setp.lt.and.u32 p, %3, %4, p;\n\t"
              "selp.b32 %0, %5, %6, p;\n\t}"
              : "=r"(off)
              : "r"(ix), "r"(p.src_w), "r"(iy), "r"(p.src_h), "r"(in_off), "r"(kNoPixel));
        }
        so[k + e] = clear ? off : guard;
      }
    }
  }

  // gathers + stores of the rows of this block
#pragma unroll
  for (int k = 0; k < P; ++k) {
    if (!kFull && k >= n_rows) break;        // warp-uniform
    const uint32_t o = so[k];
    uint32_t b0, b1, b2;
    if (kStaged) {
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b0) : "r"(o));
      asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(b1) : "r"(o));
      asm volatile("ld.shared.u8 %0, [%1+2];" : "=r"(b2) : "r"(o));
    } else {
      b0 = b1 = b2 = 0;
      if (o < kFlagPixel) {
        const uint8_t *q = p.src + (size_t)o * 3;
        b0 = __ldg(q); b1 = __ldg(q + 1); b2 = __ldg(q + 2);
      }
    }
    if (kBlend) {
      const uint32_t val = blend_px(b0 | (b1 << 8) | (b2 << 16), centre_px(p, x, y0 + k, col_ok));
      b0 = val & 0xffu; b1 = (val >> 8) & 0xffu; b2 = val >> 16;
    }
    uint8_t *d = orow + k * kOutPitch;
    d[0] = (uint8_t)b0; d[1] = (uint8_t)b1; d[2] = (uint8_t)b2;
  }

  // rare: the reference's float64 arithmetic, bytes from global memory
  bool any_guard = so[0] == guard || so[1] == guard || so[2] == guard || so[3] == guard;
  if (__any_sync(0xffffffffu, any_guard)) {
    const float *hinv = p.cell_hinv + (size_t)cell * 9;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      if ((!kFull && k >= n_rows) || so[k] != guard) continue;
      const int idx = exact_lookup(hinv, x, y0 + k, p.src_w, p.src_h);
      uint32_t val = 0;
      if (idx >= 0) {
        const uint8_t *q = p.src + (size_t)(unsigned)idx * 3;
        val = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8) | ((uint32_t)__ldg(q + 2) << 16);
      }
      if (kBlend) val = blend_px(val, centre_px(p, x, y0 + k, col_ok));
      uint8_t *d = orow + k * kOutPitch;
      d[0] = (uint8_t)val; d[1] = (uint8_t)(val >> 8); d[2] = (uint8_t)(val >> 16);
    }
  }
}

// A worker warp's share of one tile: 32 columns x up to kWarpBlocks row blocks.  Block entries and (normally)
// the cells' fast-path records come from shared memory, where the producer staged them.
template <bool kStaged, bool kBlend>
__device__ __forceinline__ void tile_warp_work(const WarpParams &p, const TileInfo &f, const Stage &stg,
                                               uint8_t *__restrict__ out_tile, const uint2 cl, int x, bool col_ok,
                                               int lane_col, int b0, int nb) {
  const float dxf = __uint_as_float(cl.y);
  const uint32_t box_addr = smem_u32(stg.box);
  const int rec_col = (int)cl.x - f.c_lo;
  CellState c;
  c.b0 = c.b1 = c.b2 = c.m0 = c.m1 = 0.f; c.m2 = 1.f; c.hme = -1.f;
  c.qbx = c.qby = 0; c.cell_row = -1; c.cell = 0; c.outside = false;
  int kbase = 0;
#pragma unroll 1
  for (int bi = 0; bi < nb; ++bi) {
    const uint2 e = f.blk[b0 + bi];
    const int i0 = (int)(e.x & 0x0fffffffu), n = (int)(e.x >> 28);
    const int cell_row = (int)(e.y & 0xffffu);
    if (cell_row != c.cell_row) {               // warp-uniform: the strip enters a new cell row
      c.cell_row = cell_row;
      float4 u, v, w;
      if (f.rec_ok) {
        const float4 *rec = stg.rec[f.slot[b0 + bi]][rec_col];
        u = rec[0]; v = rec[1]; w = rec[2];
      } else {
        const float4 *rec = p.cell_fast + (size_t)(cell_row * p.grid_cols + (int)cl.x) * 3;
        u = __ldg(rec); v = __ldg(rec + 1); w = __ldg(rec + 2);
      }
      c.m0 = fmaf(u.x, dxf, u.z); c.b0 = u.y;
      c.m1 = fmaf(u.w, dxf, v.y); c.b1 = v.x;
      c.m2 = fmaf(v.z, dxf, w.x); c.b2 = v.w;
      c.qbx = __float_as_int(w.y);
      c.qby = __float_as_int(w.z);
      // g = 0.5 - eps; a record with g < 0 (degenerate cell, forced) never passes; NaN never passes
      c.hme = p.force_exact ? -1.f : w.w;
      c.outside = w.w > 1.f && !p.force_exact;
      // staged: address = box + (iy - y_lo) * pitch + (ix - x_lo) * 3 + shift, iy = ty_bits + qby, ix = tx_bits + qbx
      kbase = (int)box_addr + (c.qby - f.y_lo) * f.pitch + (c.qbx - f.x_lo) * 3 + f.shift;
    }
    uint8_t *orow = out_tile + (uint32_t)(i0 - f.r_first) * kOutPitch + lane_col * 3;
    const float dy0 = (float)(e.y >> 16);
    const int cell = cell_row * p.grid_cols + (int)cl.x;
    if (n == kBlockRows)
      tile_block<kStaged, true, kBlend>(p, c, f.pitch, kbase, box_addr + kFlagOff, orow, dy0, n, x, i0 - p.off_y, col_ok, cell);
    else
      tile_block<kStaged, false, kBlend>(p, c, f.pitch, kbase, box_addr + kFlagOff, orow, dy0, n, x, i0 - p.off_y, col_ok, cell);
  }
}

template <bool kBlend>
__global__ void __launch_bounds__(kTileThreads, APAP_TILE_CTAS) k_warp_tile(const __grid_constant__ TileParams tp) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  TileSmem &sm = *reinterpret_cast<TileSmem *>(smem_raw);
  const WarpParams &p = tp.w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&sm.full[0], 1); mbar_init(&sm.full[1], 1);
    mbar_init(&sm.empty[0], kWorkerWarps); mbar_init(&sm.empty[1], kWorkerWarps);
    mbar_fence_init();
  }
  if (tid < 8) reinterpret_cast<uint32_t *>(sm.st[tid >> 2].zero)[tid & 3] = 0u;
  __syncthreads();

  const int tiles_x = tp.tiles_x;
  int txi = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;  // coordinates of this CTA's current tile

  if (warp == kWorkerWarps) {
    // =============================== producer warp ==================================================
    const int b = lane & 7, ccs = lane >> 3;               // this lane's footprint items: row block b, cell columns ccs + 4 u
    for (int k = 0, tile = blockIdx.x; tile < tp.n_tiles; ++k, tile += gridDim.x) {
      const int s = k & 1;
      const int j0 = txi * kTileCols, jw = min(kTileCols, p.canvas_w - j0), j1 = j0 + jw - 1;
      const int tb0 = ty * kTileBlocks, nbt = min(kTileBlocks, p.n_blocks - tb0);
      txi += tp.step_x; ty += tp.step_y;
      if (txi >= tiles_x) { txi -= tiles_x; ++ty; }
      const int c_lo = (int)__ldg(p.col_lut + j0).x, c_hi = (int)__ldg(p.col_lut + j1).x;
      const uint2 e = __ldg(p.row_blocks + tb0 + min(b, nbt - 1));
      const int i0 = (int)(e.x & 0x0fffffffu), n = (int)(e.x >> 28), cr = (int)(e.y & 0xffffu);
      bool odd = c_hi - c_lo >= kMaxCellCols || c_hi < c_lo || !tp.src_tma_ok;
      // every column's cell inside [c_lo, c_hi]?  (always, for the monotone LUT of a sorted mesh)
#pragma unroll
      for (int q = 0; q < kTileCols / 32; ++q) {
        const int cc = (int)__ldg(p.col_lut + min(j0 + lane + 32 * q, j1)).x;
        odd = odd || cc < c_lo || cc > c_hi;
      }
      odd = __any_sync(0xffffffffu, odd);
      const bool bad_lut = odd && tp.src_tma_ok;              // (conservative) the cell columns are not the range c_lo .. c_hi
      float bx0 = 3e9f, bx1 = -3e9f, by0 = 3e9f, by1 = -3e9f;
      bool all_out = true;
      if (!odd) {
        const float ya = (float)(i0 - p.off_y), yb = (float)(i0 + n - 1 - p.off_y);
        const int trips = (c_hi - c_lo + 4) >> 2;            // warp-uniform
#pragma unroll 2
        for (int u = 0; u < trips; ++u) {
          const int cc = c_lo + ccs + 4 * u;
          const int ccl = min(cc, c_hi);                    // loads stay in range; the item counts only if cc <= c_hi
          const int2 ce = __ldg(tp.col_ext + ccl);
          const size_t cell = (size_t)(cr * p.grid_cols + ccl);
          const float g = __ldg(reinterpret_cast<const float *>(p.cell_fast) + cell * kHinvRow + 11);
          const float *h = p.cell_hinv + cell * 9;
          const float h0 = __ldg(h + 0), h1 = __ldg(h + 1), h2 = __ldg(h + 2), h3 = __ldg(h + 3), h4 = __ldg(h + 4);
          const float h5 = __ldg(h + 5), h6 = __ldg(h + 6), h7 = __ldg(h + 7), h8 = __ldg(h + 8);
          const int xa = max(ce.x, j0), xb = min(ce.y, j1);
          const bool item = cc <= c_hi && b < nbt && xa <= xb;
          const float xfa = (float)(xa - p.off_x), xfb = (float)(xb - p.off_x);
          float qx[4], qy[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float xf = (v & 1) ? xfb : xfa, yf = (v & 2) ? yb : ya;
            const float rt = rcp_approx(fmaf(h6, xf, fmaf(h7, yf, h8)));
            qx[v] = fmaf(h0, xf, fmaf(h1, yf, h2)) * rt;
            qy[v] = fmaf(h3, xf, fmaf(h4, yf, h5)) * rt;
          }
          const float lx = fminf(fminf(qx[0], qx[1]), fminf(qx[2], qx[3])), ux = fmaxf(fmaxf(qx[0], qx[1]), fmaxf(qx[2], qx[3]));
          const float ly = fminf(fminf(qy[0], qy[1]), fminf(qy[2], qy[3])), uy = fmaxf(fmaxf(qy[0], qy[1]), fmaxf(qy[2], qy[3]));
          // finite and small?  (fminf / fmaxf drop a NaN, so test every corner)
          bool fin = true;
#pragma unroll
          for (int v = 0; v < 4; ++v) fin = fin && fabsf(qx[v]) < 1e9f && fabsf(qy[v]) < 1e9f;
          if (item) {
            all_out = all_out && g > 1.f;
            if (fin) { bx0 = fminf(bx0, lx); bx1 = fmaxf(bx1, ux); by0 = fminf(by0, ly); by1 = fmaxf(by1, uy); }
            else odd = true;                                // cannot bound this rectangle's image
          }
        }
      }
      odd = __any_sync(0xffffffffu, odd);
      all_out = __all_sync(0xffffffffu, all_out) && !odd;
      // the pixel picked is floor(q): widen by half a pixel for the float32 evaluation, then floor
      const int ix0 = __reduce_min_sync(0xffffffffu, __float2int_rd(bx0 - 0.5f));
      const int ix1 = __reduce_max_sync(0xffffffffu, __float2int_rd(bx1 + 0.5f));
      const int iy0 = __reduce_min_sync(0xffffffffu, __float2int_rd(by0 - 0.5f));
      const int iy1 = __reduce_max_sync(0xffffffffu, __float2int_rd(by1 + 0.5f));

      // cell-row runs of the tile's blocks (lanes 0 .. nbt-1 hold block b = lane): run index = record slot
      const int cr_prev = __shfl_up_sync(0xffffffffu, cr, 1);
      const unsigned starts = __ballot_sync(0xffffffffu, lane < nbt && (lane == 0 || cr != cr_prev));
      const int slot = __popc(starts & ((2u << b) - 1u)) - 1;
      const int n_runs = __popc(starts), ncc = c_hi - c_lo + 1;
      const bool rec_ok = n_runs <= kRecRuns && ncc >= 1 && ncc <= kRecCols && !all_out && !bad_lut;
      const int run_first = __fns(starts, 0, min(lane, n_runs - 1) + 1);    // lane r: first block of run r
      const int run_cr = __shfl_sync(0xffffffffu, cr, run_first);

      int mode = kGlobal, x_lo = 0, y_lo = 0, bpitch = p.src_w, shift = 0, shape = -1, c0 = 0;
      if (!odd && all_out) {
        mode = kBlack;
      } else if (!odd && ix0 <= ix1 && iy0 <= iy1) {
        // the box starts at a 16-byte aligned byte of the row (TMA: innermost coordinate x element size % 16 == 0)
        const long long need_w = ((long long)ix1 - ix0 + 1) * 3 + 15, need_h = (long long)iy1 - iy0 + 1;
#pragma unroll
        for (int m = kBoxShapes - 1; m >= 0; --m)
          if (need_w <= box_w(m) && need_h <= box_h(m)) shape = m;
        if (shape >= 0) {
          c0 = ((ix0 * 3) >> 4) << 2;                       // floor to 16 bytes: first uint32 element of the box
          mode = kStagedMode;
          x_lo = ix0; y_lo = iy0;
          bpitch = shape == 0 ? box_w(0) : shape == 1 ? box_w(1) : box_w(2);
          shift = ix0 * 3 - (c0 << 2);
        }
      }
      const int r_first = __shfl_sync(0xffffffffu, i0, 0);
      const int r_last = __shfl_sync(0xffffffffu, i0 + n - 1, nbt - 1);

      mbar_wait(&sm.empty[s], ((k >> 1) & 1) ^ 1);           // the workers have left this stage
      TileInfo &f = sm.info[s];
      if (lane < kTileBlocks) { f.blk[lane] = e; f.slot[lane] = slot; }
      if (lane == 0) {
        f.x_lo = x_lo; f.y_lo = y_lo; f.pitch = bpitch; f.shift = shift; f.mode = mode;
        f.r_first = r_first; f.n_rows = r_last - r_first + 1; f.c_lo = c_lo; f.rec_ok = rec_ok ? 1 : 0;
      }
      __syncwarp();
      const uint32_t box_bytes = shape < 0 ? 0u : (uint32_t)(bpitch * (shape == 0 ? box_h(0) : shape == 1 ? box_h(1) : box_h(2)));
      const uint32_t rec_bytes = rec_ok ? (uint32_t)(ncc * kHinvRow * 4) : 0u;
      if (lane == 0) {
        if (box_bytes + rec_bytes) mbar_arrive_expect_tx(&sm.full[s], box_bytes + rec_bytes * n_runs);
        else mbar_arrive(&sm.full[s]);
        if (shape >= 0) tma_load_2d(sm.st[s].box, &tp.maps[shape], c0, iy0, &sm.full[s]);
      }
      __syncwarp();
      if (rec_ok && lane < n_runs)                           // one bulk copy per cell row: the records of cells c_lo .. c_hi
        bulk_g2s(&sm.st[s].rec[lane][0][0], p.cell_fast + (size_t)(run_cr * p.grid_cols + c_lo) * 3, rec_bytes, &sm.full[s]);
    }
    return;
  }

  // ================================= worker warps ===================================================
  const int lane_col = (warp & 3) * 32 + lane;
  const int half = warp >> 2;
  const uint32_t pitch = (uint32_t)p.canvas_w * 3u;
  uint2 cl_next = __ldg(p.col_lut + min(txi * kTileCols + lane_col, p.canvas_w - 1));
  for (int k = 0, tile = blockIdx.x; tile < tp.n_tiles; ++k, tile += gridDim.x) {
    const int s = k & 1;
    const int j0 = txi * kTileCols, jw = min(kTileCols, p.canvas_w - j0);
    const int nbt = min(kTileBlocks, p.n_blocks - ty * kTileBlocks);
    txi += tp.step_x; ty += tp.step_y;
    if (txi >= tiles_x) { txi -= tiles_x; ++ty; }
    const int j = j0 + lane_col;
    const bool col_ok = j < p.canvas_w;
    const uint2 cl = cl_next;
    const int x = j - p.off_x;
    const int nb = max(0, min(kWarpBlocks, nbt - half * kWarpBlocks));
    uint8_t *out_tile = sm.out[s];

    if (tp.store_mode == 2) bulk_wait_read_but_one();        // this thread's row store of two tiles ago has read out[s]
    mbar_wait(&sm.full[s], (k >> 1) & 1);                    // box, records and info of this tile have landed
    if (tile + (int)gridDim.x < tp.n_tiles)                  // the next tile's column entry: in flight during this tile
      cl_next = __ldg(p.col_lut + min(txi * kTileCols + lane_col, p.canvas_w - 1));
    const TileInfo &f = sm.info[s];
    const int mode = f.mode, r_first = f.r_first, n_rows = f.n_rows;
    if (mode == kStagedMode) {
      tile_warp_work<true, kBlend>(p, f, sm.st[s], out_tile, cl, x, col_ok, lane_col, half * kWarpBlocks, nb);
    } else if (mode == kGlobal) {
      tile_warp_work<false, kBlend>(p, f, sm.st[s], out_tile, cl, x, col_ok, lane_col, half * kWarpBlocks, nb);
    } else {                                                  // kBlack: clear (or fill with the centre image) the tile
      if (!kBlend) {
        for (int q = tid; q < n_rows * (kOutPitch / 16); q += kWorkerThreads)
          reinterpret_cast<uint4 *>(out_tile)[q] = make_uint4(0u, 0u, 0u, 0u);
      } else {
        for (int bi = 0; bi < nb; ++bi) {
          const uint2 e = f.blk[half * kWarpBlocks + bi];
          const int i0 = (int)(e.x & 0x0fffffffu), n = (int)(e.x >> 28);
          for (int r = 0; r < n; ++r) {
            const uint32_t val = centre_px(p, x, i0 + r - p.off_y, col_ok);
            uint8_t *d = out_tile + (uint32_t)(i0 - r_first + r) * kOutPitch + lane_col * 3;
            d[0] = (uint8_t)val; d[1] = (uint8_t)(val >> 8); d[2] = (uint8_t)(val >> 16);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.empty[s]);                // this warp no longer reads the stage
    if (tp.store_mode == 2) fence_proxy_async_smem();        // this thread's tile bytes -> visible to the bulk stores
    worker_barrier();                                        // the output tile is complete
    const uint32_t row_bytes = (uint32_t)jw * 3u;
    uint8_t *gout = p.out + ((size_t)(r_first - p.row0) * pitch + (size_t)j0 * 3);
    if (tp.store_mode == 2) {                                // every warp stores four rows: lanes 0-3, one bulk copy each
      const int r = warp * (kTileRows / kWorkerWarps) + lane;
      if (lane < kTileRows / kWorkerWarps && r < n_rows)
        bulk_s2g(gout + (size_t)r * pitch, out_tile + r * kOutPitch, row_bytes);
      bulk_commit();
    } else if (tp.store_mode == 3) {
      const int per_row = (int)(row_bytes >> 4);
      for (int q = tid; q < n_rows * per_row; q += kWorkerThreads) {
        const int r = q / per_row, u = q - r * per_row;
        const uint4 v = *reinterpret_cast<const uint4 *>(out_tile + r * kOutPitch + u * 16);
        multimem_st_v4(gout + (size_t)r * pitch + u * 16, v.x, v.y, v.z, v.w);
      }
    } else if (tp.store_mode == 1) {
      const int per_row = (int)(row_bytes >> 2);
      for (int q = tid; q < n_rows * per_row; q += kWorkerThreads) {
        const int r = q / per_row, u = q - r * per_row;
        *reinterpret_cast<uint32_t *>(gout + (size_t)r * pitch + u * 4) =
            *reinterpret_cast<const uint32_t *>(out_tile + r * kOutPitch + u * 4);
      }
    } else {
      for (int q = tid; q < n_rows * (int)row_bytes; q += kWorkerThreads) {
        const int r = q / (int)row_bytes, u = q - r * (int)row_bytes;
        gout[(size_t)r * pitch + u] = out_tile[r * kOutPitch + u];
      }
    }
  }
  if (tp.store_mode == 2) bulk_wait_read_all();               // shared memory must outlive the last stores' reads
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

int launch_warp_tile(const WarpParams &w, const int *col_ext, bool words, cudaStream_t st) {
  TileParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.w = w;
  tp.col_ext = reinterpret_cast<const int2 *>(col_ext);
  const size_t src_pitch = (size_t)w.src_w * 3;
  tp.src_tma_ok = (src_pitch % 16 == 0) && !(reinterpret_cast<uintptr_t>(w.src) & 15u) && !w.force_exact;
  if (getenv("APAP_TILE_NO_TMA")) tp.src_tma_ok = 0;      // lab switch: every tile gathers from global memory
  if (tp.src_tma_ok) {                                     // the source as uint32 [src_h][src_w * 3 / 4], three box shapes
    EncodeTiledFn enc = encode_tiled_fn();
    const cuuint64_t dims[2] = {(cuuint64_t)(src_pitch / 4), (cuuint64_t)w.src_h};
    const cuuint64_t strides[1] = {(cuuint64_t)src_pitch};
    const cuuint32_t ones[2] = {1, 1};
    for (int m = 0; m < kBoxShapes && tp.src_tma_ok; ++m) {
      const cuuint32_t box[2] = {(cuuint32_t)(box_w(m) / 4), (cuuint32_t)box_h(m)};
      if (!enc || enc(&tp.maps[m], CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t *>(w.src), dims, strides, box, ones,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        tp.src_tma_ok = 0;                                 // no tensor maps: every tile gathers from global memory
    }
  }
  const bool rows16 = (w.canvas_w % 16 == 0) && !(reinterpret_cast<uintptr_t>(w.out) & 15u);
  tp.store_mode = w.multicast ? 3 : rows16 ? 2 : words ? 1 : 0;
  tp.tiles_x = (w.canvas_w + kTileCols - 1) / kTileCols;
  const long long tiles_y = (w.n_blocks + kTileBlocks - 1) / kTileBlocks;
  if (tiles_y * tp.tiles_x > 2147483647LL) return fail(APAP_E_TOOBIG, "warp: too many tiles in one launch (split the band)");
  tp.n_tiles = (int)(tiles_y * tp.tiles_x);
  static bool configured[64] = {};                       // > 48 KB of dynamic shared memory is opt-in, per device
  const size_t smem = sizeof(TileSmem);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice");
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    e = cudaFuncSetAttribute(k_warp_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k_warp_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "k_warp_tile shared memory opt-in");
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int grid = min(tp.n_tiles, sm_count_cached() * APAP_TILE_CTAS);
  tp.step_x = grid % tp.tiles_x;
  tp.step_y = grid / tp.tiles_x;
  if (w.centre) k_warp_tile<true><<<grid, kTileThreads, smem, st>>>(tp);
  else k_warp_tile<false><<<grid, kTileThreads, smem, st>>>(tp);
  return check_cuda(cudaGetLastError(), "k_warp_tile launch");
}

}  // namespace apap
