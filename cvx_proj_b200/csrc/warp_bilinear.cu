// Opt-in bilinear mode of the mesh warp (BASELINE.json north_star "bilinear sample"; SURVEY.md finding 5).
//
// The reference's APAP.local_warp (pyviz/apap.py:206-215) has no interpolation: it truncates the mapped coordinate
// (the parity mode, csrc/warp_tile.cu / warp_blend.cu, bit-exact).  This kernel keeps everything else of that loop --
// the cell lookup, t = H^-1 [x, y, 1] in float64 from the float32 grid, the division, the strict test
// 0 < tx < src_w, 0 < ty < src_h that decides WHICH canvas pixels are written -- and replaces the truncating gather by
// a bilinear sample with the convention of the reference's only bilinear sampler, cv.warpPerspective (pyviz/utils.py:114):
// pixel centres at integer coordinates, x0 = floor(tx), fx = tx - x0, taps x0 and x0 + 1 clamped to the image, value
// rounded half up.  Oracle: oracle/apap_oracle.py::local_warp_bilinear (float64 numpy); tolerance +-1 LSB (the weights
// here are float32).  A quality option, not a hot path: one thread per canvas pixel, float64 coordinates.
#include "common.cuh"
#include "warp_common.cuh"

namespace apap {

__global__ void __launch_bounds__(128) k_warp_bilinear(const WarpParams p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.canvas_w) return;
  const uint2 e = __ldg(p.row_blocks + blockIdx.y);
  const int i0 = (int)(e.x & 0x0fffffffu), n = (int)(e.x >> 28), cell_row = (int)(e.y & 0xffffu);
  const int cell = cell_row * p.grid_cols + (int)__ldg(p.col_lut + j).x;
  const float *h = p.cell_hinv + (size_t)cell * 9;
  const double h0 = h[0], h1 = h[1], h2 = h[2], h3 = h[3], h4 = h[4], h5 = h[5], h6 = h[6], h7 = h[7], h8 = h[8];
  const double xd = (double)(j - p.off_x);
  const uint32_t pitch = (uint32_t)p.canvas_w * 3u;
  for (int k = 0; k < n; ++k) {
    const int i = i0 + k;
    const double yd = (double)(i - p.off_y);
    const double t0 = h0 * xd + h1 * yd + h2, t1 = h3 * xd + h4 * yd + h5, t2 = h6 * xd + h7 * yd + h8;
    const double tx = t0 / t2, ty = t1 / t2;
    uint32_t b = 0, g = 0, r = 0;
    if (0.0 < tx && tx < (double)p.src_w && 0.0 < ty && ty < (double)p.src_h) {
      const double fx0 = floor(tx), fy0 = floor(ty);
      const int x0 = (int)fx0, y0 = (int)fy0;
      const int x1 = min(x0 + 1, p.src_w - 1), y1 = min(y0 + 1, p.src_h - 1);
      const float fx = (float)(tx - fx0), fy = (float)(ty - fy0);
      const uint8_t *r0 = p.src + (size_t)y0 * p.src_w * 3, *r1 = p.src + (size_t)y1 * p.src_w * 3;
      const uint8_t *p00 = r0 + x0 * 3, *p01 = r0 + x1 * 3, *p10 = r1 + x0 * 3, *p11 = r1 + x1 * 3;
      const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
      const float vb = w00 * p00[0] + w01 * p01[0] + w10 * p10[0] + w11 * p11[0];
      const float vg = w00 * p00[1] + w01 * p01[1] + w10 * p10[1] + w11 * p11[1];
      const float vr = w00 * p00[2] + w01 * p01[2] + w10 * p10[2] + w11 * p11[2];
      b = (uint32_t)min(255, (int)floorf(vb + 0.5f));
      g = (uint32_t)min(255, (int)floorf(vg + 0.5f));
      r = (uint32_t)min(255, (int)floorf(vr + 0.5f));
    }
    uint8_t *d = p.out + (size_t)(i - p.row0) * pitch + (size_t)j * 3;
    d[0] = (uint8_t)b; d[1] = (uint8_t)g; d[2] = (uint8_t)r;
  }
}

int launch_warp_bilinear(const WarpParams &p, cudaStream_t st) {
  if (p.n_blocks == 0) return 0;
  if (p.n_blocks > 65535) return fail(APAP_E_TOOBIG, "bilinear warp: more than 65535 row blocks in one launch (split the band)");
  k_warp_bilinear<<<dim3((p.canvas_w + 127) / 128, p.n_blocks), 128, 0, st>>>(p);
  return check_cuda(cudaGetLastError(), "k_warp_bilinear launch");
}

}  // namespace apap
