// Input preparation that used to run in numpy on the host and dominated the end-to-end time of the
// public calls (measured at c2: ~6 ms of table building and 10 ms of per-cell numpy inverses against a 33 us warp
// kernel):
//   k_kp_rows     conditioned keypoint pairs -> keypoint row table  (host restatement: apap.build_kp_table)
//   k_kp_blocks   keypoint row table -> tensor-core block table (host restatement: apap.build_kp_blocks)
//   k_warp_prep   per-cell fast-path records of the mesh warp   (host restatement: apap.build_warp_tables)
//   k_inv_grid    the per-cell inverse of local_warp, accepted only where it provably rounds like numpy's
// The first three reproduce their numpy restatements bit for bit (explicit _rn arithmetic, no FMA
// contraction, same operation order), so the CPU tests of the guard band cover the device-built tables.
#include <math.h>

#include "common.cuh"

namespace apap {

// ------------------------------------------------------------------------------------ k_kp_rows
// One thread per keypoint row.  apap.build_kp_table: with (x, y) the conditioned source point, (x', y') the
// conditioned target, m = [xx, xy, x, yy, y, 1] in float64 (exact products of float32 inputs), the row is
// float32 of [m | x' m | y' m | (x'x' + y'y') m], then the raw source point times `scale`, each coordinate twice.
// Rows at and past the scene's keypoint count are zero.
__global__ void __launch_bounds__(128) k_kp_rows(const float2 *__restrict__ cf1, const float2 *__restrict__ cf2,
                                                  const float2 *__restrict__ src, const int *__restrict__ counts,
                                                  int n_points, int n_pad, double scale, float *__restrict__ rows) {
  const int scene = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const int n = counts ? min(counts[scene], n_points) : n_points;
  float4 *dst = reinterpret_cast<float4 *>(rows + ((size_t)scene * n_pad + i) * kRowFloats);
  float v[kRowFloats];
  if (i < n) {
    const size_t at = (size_t)scene * n_points + i;
    const float2 a = cf1[at], b = cf2[at], k = src[at];
    const double x = a.x, y = a.y, xp = b.x, yp = b.y;
    const double m[6] = {__dmul_rn(x, x), __dmul_rn(x, y), x, __dmul_rn(y, y), y, 1.0};
    const double r = __dadd_rn(__dmul_rn(xp, xp), __dmul_rn(yp, yp));
#pragma unroll
    for (int t = 0; t < 6; ++t) {
      v[t] = __double2float_rn(m[t]);
      v[6 + t] = __double2float_rn(__dmul_rn(xp, m[t]));
      v[12 + t] = __double2float_rn(__dmul_rn(yp, m[t]));
      v[18 + t] = __double2float_rn(__dmul_rn(r, m[t]));
    }
    v[24] = v[25] = __double2float_rn(__dmul_rn((double)k.x, scale));
    v[26] = v[27] = __double2float_rn(__dmul_rn((double)k.y, scale));
  } else {
#pragma unroll
    for (int t = 0; t < kRowFloats; ++t) v[t] = 0.f;
  }
#pragma unroll
  for (int t = 0; t < kRowFloats / 4; ++t) dst[t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
}

int launch_kp_rows(const float *cf1, const float *cf2, const float *src, const int *counts, int batch, int n_points,
                   int n_kp_padded, double scale, float *rows, cudaStream_t st) {
  if (batch == 0 || n_kp_padded == 0) return 0;
  dim3 grid((n_kp_padded + 127) / 128, batch);
  k_kp_rows<<<grid, 128, 0, st>>>(reinterpret_cast<const float2 *>(cf1), reinterpret_cast<const float2 *>(cf2),
                                  reinterpret_cast<const float2 *>(src), counts, n_points, n_kp_padded, scale, rows);
  return check_cuda(cudaGetLastError(), "k_kp_rows launch");
}

// ------------------------------------------------------------------------------------ k_condition
// The O(N) prologue of local_homography (pyviz/apap.py:129-141) for one point set of one scene per CTA:
// Hartley normaliser (centroid, mean distance -> t, pyviz/apap.py:35-59), the conditioner of the normalised points
// (per-axis mean and unbiased standard deviation, pyviz/apap.py:63-89) and the conditioned points (pyviz/apap.py:92-100).
// The reference reduces in float32 with numpy's pairwise order; here every reduction is float64 in a fixed order
// (deterministic, and closer to the exact value), t and the conditioner are rounded to float32 like the reference's
// arrays, and the per-point arithmetic is the reference's float32 arithmetic.  H is invariant to the normalisation in
// exact arithmetic; the end-to-end difference against the host path is ~1e-6 of the 1e-4 gate (tests).
// The public static methods of the Python class stay on the host, bit-exact with the reference.
constexpr int kCondThreads = 512;

__device__ __forceinline__ double block_sum(double v, double *scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                                   // scratch may still be read from the previous sum
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double total = 0.0;
  for (int w = 0; w < kCondThreads / 32; ++w) total += scratch[w];   // same order in every thread
  return total;
}

// grid = (2, batch): blockIdx.x = 0 source points, 1 target points.  out_cond [2][batch][n_points][2] float32,
// out_mats [batch][2][2][9] float32 = per scene and set {t (normaliser), conditioner}.
__global__ void __launch_bounds__(kCondThreads) k_condition(const float2 *__restrict__ src, const float2 *__restrict__ dst,
                                                            const int *__restrict__ counts, int n_points, int batch,
                                                            float2 *__restrict__ out_cond, float *__restrict__ out_mats) {
  __shared__ double scratch[kCondThreads / 32];
  const int which = blockIdx.x, scene = blockIdx.y, tid = threadIdx.x;
  const int n = counts ? min(counts[scene], n_points) : n_points;
  const float2 *pts = (which ? dst : src) + (size_t)scene * n_points;
  float2 *cond = out_cond + ((size_t)which * batch + scene) * n_points;
  // centroid
  double sx = 0.0, sy = 0.0;
  for (int i = tid; i < n; i += kCondThreads) { const float2 p = pts[i]; sx += (double)p.x; sy += (double)p.y; }
  const double cx = block_sum(sx, scratch) / (double)n, cy = block_sum(sy, scratch) / (double)n;
  // mean distance to the centroid -> scale
  double sd = 0.0;
  for (int i = tid; i < n; i += kCondThreads) {
    const float2 p = pts[i];
    const double dx = (double)p.x - cx, dy = (double)p.y - cy;
    sd += sqrt(dx * dx + dy * dy);
  }
  const double scale = 1.4142135623730951 / (block_sum(sd, scratch) / (double)n + 1e-8);
  const float t00 = (float)scale, t02 = (float)(-scale * cx), t12 = (float)(-scale * cy);     // t is a float32 array
  // normalised points nf = t [p; 1] (float32), their mean and variance per axis
  double mx = 0.0, my = 0.0;
  for (int i = tid; i < n; i += kCondThreads) {
    const float2 p = pts[i];
    const float nx = __fadd_rn(__fmul_rn(t00, p.x), t02), ny = __fadd_rn(__fmul_rn(t00, p.y), t12);
    mx += (double)nx; my += (double)ny;
  }
  const double mux = block_sum(mx, scratch) / (double)n, muy = block_sum(my, scratch) / (double)n;
  double vx = 0.0, vy = 0.0;
  for (int i = tid; i < n; i += kCondThreads) {
    const float2 p = pts[i];
    const float nx = __fadd_rn(__fmul_rn(t00, p.x), t02), ny = __fadd_rn(__fmul_rn(t00, p.y), t12);
    const double ex = (double)nx - mux, ey = (double)ny - muy;
    vx += ex * ex; vy += ey * ey;
  }
  // unbiased standard deviation (std^2 * n / (n - 1), pyviz/apap.py:76-77), zero-deviation guard (:81-82)
  double devx = sqrt(block_sum(vx, scratch) / (double)n * (double)n / (double)(n - 1));
  double devy = sqrt(block_sum(vy, scratch) / (double)n * (double)n / (double)(n - 1));
  devx = devx + (devx == 0.0 ? 1.0 : 0.0);
  devy = devy + (devy == 0.0 ? 1.0 : 0.0);
  const double kx = 1.4142135623730951 / devx, ky = 1.4142135623730951 / devy;
  const float c00 = (float)kx, c02 = (float)(-kx * mux), c11 = (float)ky, c12 = (float)(-ky * muy);   // float32 array
  // conditioned points: cf = nf * diag + translation, the reference's two float32 roundings (pyviz/apap.py:96-99)
  for (int i = tid; i < n_points; i += kCondThreads) {
    float2 o = make_float2(0.f, 0.f);
    if (i < n) {
      const float2 p = pts[i];
      const float nx = __fadd_rn(__fmul_rn(t00, p.x), t02), ny = __fadd_rn(__fmul_rn(t00, p.y), t12);
      o.x = __fadd_rn(__fmul_rn(nx, c00), c02);
      o.y = __fadd_rn(__fmul_rn(ny, c11), c12);
    }
    cond[i] = o;
  }
  if (tid == 0) {
    float *m = out_mats + ((size_t)scene * 2 + which) * 18;
    const float t[9] = {t00, 0.f, t02, 0.f, t00, t12, 0.f, 0.f, 1.f};
    const float c[9] = {c00, 0.f, c02, 0.f, c11, c12, 0.f, 0.f, 1.f};
    for (int k = 0; k < 9; ++k) { m[k] = t[k]; m[9 + k] = c[k]; }
  }
}

// T2inv = inv(N2) inv(C2), T1 = C1 N1 (pyviz/apap.py:165-166) from the float32 matrices of k_condition: the
// inverses of these [[a, 0, b], [0, c, d], [0, 0, 1]] matrices in closed form, rounded to float32 like
// np.linalg.inv of a float32 matrix, products in float64.  One thread per scene.
__global__ void k_condition_mats(const float *__restrict__ mats, int batch, double *__restrict__ tmats) {
  const int scene = blockIdx.x * blockDim.x + threadIdx.x;
  if (scene >= batch) return;
  const float *n1 = mats + (size_t)scene * 36, *c1 = n1 + 9, *n2 = n1 + 18, *c2 = n1 + 27;
  auto inv = [](const float *a, double *o) {          // a = [[a0, 0, a2], [0, a4, a5], [0, 0, 1]]
    const double i0 = 1.0 / (double)a[0], i4 = 1.0 / (double)a[4];
    const double r[9] = {i0, 0.0, -(double)a[2] * i0, 0.0, i4, -(double)a[5] * i4, 0.0, 0.0, 1.0};
    for (int k = 0; k < 9; ++k) o[k] = (double)(float)r[k];
  };
  auto mul = [](const double *a, const double *b, double *o) {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  };
  double n2i[9], c2i[9], c1d[9], n1d[9];
  inv(n2, n2i); inv(c2, c2i);
  for (int k = 0; k < 9; ++k) { c1d[k] = (double)c1[k]; n1d[k] = (double)n1[k]; }
  mul(n2i, c2i, tmats + (size_t)scene * 18);
  mul(c1d, n1d, tmats + (size_t)scene * 18 + 9);
}

int launch_condition(const float *src, const float *dst, const int *counts, int batch, int n_points, float *cond,
                     float *mats, double *tmats, cudaStream_t st) {
  if (batch == 0 || n_points == 0) return 0;
  k_condition<<<dim3(2, batch), kCondThreads, 0, st>>>(reinterpret_cast<const float2 *>(src), reinterpret_cast<const float2 *>(dst),
                                                      counts, n_points, batch, reinterpret_cast<float2 *>(cond), mats);
  k_condition_mats<<<(batch + 63) / 64, 64, 0, st>>>(mats, batch, tmats);
  return check_cuda(cudaGetLastError(), "k_condition launch");
}

// ------------------------------------------------------------------------------------ k_kp_blocks
// Block layout (include/apap_b200.h): per 8 keypoints the 8 x 64 tile [Ph | Pl] in the K-major
// core-matrix layout -- element (k, column m) at float (k/4)*256 + (m/8)*32 + (m%8)*4 + k%4 -- then
// s*kx[8], s*ky[8].  Ph = P rounded to TF32 (ties away, like cvt.rna.tf32.f32), Pl = P - Ph (exact).
__global__ void __launch_bounds__(256) k_kp_blocks(const float *__restrict__ table, int n_kb, float *__restrict__ out) {
  const int kb = blockIdx.x;
  const float *rows = table + (size_t)kb * APAP_KP_BLOCK * kRowFloats;
  float *dst = out + (size_t)kb * APAP_KP_BLOCK_FLOATS;
  for (int o = threadIdx.x; o < APAP_KP_BLOCK_FLOATS; o += blockDim.x) {
    float v;
    if (o < 512) {
      const int j = o >> 8, r1 = (o >> 5) & 7, r0 = (o >> 2) & 7, kk = o & 3;
      const int k = j * 4 + kk, m = r1 * 8 + r0, n = m & 31;
      const float p = n < kTerms ? rows[k * kRowFloats + n] : 0.f;
      const float hi = __uint_as_float((__float_as_uint(p) + 0x1000u) & 0xFFFFE000u);
      v = m < 32 ? hi : __fsub_rn(p, hi);
    } else if (o < 520) {
      v = rows[(o - 512) * kRowFloats + 24];
    } else {
      v = rows[(o - 520) * kRowFloats + 26];
    }
    dst[o] = v;
  }
}

int launch_kp_blocks(const float *table, int batch, int n_kp_padded, float *out, cudaStream_t st) {
  const int n_kb = batch * (n_kp_padded / APAP_KP_BLOCK);
  if (n_kb == 0) return 0;
  k_kp_blocks<<<n_kb, 256, 0, st>>>(table, n_kb, out);
  return check_cuda(cudaGetLastError(), "k_kp_blocks launch");
}

// ------------------------------------------------------------------------------------ k_warp_prep
// One thread per cell.  See apap.build_warp_tables for the derivation: the cell's H^-1 rewritten
// relative to the cell's first pixel and an integer base, scaled so the denominator is ~1, plus a
// rigorous bound eps on the float32 quotient's error inside the cell.
constexpr double kU = 5.9604644775390625e-08;          // 2^-24, float32 unit roundoff
constexpr int kMagicBits = 0x4B400000;                 // float32 bits of 1.5 * 2^23

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double lin(double a, double x, double b, double y, double c) {   // (a x + b y) + c
  return add(add(mul(a, x), mul(b, y)), c);
}

__global__ void __launch_bounds__(128) k_warp_prep(const float *__restrict__ inv_h, const int2 *__restrict__ col_ext,
                                                    const int2 *__restrict__ row_ext, int grid_rows, int grid_cols,
                                                    int off_x, int off_y, int src_w, int src_h,
                                                    float *__restrict__ rec_out) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= grid_rows * grid_cols) return;
  const int r = cell / grid_cols, c = cell - r * grid_cols;
  int2 ce = col_ext[c], re = row_ext[r];               // {first, last} canvas column / row of the cell; first > last = unused
  const bool col_used = ce.x <= ce.y, row_used = re.x <= re.y;
  if (!col_used) ce = make_int2(0, 0);
  if (!row_used) re = make_int2(0, 0);
  double h[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) h[k] = (double)inv_h[(size_t)cell * 9 + k];
  const double x0 = (double)(ce.x - off_x), y0 = (double)(re.x - off_y);
  const double dxm = (double)(ce.y - ce.x), dym = (double)(re.y - re.x);
  bool ok = col_used && row_used;

  const double t0 = lin(h[0], x0, h[1], y0, h[2]);
  const double t1 = lin(h[3], x0, h[4], y0, h[5]);
  const double t2 = lin(h[6], x0, h[7], y0, h[8]);
  const double hx = mul(0.5, dxm), hy = mul(0.5, dym);
  // integer base: the source position of the cell centre
  const double c0 = add(add(t0, mul(h[0], hx)), mul(h[1], hy));
  const double c1 = add(add(t1, mul(h[3], hx)), mul(h[4], hy));
  const double c2 = add(add(t2, mul(h[6], hx)), mul(h[7], hy));
  double bx = rint(dvd(c0, c2)), by = rint(dvd(c1, c2));
  ok = ok && isfinite(bx) && isfinite(by) && fabs(bx) < 1073741824.0 && fabs(by) < 1073741824.0 && c2 != 0.0;
  if (!ok) bx = by = 0.0;
  const double s = ok ? dvd(1.0, c2) : 0.0;
  double coef[9];
  coef[0] = mul(sub(h[0], mul(bx, h[6])), s);
  coef[1] = mul(sub(h[1], mul(bx, h[7])), s);
  coef[2] = mul(sub(t0, mul(bx, t2)), s);
  coef[3] = mul(sub(h[3], mul(by, h[6])), s);
  coef[4] = mul(sub(h[4], mul(by, h[7])), s);
  coef[5] = mul(sub(t1, mul(by, t2)), s);
  coef[6] = mul(h[6], s);
  coef[7] = mul(h[7], s);
  coef[8] = mul(t2, s);
  bool fin = true;
#pragma unroll
  for (int k = 0; k < 9; ++k) fin = fin && isfinite(coef[k]);
  ok = ok && fin;
#pragma unroll
  for (int k = 0; k < 9; ++k) coef[k] = ok ? (double)(float)coef[k] : 0.0;   // what the warp kernel sees
  double m[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) m[k] = lin(fabs(coef[3 * k]), dxm, fabs(coef[3 * k + 1]), dym, fabs(coef[3 * k + 2]));
  // corners of the cell's pixel rectangle, in the order (0,0) (0,dym) (dxm,0) (dxm,dym)
  const double cx[4] = {mul(0.0, dxm), mul(0.0, dxm), dxm, dxm};
  const double cy[4] = {mul(0.0, dym), dym, mul(0.0, dym), dym};
  double corner[4], cmin = 0.0;
  bool same_sign = true;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    corner[i] = lin(coef[6], cx[i], coef[7], cy[i], coef[8]);
    same_sign = same_sign && corner[i] > 0.0;
    cmin = i == 0 ? corner[0] : fmin(cmin, corner[i]);
  }
  const double d_err = mul(3.0 * kU, m[2]);
  const double d_min = sub(cmin, d_err);               // the computed denominator is at least this
  ok = ok && same_sign && d_min >= 0.25;
  const double d_safe = ok ? d_min : 1.0;
  double eps = 0.0;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const double q = dvd(m[k], d_safe);
    const double e = add(dvd(add(mul(3.0 * kU, m[k]), mul(q, d_err)), d_safe), mul(q, 2.384185791015625e-07 + kU));
    ok = ok && isfinite(q) && q < 1048576.0;
    eps = fmax(eps, isfinite(e) ? e : 1.0);
  }
  eps = add(mul(1.25, eps), 1e-7);
  ok = ok && eps < 0.25;
  // cells that map entirely outside the source image: decided at the corners (the maps are ratios of
  // functions affine in (dx, dy) with a positive denominator)
  bool xl = true, xh = true, yl = true, yh = true;
  const double wlim = (double)src_w + 0.5, hlim = (double)src_h + 0.5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double dc = ok ? corner[i] : 1.0;
    const double qx = add(dvd(lin(coef[0], cx[i], coef[1], cy[i], coef[2]), dc), bx);
    const double qy = add(dvd(lin(coef[3], cx[i], coef[4], cy[i], coef[5]), dc), by);
    xl = xl && qx < -0.5; xh = xh && qx > wlim;
    yl = yl && qy < -0.5; yh = yh && qy > hlim;
  }
  const bool outside = ok && (xl || xh || yl || yh);

  float rec[kHinvRow];
#pragma unroll
  for (int k = 0; k < 9; ++k) rec[k] = ok ? (float)coef[k] : 0.f;
  if (!ok) rec[8] = 1.f;
  rec[9] = __int_as_float((int)((long long)(ok ? bx : 0.0) - (long long)kMagicBits));
  rec[10] = __int_as_float((int)((long long)(ok ? by : 0.0) - (long long)kMagicBits));
  const double hme = ok ? sub(0.5, eps) : -1.0;
  float hme32 = (float)hme;
  if ((double)hme32 > hme) hme32 = nextafterf(hme32, -2.f);   // round down
  rec[11] = outside ? 2.f : hme32;                            // 2 = every pixel of the cell is left black
  float4 *dst = reinterpret_cast<float4 *>(rec_out + (size_t)cell * kHinvRow);
  dst[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
  dst[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
  dst[2] = make_float4(rec[8], rec[9], rec[10], rec[11]);
}

int launch_warp_prep(const float *inv_h, const int *col_ext, const int *row_ext, int grid_rows, int grid_cols, int off_x,
                     int off_y, int src_w, int src_h, float *rec_out, cudaStream_t st) {
  const long long cells = (long long)grid_rows * grid_cols;
  if (cells == 0) return 0;
  if (cells > 2147483647LL / 12) return fail(APAP_E_TOOBIG, "warp tables: too many cells");
  k_warp_prep<<<(unsigned)((cells + 127) / 128), 128, 0, st>>>(inv_h, reinterpret_cast<const int2 *>(col_ext),
                                                               reinterpret_cast<const int2 *>(row_ext), grid_rows,
                                                               grid_cols, off_x, off_y, src_w, src_h, rec_out);
  return check_cuda(cudaGetLastError(), "k_warp_prep launch");
}

// ------------------------------------------------------------------------------------ k_inv_grid
// The per-cell inverse of local_warp (pyviz/apap.py:201-203).  The reference's float32 result is numpy's:
// LAPACK dgesv on the float64 promotion (Gaussian elimination with partial pivoting), rounded to float32 once.
// One thread per cell does the same factorisation in float64 and rounds; because the two float64 results may differ
// in their last bits (operation order, FMA), the cell is only accepted when every entry is provably far from a
// float32 rounding boundary: both results lie within  E = c u |X| (P^T |L||U|) |X|  of the exact inverse (the
// componentwise forward error of a GEPP solve, Higham, Accuracy and Stability of Numerical Algorithms, Thm 9.4,
// with c = 64 in place of 3n = 9), so if [x - 2E, x + 2E] contains no midpoint between adjacent float32 numbers,
// round(x) = round(dgesv's x).  Everything else -- an entry near a boundary, a (near-)tie in the pivot search (the
// other factorisation could pivot differently), zero / tiny / huge / non-finite entries, singular cells -- is
// flagged and left to the host, which calls numpy on those cells (~1 in 10^4 on homography grids).
__global__ void __launch_bounds__(128) k_inv_grid(const float *__restrict__ h, int cells, float *__restrict__ out,
                                                   unsigned char *__restrict__ flags) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= cells) return;
  double lu[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) lu[i][j] = (double)h[(size_t)cell * 9 + i * 3 + j];
  int perm[3] = {0, 1, 2};
  bool flag = false;
  constexpr double kTie = 1.0 - 1e-9;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    int piv = k;
    double best = fabs(lu[k][k]);
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      const double v = fabs(lu[i][k]);
      if (v > best) { best = v; piv = i; }
    }
#pragma unroll
    for (int i = k; i < 3; ++i)
      if (i != piv && fabs(lu[i][k]) >= best * kTie) flag = true;          // (near-)tie: pivot order not certain
    if (!(best > 0.0) || !isfinite(best)) { flag = true; best = 1.0; }
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      if (i == piv) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { const double t = lu[k][j]; lu[k][j] = lu[i][j]; lu[i][j] = t; }
        const int t = perm[k]; perm[k] = perm[i]; perm[i] = t;
      }
    }
    const double pivot = flag && lu[k][k] == 0.0 ? 1.0 : lu[k][k];
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      const double l = lu[i][k] / pivot;
      lu[i][k] = l;
#pragma unroll
      for (int j = k + 1; j < 3; ++j) lu[i][j] -= l * lu[k][j];
    }
  }
  if (!(fabs(lu[2][2]) > 0.0) || !isfinite(lu[2][2])) { flag = true; lu[2][2] = 1.0; }
  // X = U^-1 L^-1 P
  double x[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double b0 = perm[0] == j ? 1.0 : 0.0, b1 = perm[1] == j ? 1.0 : 0.0, b2 = perm[2] == j ? 1.0 : 0.0;
    const double y0 = b0, y1 = b1 - lu[1][0] * y0, y2 = b2 - lu[2][0] * y0 - lu[2][1] * y1;
    const double x2 = y2 / lu[2][2];
    const double x1 = (y1 - lu[1][2] * x2) / lu[1][1];
    const double x0 = (y0 - lu[0][1] * x1 - lu[0][2] * x2) / lu[0][0];
    x[0][j] = x0; x[1][j] = x1; x[2][j] = x2;
  }
  // W = P^T |L||U| (rows back in A's order), then E = c u |X| W |X|
  double w[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double m = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double l = k < i ? fabs(lu[i][k]) : (k == i ? 1.0 : 0.0);
        const double u = k <= j ? fabs(lu[k][j]) : 0.0;
        m += l * u;
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
        if (perm[i] == r) w[r][j] = m;
    }
  double wx[3][3];                                  // W |X|
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      wx[i][j] = w[i][0] * fabs(x[0][j]) + w[i][1] * fabs(x[1][j]) + w[i][2] * fabs(x[2][j]);
  constexpr double kCu = 64.0 * 1.1102230246251565e-16;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double v = x[i][j];
      const double e = 2.0 * kCu * (fabs(x[i][0]) * wx[0][j] + fabs(x[i][1]) * wx[1][j] + fabs(x[i][2]) * wx[2][j]);
      const float f = (float)v;
      const double av = fabs(v);
      if (!(av > 1e-30) || !(av < 1e30) || !isfinite(e)) {
        flag = true;
      } else {
        const double lo = 0.5 * ((double)f + (double)nextafterf(f, -INFINITY));
        const double hi = 0.5 * ((double)f + (double)nextafterf(f, INFINITY));
        if (!(v - e > lo) || !(v + e < hi)) flag = true;
      }
      out[(size_t)cell * 9 + i * 3 + j] = f;
    }
  flags[cell] = flag ? 1 : 0;
}

int launch_inv_grid(const float *h, int cells, float *out, unsigned char *flags, cudaStream_t st) {
  if (cells == 0) return 0;
  k_inv_grid<<<(cells + 127) / 128, 128, 0, st>>>(h, cells, out, flags);
  return check_cuda(cudaGetLastError(), "k_inv_grid launch");
}

// ------------------------------------------------------------------------------------ k_weight_bound
// t_bound[scene] >= |s v - s x| for every (anchor v, keypoint x) of the scene: the diagonal of the two sets' bounding
// boxes taken together.  K1 drops the clamp max(w, gamma^2) for a scene whose bound shows that no weight reaches it.
constexpr int kBoundThreads = 1024;   // one CTA per scene: 40 000 anchors are 39 loads per thread, eight in flight
__global__ void __launch_bounds__(kBoundThreads) k_weight_bound(const float2 *__restrict__ src_raw, const int *__restrict__ counts,
                                                      int n_points, double scale, const float2 *__restrict__ anchors,
                                                      int cells, float *__restrict__ t_bound) {
  const int scene = blockIdx.x, tid = threadIdx.x;
  const int n = counts ? min(max(counts[scene], 0), n_points) : n_points;
  src_raw += (size_t)scene * n_points;
  anchors += (size_t)scene * cells;
  float lo[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, hi[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 8
  for (int i = tid; i < n; i += kBoundThreads) {             // unrolled: eight loads in flight per thread (one CTA covers a scene)
    const float2 v = src_raw[i];
    lo[0] = fminf(lo[0], v.x); hi[0] = fmaxf(hi[0], v.x);
    lo[1] = fminf(lo[1], v.y); hi[1] = fmaxf(hi[1], v.y);
  }
#pragma unroll 8
  for (int i = tid; i < cells; i += kBoundThreads) {
    const float2 v = anchors[i];
    lo[2] = fminf(lo[2], v.x); hi[2] = fmaxf(hi[2], v.x);
    lo[3] = fminf(lo[3], v.y); hi[3] = fmaxf(hi[3], v.y);
  }
  __shared__ float red[2][4][kBoundThreads / 32];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    if ((tid & 31) == 0) { red[0][k][tid >> 5] = lo[k]; red[1][k][tid >> 5] = hi[k]; }
  }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      for (int w = 1; w < kBoundThreads / 32; ++w) { lo[k] = fminf(lo[0 + k], red[0][k][w]); hi[k] = fmaxf(hi[k], red[1][k][w]); }
    // keypoints go into the table as float(s * x): the same product here, in double, then the widest gap per axis
    const double kx0 = scale * lo[0], kx1 = scale * hi[0], ky0 = scale * lo[1], ky1 = scale * hi[1];
    const double dx = fmax(fabs(kx1 - lo[2]), fabs(hi[2] - kx0)), dy = fmax(fabs(ky1 - lo[3]), fabs(hi[3] - ky0));
    const double t = sqrt(dx * dx + dy * dy) * (1.0 + 1e-6);
    // no keypoints or no cells, NaN or infinite coordinates: no bound (+inf keeps the clamp)
    t_bound[scene] = (n > 0 && cells > 0 && isfinite(t)) ? (float)t + 1e-6f : INFINITY;
  }
}

int launch_weight_bound(const float *src_raw, const int *counts, int batch, int n_points, double scale,
                        const float *anchors, int cells, float *t_bound, cudaStream_t st) {
  k_weight_bound<<<batch, kBoundThreads, 0, st>>>(reinterpret_cast<const float2 *>(src_raw), counts, n_points, scale,
                                        reinterpret_cast<const float2 *>(anchors), cells, t_bound);
  return check_cuda(cudaGetLastError(), "k_weight_bound launch");
}

// ------------------------------------------------------------------------------------ k_scale_anchors
// anchors = float32(vertices * s): numpy's `(vertices * scale).astype(np.float32)` (apap.scale_anchors), on the device
__global__ void __launch_bounds__(256) k_scale_anchors(const double *__restrict__ v, size_t n, double scale,
                                                       float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n) out[i] = __double2float_rn(__dmul_rn(v[i], scale));
}

// The device buffers of one whole local_homography call (apap_local_homography_points), in one workspace
struct PassLayout {
  size_t cond, mats, tmats, anchors, bound, rows, blocks, partials, total;
  int n_pad;
};

static PassLayout pass_layout(int batch, int n_points, int cells, int engine) {
  PassLayout l;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t at = off; off = (off + bytes + 255) / 256 * 256; return at; };
  l.n_pad = (n_points + kChunk - 1) / kChunk * kChunk;
  if (l.n_pad < kChunk) l.n_pad = kChunk;
  const GramPlan p = make_gram_plan(cells, l.n_pad, engine);
  l.cond = take((size_t)2 * batch * n_points * 2 * sizeof(float));
  l.mats = take((size_t)batch * 36 * sizeof(float));
  l.tmats = take((size_t)batch * 18 * sizeof(double));
  l.anchors = take((size_t)batch * cells * 2 * sizeof(float));
  l.bound = take((size_t)batch * sizeof(float));
  l.rows = take((size_t)batch * l.n_pad * kRowFloats * sizeof(float));
  l.blocks = take(engine == APAP_GRAM_TCGEN05 ? (size_t)batch * (l.n_pad / APAP_KP_BLOCK) * APAP_KP_BLOCK_FLOATS * sizeof(float) : 0);
  l.partials = take((size_t)batch * p.k_splits * kTerms * p.cells_padded * sizeof(float));
  l.total = off;
  return l;
}

}  // namespace apap

using namespace apap;

extern "C" {

int apap_kp_rows(const float *src_cond, const float *dst_cond, const float *src_raw, const int *counts, int batch,
                 int n_points, int n_kp_padded, double scale, float *kp_table, void *stream) {
  if (!src_cond || !dst_cond || !src_raw || !kp_table) return fail(APAP_E_BADARG, "null pointer");
  if (batch <= 0 || n_points <= 0 || n_kp_padded < n_points || n_kp_padded % kChunk)
    return fail(APAP_E_BADARG, "kp_rows: bad sizes");
  if ((reinterpret_cast<uintptr_t>(src_cond) | reinterpret_cast<uintptr_t>(dst_cond) |
       reinterpret_cast<uintptr_t>(src_raw)) & 7u)
    return fail(APAP_E_ALIGN, "kp_rows: the point arrays must be 8-byte aligned");
  if (reinterpret_cast<uintptr_t>(kp_table) & 15u) return fail(APAP_E_ALIGN, "kp_rows: kp_table must be 16-byte aligned");
  return launch_kp_rows(src_cond, dst_cond, src_raw, counts, batch, n_points, n_kp_padded, scale, kp_table,
                        static_cast<cudaStream_t>(stream));
}

int apap_weight_bound(const float *src_raw, const int *counts, int batch, int n_points, double scale,
                      const float *anchors, int cells, float *t_bound, void *stream) {
  if (!src_raw || !anchors || !t_bound) return fail(APAP_E_BADARG, "null pointer");
  if (batch <= 0 || n_points <= 0 || cells <= 0 || batch > 65535) return fail(APAP_E_BADARG, "weight_bound: bad sizes");
  if ((reinterpret_cast<uintptr_t>(src_raw) | reinterpret_cast<uintptr_t>(anchors)) & 7u)
    return fail(APAP_E_ALIGN, "weight_bound: the point arrays must be 8-byte aligned");
  return launch_weight_bound(src_raw, counts, batch, n_points, scale, anchors, cells, t_bound,
                             static_cast<cudaStream_t>(stream));
}

int apap_condition(const float *src, const float *dst, const int *counts, int batch, int n_points, float *cond,
                   float *mats, double *tmats, void *stream) {
  if (!src || !dst || !cond || !mats || !tmats) return fail(APAP_E_BADARG, "null pointer");
  if (batch <= 0 || n_points <= 0 || batch > 65535) return fail(APAP_E_BADARG, "condition: bad sizes");
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(cond)) & 7u)
    return fail(APAP_E_ALIGN, "condition: the point arrays must be 8-byte aligned");
  return launch_condition(src, dst, counts, batch, n_points, cond, mats, tmats, static_cast<cudaStream_t>(stream));
}

int apap_kp_blocks(const float *kp_table, int batch, int n_kp_padded, float *kp_blocks, void *stream) {
  if (!kp_table || !kp_blocks) return fail(APAP_E_BADARG, "null pointer");
  if (batch <= 0 || n_kp_padded <= 0 || n_kp_padded % kChunk) return fail(APAP_E_BADARG, "kp_blocks: bad sizes");
  return launch_kp_blocks(kp_table, batch, n_kp_padded, kp_blocks, static_cast<cudaStream_t>(stream));
}

int apap_pass_workspace_bytes(int batch, int n_points, int cells, int engine, size_t *bytes) {
  if (!bytes) return fail(APAP_E_BADARG, "null pointer");
  if (batch <= 0 || n_points <= 0 || cells <= 0) return fail(APAP_E_BADARG, "pass workspace: bad sizes");
  if (engine != APAP_GRAM_TCGEN05 && engine != APAP_GRAM_FFMA2) return fail(APAP_E_BADARG, "pass workspace: unknown engine");
  *bytes = pass_layout(batch, n_points, cells, engine).total;
  return 0;
}

int apap_local_homography_points(const float *src, const float *dst, const int *counts, int batch, int n_points,
                                 const double *vertices, int cells, double scale, float gamma_sq, int engine, int solver,
                                 void *workspace, size_t workspace_bytes, int *tile_counters, float *out_h,
                                 int *out_sweeps, void *stream) {
  if (!src || !dst || !vertices || !workspace || !out_h) return fail(APAP_E_BADARG, "null pointer");
  if (batch <= 0 || n_points <= 0 || cells <= 0) return fail(APAP_E_BADARG, "local_homography_points: bad sizes");
  if (engine != APAP_GRAM_TCGEN05 && engine != APAP_GRAM_FFMA2) return fail(APAP_E_BADARG, "local_homography_points: unknown engine");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255u) || (reinterpret_cast<uintptr_t>(vertices) & 7u))
    return fail(APAP_E_ALIGN, "local_homography_points: workspace must be 256-byte aligned, vertices 8-byte aligned");
  const PassLayout l = pass_layout(batch, n_points, cells, engine);
  if (workspace_bytes < l.total) return fail(APAP_E_BADARG, "local_homography_points: workspace smaller than apap_pass_workspace_bytes");
  char *ws = static_cast<char *>(workspace);
  float *cond = reinterpret_cast<float *>(ws + l.cond), *mats = reinterpret_cast<float *>(ws + l.mats);
  double *tmats = reinterpret_cast<double *>(ws + l.tmats);
  float *anchors = reinterpret_cast<float *>(ws + l.anchors), *bound = reinterpret_cast<float *>(ws + l.bound);
  float *rows = reinterpret_cast<float *>(ws + l.rows), *blocks = reinterpret_cast<float *>(ws + l.blocks);
  float *partials = reinterpret_cast<float *>(ws + l.partials);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n_anchor = (size_t)batch * cells * 2;
  k_scale_anchors<<<(unsigned)((n_anchor + 255) / 256), 256, 0, st>>>(vertices, n_anchor, scale, anchors);
  int rc = check_cuda(cudaGetLastError(), "k_scale_anchors launch");
  if (rc) return rc;
  if ((rc = apap_condition(src, dst, counts, batch, n_points, cond, mats, tmats, stream))) return rc;
  const float *cf1 = cond, *cf2 = cond + (size_t)batch * n_points * 2;
  if ((rc = apap_kp_rows(cf1, cf2, src, counts, batch, n_points, l.n_pad, scale, rows, stream))) return rc;
  if ((rc = apap_weight_bound(src, counts, batch, n_points, scale, anchors, cells, bound, stream))) return rc;
  const float *table = rows;
  if (engine == APAP_GRAM_TCGEN05) {
    if ((rc = apap_kp_blocks(rows, batch, l.n_pad, blocks, stream))) return rc;
    table = blocks;
  }
  return apap_local_homography(table, anchors, tmats, batch, cells, l.n_pad, gamma_sq, engine, solver, bound, partials,
                               tile_counters, out_h, out_sweeps, stream);
}

int apap_invert_grid(const float *grid, int cells, float *grid_inv, unsigned char *flags, void *stream) {
  if (!grid || !grid_inv || !flags || cells < 0) return fail(APAP_E_BADARG, "invert_grid: bad arguments");
  return launch_inv_grid(grid, cells, grid_inv, flags, static_cast<cudaStream_t>(stream));
}

int apap_warp_tables(const float *cell_hinv, const int *col_extent, const int *row_extent, int grid_rows, int grid_cols,
                     int off_x, int off_y, int src_w, int src_h, float *cell_fast, void *stream) {
  if (!cell_hinv || !col_extent || !row_extent || !cell_fast) return fail(APAP_E_BADARG, "null pointer");
  if (grid_rows <= 0 || grid_cols <= 0) return fail(APAP_E_BADARG, "warp tables: bad grid");
  if ((reinterpret_cast<uintptr_t>(cell_fast) & 15u) || (reinterpret_cast<uintptr_t>(col_extent) & 7u) ||
      (reinterpret_cast<uintptr_t>(row_extent) & 7u))
    return fail(APAP_E_ALIGN, "warp tables: cell_fast must be 16-byte, the extents 8-byte aligned");
  return launch_warp_prep(cell_hinv, col_extent, row_extent, grid_rows, grid_cols, off_x, off_y, src_w, src_h, cell_fast,
                          static_cast<cudaStream_t>(stream));
}

}  // extern "C"
