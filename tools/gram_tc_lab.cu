// Lab for the tensor-core Gram kernel (K1 on tcgen05): weights by SIMT producers into TMEM, the
// [128 cells x 8 keypoints] . [8 x 24] products as 3xTF32 tcgen05.mma with the accumulator in TMEM.
// Validates against a plain FP32 SIMT kernel on the same synthetic data and times both.
//   build: make -C tools gram_tc_lab ; run on the GPU box: timeout 60 tools/gram_tc_lab [cells] [n_kp]
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../cvx_proj_b200/csrc/common.cuh"

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
using namespace apap;

// ------------------------------------------------------------------------------------------ layout
constexpr int kKB = 8;                      // keypoints per k-block (one tf32 MMA K step)
constexpr int kNT = 32;                     // accumulator columns (24 terms padded to the MMA N)
constexpr int kKbFloats = 2 * kNT * kKB + 2 * kKB;   // Bh tile, Bl tile, kx[8], ky[8] = 528 floats
constexpr int kKbBytes = kKbFloats * 4;     // 2112
constexpr int kStageKb = 8;                 // k-blocks per shared-memory stage (64 keypoints)
constexpr int kStageBytesTc = kStageKb * kKbBytes;   // 16896
constexpr int kSmemStages = 3;
constexpr int kTmemStages = 4;              // A ring in TMEM: 4 x (8 hi + 8 lo columns)
constexpr int kTmemCols = 128;              // 32 (D1) + 32 (D2) + 4 * 16 (A)
constexpr int kSegKb = 32;                  // k-blocks per accumulation segment (256 keypoints), then drained
constexpr int kTcThreads = 192;             // warps 0-3 producers / epilogue, warp 4 MMA, warp 5 TMA

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem descriptor], kind::tf32, issued by one thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B; LBO = bytes between the two 16-B K chunks,
// SBO = bytes between 8-row groups (cute/arch/mma_sm100_desc.hpp SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kNT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// ------------------------------------------------------------------------------------------ kernel
struct TcSmem {
  alignas(128) float stage[kSmemStages][kStageBytesTc / 4];
  uint64_t smem_full[kSmemStages];      // TMA -> producers + MMA
  uint64_t smem_empty[kSmemStages];     // MMA (commit) -> TMA
  uint64_t a_full[kTmemStages];         // producers -> MMA   (one arrival per producer warp)
  uint64_t a_empty[kTmemStages];        // MMA (commit) -> producers
  uint64_t d_full;                      // MMA (commit) -> producers: a segment's sums are complete
  uint64_t d_empty;                     // producers -> MMA: the accumulators have been drained (one per warp)
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kTcThreads) gram_tc(const float *__restrict__ table_tc, const float *__restrict__ anchors,
                                                       int cells, int cells_padded, int n_kb, int kb_per_split, float gamma_sq,
                                                       float *__restrict__ partials, int dbg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem &sm = *reinterpret_cast<TcSmem *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.y;
  const int kb0 = split * kb_per_split;
  const int nkb = min(n_kb, kb0 + kb_per_split) - kb0;            // k-blocks of this CTA (multiple of kStageKb)
  const int n_stage = nkb / kStageKb;

  if (tid == 0) {
    for (int s = 0; s < kSmemStages; ++s) { mbar_init(&sm.smem_full[s], 1); mbar_init(&sm.smem_empty[s], 1); }
    for (int s = 0; s < kTmemStages; ++s) { mbar_init(&sm.a_full[s], 4); mbar_init(&sm.a_empty[s], 1); }
    mbar_init(&sm.d_full, 1);
    mbar_init(&sm.d_empty, 4);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t tmem_d = tmem;                    // columns [0, 32): hi x hi sums; [32, 64): the two cross terms
  const uint32_t tmem_a = tmem + 2 * kNT;          // columns [64, 128): stage s -> hi at 16 s, lo at 16 s + 8
  const int n_seg = (nkb + kSegKb - 1) / kSegKb;

  if (warp < 4) {
    // ================= producers: thread = cell row of the tile = TMEM lane ==================
    const int cell = min(blockIdx.x * 128 + tid, cells - 1);
    const float2 av = reinterpret_cast<const float2 *>(anchors)[cell];
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float acc[24];
#pragma unroll
    for (int t = 0; t < 24; ++t) acc[t] = 0.f;
    int seg_done = 0;
    auto drain = [&]() {                           // add the finished segment's TMEM sums into registers
      mbar_wait(&sm.d_full, seg_done & 1);
      tc_fence_after();
      uint32_t v0[16], v1[16], u0[16], u1[16];
      tmem_ld16(tmem_d + lane_base, v0);
      tmem_ld16(tmem_d + lane_base + 16, v1);
      tmem_ld16(tmem_d + lane_base + 32, u0);
      tmem_ld16(tmem_d + lane_base + 48, u1);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.d_empty);
#pragma unroll
      for (int t = 0; t < 16; ++t) acc[t] += __uint_as_float(v0[t]) + __uint_as_float(u0[t]);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[16 + t] += __uint_as_float(v1[t]) + __uint_as_float(u1[t]);
      ++seg_done;
    };
    int kb = 0;
    for (int st = 0; st < n_stage; ++st) {
      const int ss = st % kSmemStages;
      mbar_wait(&sm.smem_full[ss], (st / kSmemStages) & 1);
      const float *stg = sm.stage[ss];
      for (int g = 0; g < kStageKb; ++g, ++kb) {
        const int ts = kb % kTmemStages;
        mbar_wait(&sm.a_empty[ts], ((kb / kTmemStages) & 1) ^ 1);      // first pass: free
        tc_fence_after();
        const float4 *co = reinterpret_cast<const float4 *>(stg + g * kKbFloats + 2 * kNT * kKB);
        const float4 x0 = co[0], x1 = co[1], y0 = co[2], y1 = co[3];
        const float2 kx[4] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y), make_float2(x1.z, x1.w)};
        const float2 ky[4] = {make_float2(y0.x, y0.y), make_float2(y0.z, y0.w), make_float2(y1.x, y1.y), make_float2(y1.z, y1.w)};
        uint32_t hi[8], lo[8];
        const float2 ax2 = make_float2(av.x, av.x), ay2 = make_float2(av.y, av.y);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 dx = __fadd2_rn(ax2, make_float2(-kx[k].x, -kx[k].y));
          const float2 dy = __fadd2_rn(ay2, make_float2(-ky[k].x, -ky[k].y));
          const float2 d2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
          const float w0 = dbg >= 3 ? fmaxf(d2.x, gamma_sq) : fmaxf(ex2_approx(-sqrt_approx(d2.x)), gamma_sq);
          const float w1 = dbg >= 3 ? fmaxf(d2.y, gamma_sq) : fmaxf(ex2_approx(-sqrt_approx(d2.y)), gamma_sq);
          const uint32_t h0 = (__float_as_uint(w0) + 0x1000u) & 0xFFFFE000u;    // round to tf32
          const uint32_t h1 = (__float_as_uint(w1) + 0x1000u) & 0xFFFFE000u;
          hi[2 * k] = h0; hi[2 * k + 1] = h1;
          lo[2 * k] = __float_as_uint(w0 - __uint_as_float(h0));
          lo[2 * k + 1] = __float_as_uint(w1 - __uint_as_float(h1));
        }
        // publish the PREVIOUS k-block now: its TMEM stores have had this k-block's arithmetic to land
        if (kb > 0) {
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.a_full[(kb - 1) % kTmemStages]);
        }
        tmem_st8(tmem_a + lane_base + ts * 16, hi);
        tmem_st8(tmem_a + lane_base + ts * 16 + 8, lo);
        // a segment behind: its MMAs have had kTmemStages k-blocks of time to retire
        if (kb % kSegKb == kTmemStages - 1 && kb >= kSegKb) drain();
      }
    }
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&sm.a_full[(nkb - 1) % kTmemStages]);
    while (seg_done < n_seg) drain();
    const int c = blockIdx.x * 128 + tid;
    if (c < cells) {
      float *dst = partials + (size_t)split * 24 * cells_padded + c;
#pragma unroll
      for (int t = 0; t < 24; ++t) dst[(size_t)t * cells_padded] = acc[t];
    }
  } else if (warp == 4) {
    // ================= MMA issuer (one thread) ===============================================
    if (lane == 0) {
      int kb = 0;
      for (int st = 0; st < n_stage; ++st) {
        const int ss = st % kSmemStages;
        mbar_wait(&sm.smem_full[ss], (st / kSmemStages) & 1);
        const uint32_t sbase = smem_u32(sm.stage[ss]);
        for (int g = 0; g < kStageKb; ++g, ++kb) {
          const int ts = kb % kTmemStages;
          mbar_wait(&sm.a_full[ts], (kb / kTmemStages) & 1);
          tc_fence_after();
          const uint64_t bh = smem_desc_kmajor(sbase + g * kKbBytes, 512, 128);
          const uint64_t bl = smem_desc_kmajor(sbase + g * kKbBytes + kNT * kKB * 4, 512, 128);
          const uint32_t a_hi = tmem_a + ts * 16, a_lo = a_hi + 8;
          const int in_seg = kb % kSegKb;
          if (in_seg == 0 && kb > 0) {             // the producers must have drained the previous segment
            mbar_wait(&sm.d_empty, ((kb / kSegKb) - 1) & 1);
            tc_fence_after();
          }
          mma_tf32_ts(tmem_d, a_hi, bh, kIdescTf32, in_seg > 0 ? 1u : 0u);            // hi x hi
          if (dbg < 2) mma_tf32_ts(tmem_d + kNT, a_lo, bh, kIdescTf32, in_seg > 0 ? 1u : 0u);      // cross terms apart:
          if (dbg < 1) mma_tf32_ts(tmem_d + kNT, a_hi, bl, kIdescTf32, 1u);                        //  they never touch the big sums
          mma_commit(&sm.a_empty[ts]);             // the A columns (and, below, the smem stage) are free when these retire
          if (in_seg == kSegKb - 1 || kb == nkb - 1) mma_commit(&sm.d_full);
        }
        mma_commit(&sm.smem_empty[ss]);
      }
    }
  } else {
    // ================= TMA producer (one thread) =============================================
    if (lane == 0) {
      const char *src = reinterpret_cast<const char *>(table_tc) + (size_t)kb0 * kKbBytes;
      for (int st = 0; st < n_stage; ++st) {
        const int ss = st % kSmemStages;
        mbar_wait(&sm.smem_empty[ss], ((st / kSmemStages) & 1) ^ 1);   // first pass: free
        mbar_arrive_expect_tx(&sm.smem_full[ss], kStageBytesTc);
        bulk_g2s(sm.stage[ss], src + (size_t)st * kStageBytesTc, kStageBytesTc, &sm.smem_full[ss]);
      }
    }
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------- FP32 SIMT reference
__global__ void gram_ref(const float *__restrict__ kp, const float *__restrict__ anchors, int cells, int cells_padded,
                         int n_kp, int kp_per_split, float gamma_sq, float *__restrict__ partials) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int split = blockIdx.y;
  if (c >= cells) return;
  const float2 av = reinterpret_cast<const float2 *>(anchors)[c];
  double acc[24];
  for (int t = 0; t < 24; ++t) acc[t] = 0.0;
  const int i1 = min(n_kp, (split + 1) * kp_per_split);
  for (int i = split * kp_per_split; i < i1; ++i) {
    const float *row = kp + (size_t)i * 28;
    const float dx = av.x - row[24], dy = av.y - row[26];
    const float w = fmaxf(exp2f(-sqrtf(fmaf(dy, dy, dx * dx))), gamma_sq);
    for (int t = 0; t < 24; ++t) acc[t] += (double)w * (double)row[t];
  }
  for (int t = 0; t < 24; ++t) partials[((size_t)split * 24 + t) * cells_padded + c] = (float)acc[t];
}

int main(int argc, char **argv) {
  const int cells = argc > 1 ? atoi(argv[1]) : 40000;
  const int n_kp = argc > 2 ? atoi(argv[2]) : 5000;
  const int n_pad = (n_kp + 127) / 128 * 128;
  const int n_chunks = n_pad / 128;
  int cps = (n_chunks + 3) / 4; if (cps > 8) cps = 8;
  if (argc > 3) cps = atoi(argv[3]);
  const int dbg = argc > 4 ? atoi(argv[4]) : 0;
  const int splits = (n_chunks + cps - 1) / cps;
  const int cells_padded = (cells + 511) / 512 * 512;
  const int n_kb = n_pad / kKB, kb_per_split = cps * 128 / kKB;
  printf("cells %d n_kp %d (padded %d) splits %d k-blocks %d (%d per split)\n", cells, n_kp, n_pad, splits, n_kb, kb_per_split);
  std::vector<float> table((size_t)n_pad * 28, 0.f), anchors((size_t)cells * 2);
  srand(1);
  const float s = 2.0f * 1.4426950f / 1e4f;
  for (int i = 0; i < n_kp; ++i) {
    for (int t = 0; t < 24; ++t) table[(size_t)i * 28 + t] = (rand() % 2001 - 1000) * 1e-3f * (1.f + 0.37f * t);
    const float kx = (rand() % 3840) * s, ky = (rand() % 2160) * s;
    table[(size_t)i * 28 + 24] = kx; table[(size_t)i * 28 + 25] = kx;
    table[(size_t)i * 28 + 26] = ky; table[(size_t)i * 28 + 27] = ky;
  }
  const int side = (int)ceil(sqrt((double)cells));
  for (int c = 0; c < cells; ++c) { anchors[2 * c] = (c % side) * 4448.f / side * s; anchors[2 * c + 1] = (c / side) * 2332.f / side * s; }
  // tensor-core table: per k-block Bh tile, Bl tile (K-major core matrices), kx[8], ky[8]
  std::vector<float> tc((size_t)n_kb * kKbFloats, 0.f);
  for (int kb = 0; kb < n_kb; ++kb) {
    float *blk = tc.data() + (size_t)kb * kKbFloats;
    for (int k = 0; k < kKB; ++k) {
      const float *row = table.data() + (size_t)(kb * kKB + k) * 28;
      for (int n = 0; n < 24; ++n) {
        union { float f; uint32_t u; } v, h;
        v.f = row[n];
        h.u = (v.u + 0x1000u) & 0xFFFFE000u;
        const int idx = (k / 4) * 128 + (n / 8) * 32 + (n % 8) * 4 + (k % 4);
        blk[idx] = h.f;
        blk[kNT * kKB + idx] = v.f - h.f;
      }
      blk[2 * kNT * kKB + k] = row[24];
      blk[2 * kNT * kKB + 8 + k] = row[26];
    }
  }
  float *d_table, *d_tc, *d_anchors, *d_ref, *d_out;
  const size_t pbytes = (size_t)splits * 24 * cells_padded * 4;
  CHECK(cudaMalloc(&d_table, table.size() * 4)); CHECK(cudaMalloc(&d_tc, tc.size() * 4));
  CHECK(cudaMalloc(&d_anchors, anchors.size() * 4)); CHECK(cudaMalloc(&d_ref, pbytes)); CHECK(cudaMalloc(&d_out, pbytes));
  CHECK(cudaMemcpy(d_table, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(d_tc, tc.data(), tc.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(d_anchors, anchors.data(), anchors.size() * 4, cudaMemcpyHostToDevice));
  CHECK(cudaMemset(d_ref, 0, pbytes)); CHECK(cudaMemset(d_out, 0, pbytes));
  gram_ref<<<dim3((cells + 127) / 128, splits), 128>>>(d_table, d_anchors, cells, cells_padded, n_pad, cps * 128, 0.25f, d_ref);
  CHECK(cudaDeviceSynchronize());

  const size_t smem = sizeof(TcSmem);
  CHECK(cudaFuncSetAttribute(gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((cells + 127) / 128, splits);
  printf("gram_tc: grid %d x %d, %d threads, %zu B smem\n", grid.x, grid.y, kTcThreads, smem);
  gram_tc<<<grid, kTcThreads, smem>>>(d_tc, d_anchors, cells, cells_padded, n_kb, kb_per_split, 0.25f, d_out, dbg);
  CHECK(cudaDeviceSynchronize());
  std::vector<float> ref(pbytes / 4), out(pbytes / 4);
  CHECK(cudaMemcpy(ref.data(), d_ref, pbytes, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(out.data(), d_out, pbytes, cudaMemcpyDeviceToHost));
  double max_rel = 0, sum_rel = 0; size_t cnt = 0; double max_abs_ref = 0;
  for (int sp = 0; sp < splits; ++sp)
    for (int t = 0; t < 24; ++t) {
      double scale = 0;
      for (int c = 0; c < cells; ++c) scale = fmax(scale, fabs((double)ref[((size_t)sp * 24 + t) * cells_padded + c]));
      max_abs_ref = fmax(max_abs_ref, scale);
      for (int c = 0; c < cells; ++c) {
        const size_t i = ((size_t)sp * 24 + t) * cells_padded + c;
        const double e = fabs((double)out[i] - (double)ref[i]) / (scale + 1e-30);
        max_rel = fmax(max_rel, e); sum_rel += e; ++cnt;
      }
    }
  printf("tensor-core vs float64-accumulated reference: max err / max|ref| per (split, term) = %.3e, mean %.3e (ref max %.3e)\n",
         max_rel, sum_rel / cnt, max_abs_ref);
  printf("sample: out %.6f ref %.6f | out %.6f ref %.6f\n", out[5], ref[5], out[(size_t)3 * cells_padded + 777], ref[(size_t)3 * cells_padded + 777]);

  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CHECK(cudaEventRecord(e0));
    gram_tc<<<grid, kTcThreads, smem>>>(d_tc, d_anchors, cells, cells_padded, n_kb, kb_per_split, 0.25f, d_out, dbg);
    CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
    float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
  }
  CHECK(cudaGetLastError());
  const double flops = 2.0 * 24 * (double)n_pad * cells;
  printf("gram_tc %.3f ms  -> %.2f TFLOP/s of the 24-term contraction (%.1f Gpair/s)\n", best, flops / (best * 1e-3) / 1e12,
         (double)n_pad * cells / (best * 1e-3) / 1e9);
  return 0;
}
