// K1 on the tensor cores -- moving-DLT weights + Gram contraction (reference pyviz/apap.py:150-152,159)
// as tcgen05 (5th-generation tensor core) MMAs with the accumulators in tensor memory.
//
// The contraction  S_t(c) = sum_i w2(c, i) * P[i][t]  is a GEMM  [cells x N] . [N x 24]  whose A
// operand does not exist in memory: it is generated on the fly.  Per CTA (one 128-cell tile x one
// keypoint split):
//   * warps 0-3 (producers; thread = cell = accumulator row = TMEM lane) compute, for every block
//     of 8 keypoints, the 8 weights  w2 = max(2^-|s v - s x|, gamma^2)  (packed FP32x2 adds / FMA,
//     MUFU.SQRT, MUFU.EX2), split each into a TF32 head and tail (hi = cvt.rna.tf32(w2),
//     lo = w2 - hi, exact) and store both straight into tensor memory (tcgen05.st, SASS STTM) --
//     the A operand of the MMA is read from TMEM, so the weights never touch shared memory;
//   * warp 4 (one elected thread) issues, per 8-keypoint block, two TF32 MMAs (SASS UTCHMMA):
//     M128 x N64 x K8:  [D1 | D2] += hi . [Ph | Pl]   and   M128 x N32 x K8:  D2 += lo . Ph
//     (3xTF32: the product terms P are split into Ph + Pl on the host; the dropped lo . Pl term is
//     2^-22 relative).  The two cross terms accumulate in their own TMEM columns D2, so they never
//     re-round the big sums D1;
//   * warp 5 (one thread) streams the keypoint blocks -- the 64 x 8 tile [Ph | Pl] in the K-major
//     core-matrix layout of the MMA's shared-memory descriptor, and the 8 pre-scaled keypoint
//     coordinates -- by TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) into a 4-stage ring;
//   * mbarriers connect the three roles: smem_full/smem_empty (TMA <-> producers + MMA),
//     a_full/a_empty (producers <-> MMA: a 2-deep ring of 16-keypoint steps of A columns in TMEM,
//     one hand-over per step), d_full/d_empty (segment drains).
//
// Accuracy.  The tensor core accumulates in FP32 and aligns/truncates the accumulator at every
// MMA, a drift proportional to the number of MMAs that touch a big accumulator.  So (a) only one
// MMA per keypoint block touches D1, and (b) every 256 keypoints (kSegKb blocks) the producers
// drain D1 + D2 from TMEM (tcgen05.ld, SASS LDTM) into FP32 sums in shared memory and the MMAs
// restart from zero; the per-split sums (<= 1024 keypoints, like the FFMA2 kernel) are written as
// the same partial-sum layout and combined in float64 by K2.  Measured against a float64-accumulated
// reference: 5e-7 mean / 7e-6 max of the largest sum per term (tools/gram_tc_lab.cu).
#include "common.cuh"

namespace apap {

// lab knobs (tools/variants.sh builds the library with -D...; the defaults are the product)
#ifndef APAP_TC_TRUNC
#define APAP_TC_TRUNC 1          // 1: split the weight by truncation (hi = w & mask) instead of rounding
#endif
#ifndef APAP_TC_SLEEP
#define APAP_TC_SLEEP 0          // > 0: nanosleep(ns) back-off in the single-thread roles' spin loops
#endif
#ifndef APAP_TC_NOMUFU
#define APAP_TC_NOMUFU 0         // 1: diagnostic, weights without the two MUFU operations (wrong results)
#endif
#ifndef APAP_TC_PIPELINE
#define APAP_TC_PIPELINE 1       // 0: the un-pipelined producer loop (kept for A/B timing)
#endif
#ifndef APAP_TC_NODRAIN
#define APAP_TC_NODRAIN 0        // 1: diagnostic, segments are never drained (wrong results)
#endif
#ifndef APAP_TC_NOSTTM
#define APAP_TC_NOSTTM 0         // 1: diagnostic, the weights are not stored to TMEM (wrong results)
#endif
#ifndef APAP_TC_NOMMA
#define APAP_TC_NOMMA 0          // 1: diagnostic, no MMA is issued, only the commits (wrong results)
#endif
#ifndef APAP_TC_NOWAIT
#define APAP_TC_NOWAIT 0         // 1: diagnostic, producers never wait for the A ring (wrong results)
#endif

constexpr int kKB = APAP_KP_BLOCK;                 // keypoints per k-block = K of one TF32 MMA
constexpr int kNT = 32;                            // accumulator columns (24 terms padded to the MMA N)
constexpr int kKbFloats = APAP_KP_BLOCK_FLOATS;    // [Ph | Pl] tile 512, s*kx[8], s*ky[8]
constexpr int kKbBytes = kKbFloats * 4;            // 2112
constexpr int kStepKb = 2;                         // k-blocks per producer/MMA hand-over ("step" = 16 keypoints)
constexpr int kStageKb = 4;                        // k-blocks per shared-memory stage (2 steps)
constexpr int kStageBytesTc = kStageKb * kKbBytes; // 8448
#ifndef APAP_TC_SMEM_STAGES
#define APAP_TC_SMEM_STAGES 4
#endif
constexpr int kSmemStages = APAP_TC_SMEM_STAGES;
#ifndef APAP_TC_TMEM_STAGES
#define APAP_TC_TMEM_STAGES 2
#endif
#ifndef APAP_TC_SMEM_PAD
#define APAP_TC_SMEM_PAD 0       // lab: extra dynamic shared memory (limits the CTAs per SM)
#endif
constexpr int kTmemStages = APAP_TC_TMEM_STAGES;   // A ring in TMEM: steps x 2 k-blocks x (8 hi + 8 lo columns)
constexpr int kTmemCols = kTmemStages <= 2 ? 128 : kTmemStages <= 6 ? 256 : 512;   // 32 (D1) + 32 (D2) + stages * 32 (A)
constexpr int kSegKb = 32;                         // k-blocks per accumulation segment (256 keypoints)
constexpr int kSegSteps = kSegKb / kStepKb;
constexpr int kTcThreads = 192;                    // warps 0-3 producers, warp 4 MMA, warp 5 TMA
static_assert(kKbFloats == 2 * kNT * kKB + 2 * kKB, "k-block layout");
static_assert((kChunk / kKB) % kStageKb == 0 && kStageKb % kStepKb == 0, "chunks, stages and steps nest");
static_assert(kStageKb / kStepKb == 2 && kStepKb == 2, "the producers' software pipeline is written for 2 steps per stage");

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
#ifndef APAP_TC_NOFENCE
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
}
__device__ __forceinline__ void tc_fence_after() {
#ifndef APAP_TC_NOFENCE
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif
}
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
__device__ __forceinline__ void tmem_wait_st() {
#ifndef APAP_TC_NOWAITST
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem descriptor], kind::tf32, M = 128, issued by one thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when every MMA issued so far by this thread has retired (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
#if defined(APAP_TC_PLAINARRIVE)
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");   // diagnostic (with NOMMA)
  return;
#endif
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// Round a finite float to the nearest TF32 (low 13 mantissa bits zero; ties away, like cvt.rna.tf32.f32
// without its Inf/NaN handling -- the weights are in (0, 1]).
__device__ __forceinline__ uint32_t to_tf32(float x) {
#if APAP_TC_TRUNC
  return __float_as_uint(x) & 0xFFFFE000u;
#else
  return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
#endif
}
// spin on an mbarrier from a single-thread role: optional back-off so the spin does not steal issue slots
__device__ __forceinline__ void mbar_wait_role(uint64_t *bar, uint32_t parity) {
#if APAP_TC_SLEEP > 0
  while (!mbar_try_wait(bar, parity)) __nanosleep(APAP_TC_SLEEP);
#else
  mbar_wait(bar, parity);
#endif
}
// Shared-memory matrix descriptor, K-major, no swizzle: a core matrix is 8 rows x 16 B (128 contiguous
// bytes); LBO = bytes between the two 16-byte K chunks of a row, SBO = bytes between 8-row groups;
// bits 46-47 = descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Instruction descriptor: D = F32 (bit 4), A = B = TF32 (2 at bits 7 and 10), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ lab trace
#ifdef APAP_TC_TRACE
__device__ long long g_trace[4][160][4];           // [role][step][mark] SM clock of CTA (0,0,0)
#define TRACE(role, step, mark) do { if (blockIdx.x == APAP_TC_TRACE && blockIdx.y == 0 && (step) < 160) g_trace[role][step][mark] = clock64(); } while (0)
__device__ long long g_cta[8192][4];               // per CTA: SM id, globaltimer at entry, after the TMEM allocation, at exit
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#else
#define TRACE(role, step, mark) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------ kernel
struct TcSmem {
  alignas(128) float stage[kSmemStages][kStageBytesTc / 4];
  float acc[kTerms][128];               // per-split sums of the producers' drains (thread = column: conflict-free)
  uint64_t smem_full[kSmemStages];      // TMA -> producers + MMA
  uint64_t smem_empty[kSmemStages];     // MMA (commit) -> TMA
  uint64_t a_full[kTmemStages];         // producers -> MMA (one arrival per producer warp)
  uint64_t a_empty[kTmemStages];        // MMA (commit) -> producers
  uint64_t d_full;                      // MMA (commit) -> producers: a segment's sums are complete
  uint64_t d_empty;                     // producers -> MMA: the accumulators have been drained
  uint32_t tmem_base;
};
static_assert(sizeof(TcSmem) <= 227 * 1024, "fits the shared memory of an SM");

__global__ void __launch_bounds__(kTcThreads, 4) k_gram_tc(const float *__restrict__ kp_blocks,
                                                           const float *__restrict__ anchors, int cells,
                                                           int cells_padded, int n_kb, int kb_per_split, int k_splits,
                                                           float gamma_sq, float *__restrict__ partials) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TcSmem &sm = *reinterpret_cast<TcSmem *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.y, scene = blockIdx.z;
  const int kb0 = split * kb_per_split;
  const int nkb = min(n_kb, kb0 + kb_per_split) - kb0;            // k-blocks of this CTA: a multiple of kStageKb
  const int n_stage = nkb / kStageKb;
  const int n_step = nkb / kStepKb;
  const int n_seg = (n_step + kSegSteps - 1) / kSegSteps;
  kp_blocks += (size_t)scene * n_kb * kKbFloats;
  anchors += (size_t)scene * cells * 2;
  partials += (size_t)scene * k_splits * kTerms * cells_padded;

#ifdef APAP_TC_TRACE
  const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
  if (tid == 0 && cta_lin < 8192) {
    uint32_t smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    g_cta[cta_lin][0] = smid; g_cta[cta_lin][1] = gtimer();
  }
#endif
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
#ifdef APAP_TC_TRACE
  if (tid == 0 && cta_lin < 8192) g_cta[cta_lin][2] = gtimer();
#endif
  const uint32_t tmem = sm.tmem_base;
  const uint32_t tmem_d = tmem;                    // columns [0, 32): hi x Ph sums; [32, 64): the two cross terms
  const uint32_t tmem_a = tmem + 2 * kNT;          // columns [64, 128): step slot s, k-block e -> hi at 32 s + 16 e, lo + 8

  if (warp < 4) {
    // ================= producers: thread = cell of the tile = accumulator row = TMEM lane =====
    const int c = blockIdx.x * 128 + tid;
    const float2 av = reinterpret_cast<const float2 *>(anchors)[min(c, cells - 1)];
    const float2 ax2 = make_float2(av.x, av.x), ay2 = make_float2(av.y, av.y);
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    int seg_done = 0;
    auto drain = [&]() {                           // add the finished segment's TMEM sums into the split's sums
      mbar_wait(&sm.d_full, seg_done & 1);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {       // terms 0..15, then 16..23 (+ 8 padding columns)
        uint32_t v[16], u[16];
        tmem_ld16(tmem_d + lane_base + 16 * half, v);
        tmem_ld16(tmem_d + lane_base + kNT + 16 * half, u);
        tmem_wait_ld();
#pragma unroll
        for (int t = 0; t < (half ? 8 : 16); ++t) {
          const float add = __uint_as_float(v[t]) + __uint_as_float(u[t]);
          float &dst = sm.acc[16 * half + t][tid];
          dst = seg_done ? dst + add : add;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.d_empty);
      ++seg_done;
    };
#if APAP_TC_PIPELINE
    // Software pipeline over the steps: the weights of step s + 1 (LDS, packed geometry, 32 MUFU) are
    // issued BEFORE the synchronisation of step s (wait for the A slot, STTM, wait::st, fences, arrive --
    // ~250 cycles of latency on the barrier unit), so the XU pipe works through that latency instead of
    // idling; the full-barrier of the stage that step s + 2 reads is also waited for in that shadow.
    auto weights = [&](int s, float (&w)[2 * kStepKb * 4]) {      // w[8 e + 2 k + {0,1}] = keypoint 2k, 2k+1 of k-block e
      const int stg = s / (kStageKb / kStepKb);
      const float4 *co = reinterpret_cast<const float4 *>(sm.stage[stg % kSmemStages] + 2 * kNT * kKB) +
                         (s % (kStageKb / kStepKb)) * kStepKb * (kKbFloats / 4);
#pragma unroll
      for (int e = 0; e < kStepKb; ++e, co += kKbFloats / 4) {
        const float4 x0 = co[0], x1 = co[1], y0 = co[2], y1 = co[3];
        const float2 kx[4] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y),
                              make_float2(x1.z, x1.w)};
        const float2 ky[4] = {make_float2(y0.x, y0.y), make_float2(y0.z, y0.w), make_float2(y1.x, y1.y),
                              make_float2(y1.z, y1.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 dx = __fadd2_rn(ax2, make_float2(-kx[k].x, -kx[k].y));
          const float2 dy = __fadd2_rn(ay2, make_float2(-ky[k].x, -ky[k].y));
          const float2 d2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
#if APAP_TC_NOMUFU
          w[8 * e + 2 * k] = fmaxf(1.f - d2.x, gamma_sq);
          w[8 * e + 2 * k + 1] = fmaxf(1.f - d2.y, gamma_sq);
#else
          w[8 * e + 2 * k] = fmaxf(ex2_approx(-sqrt_approx(d2.x)), gamma_sq);
          w[8 * e + 2 * k + 1] = fmaxf(ex2_approx(-sqrt_approx(d2.y)), gamma_sq);
#endif
        }
      }
    };
    auto wait_stage_of = [&](int s) {                // the stage step s reads must have landed
      const int stg = s / (kStageKb / kStepKb);
      mbar_wait(&sm.smem_full[stg % kSmemStages], (stg / kSmemStages) & 1);
    };
    auto hand_over = [&](int s, const float (&w)[2 * kStepKb * 4]) {
      const int ts = s % kTmemStages;
      if (lane == 0 && warp == 0) TRACE(0, s, 0);
#if !APAP_TC_NOWAIT
      mbar_wait(&sm.a_empty[ts], ((s / kTmemStages) & 1) ^ 1);   // first pass: free
#endif
      tc_fence_after();
      if (lane == 0 && warp == 0) TRACE(0, s, 1);
#pragma unroll
      for (int e = 0; e < kStepKb; ++e) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          hi[k] = to_tf32(w[8 * e + k]);
          lo[k] = __float_as_uint(w[8 * e + k] - __uint_as_float(hi[k]));
        }
        tmem_st8(tmem_a + lane_base + ts * 32 + e * 16, hi);
        tmem_st8(tmem_a + lane_base + ts * 32 + e * 16 + 8, lo);
      }
      if (lane == 0 && warp == 0) TRACE(0, s, 2);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.a_full[ts]);
      if (lane == 0 && warp == 0) TRACE(0, s, 3);
      // a segment behind: its MMAs have had a step's worth of time to retire
      if ((s & (kSegSteps - 1)) == 0 && s >= kSegSteps) drain();
    };
    float wa[2 * kStepKb * 4], wb[2 * kStepKb * 4];
    wait_stage_of(0);
    weights(0, wa);
    for (int s = 0; s < n_step; s += 2) {            // n_step is even (a stage = 2 steps)
      weights(s + 1, wb);                            // same stage as step s
      if (s + 2 < n_step) wait_stage_of(s + 2);
      hand_over(s, wa);
      if (s + 2 < n_step) weights(s + 2, wa);
      hand_over(s + 1, wb);
    }
#else
    int step = 0, ts = 0;
    uint32_t a_phase = 1;                          // parity to wait on for a_empty: the first pass is free
    for (int st = 0; st < n_stage; ++st) {
      const int ss = st % kSmemStages;
      mbar_wait(&sm.smem_full[ss], (st / kSmemStages) & 1);
      const float4 *co = reinterpret_cast<const float4 *>(sm.stage[ss] + 2 * kNT * kKB);
#pragma unroll 1
      for (int g = 0; g < kStageKb / kStepKb; ++g, ++step) {
        if (lane == 0 && warp == 0) TRACE(0, step, 0);
        if (lane == 0 && warp == 3) TRACE(3, step, 0);
#if !APAP_TC_NOWAIT
        mbar_wait(&sm.a_empty[ts], a_phase);
#endif
        tc_fence_after();
        if (lane == 0 && warp == 0) TRACE(0, step, 1);
        if (lane == 0 && warp == 3) TRACE(3, step, 1);
#pragma unroll
        for (int e = 0; e < kStepKb; ++e, co += kKbFloats / 4) {
          const float4 x0 = co[0], x1 = co[1], y0 = co[2], y1 = co[3];
          const float2 kx[4] = {make_float2(x0.x, x0.y), make_float2(x0.z, x0.w), make_float2(x1.x, x1.y),
                                make_float2(x1.z, x1.w)};
          const float2 ky[4] = {make_float2(y0.x, y0.y), make_float2(y0.z, y0.w), make_float2(y1.x, y1.y),
                                make_float2(y1.z, y1.w)};
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 dx = __fadd2_rn(ax2, make_float2(-kx[k].x, -kx[k].y));
            const float2 dy = __fadd2_rn(ay2, make_float2(-ky[k].x, -ky[k].y));
            const float2 d2 = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
#if APAP_TC_NOMUFU
            const float w0 = fmaxf(1.f - d2.x, gamma_sq), w1 = fmaxf(1.f - d2.y, gamma_sq);
#else
            const float w0 = fmaxf(ex2_approx(-sqrt_approx(d2.x)), gamma_sq);
            const float w1 = fmaxf(ex2_approx(-sqrt_approx(d2.y)), gamma_sq);
#endif
            hi[2 * k] = to_tf32(w0);
            hi[2 * k + 1] = to_tf32(w1);
            lo[2 * k] = __float_as_uint(w0 - __uint_as_float(hi[2 * k]));
            lo[2 * k + 1] = __float_as_uint(w1 - __uint_as_float(hi[2 * k + 1]));
          }
#if APAP_TC_NOSTTM
          if (hi[0] + hi[3] + lo[1] + lo[7] + hi[5] + lo[4] + hi[1] + hi[2] + hi[4] + hi[6] + hi[7] + lo[0] + lo[2] + lo[3] + lo[5] + lo[6] == 0x12345u) sm.acc[0][tid] = 1.f;
#else
          tmem_st8(tmem_a + lane_base + ts * 32 + e * 16, hi);
          tmem_st8(tmem_a + lane_base + ts * 32 + e * 16 + 8, lo);
#endif
        }
        if (lane == 0 && warp == 0) TRACE(0, step, 2);
        if (lane == 0 && warp == 3) TRACE(3, step, 2);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.a_full[ts]);
        if (lane == 0 && warp == 0) TRACE(0, step, 3);
        if (lane == 0 && warp == 3) TRACE(3, step, 3);
        if (++ts == kTmemStages) { ts = 0; a_phase ^= 1; }
        // a segment behind: its MMAs have had a step's worth of time to retire
#if !APAP_TC_NODRAIN
        if ((step & (kSegSteps - 1)) == 0 && step >= kSegSteps) drain();
#endif
      }
    }
#endif   // APAP_TC_PIPELINE
#if APAP_TC_NODRAIN
    mbar_wait(&sm.d_full, 0);
    seg_done = n_seg;
#endif
    while (seg_done < n_seg) drain();
    if (c < cells) {
      // partials[split][t][cell]: consecutive threads write consecutive cells (coalesced)
      float *dst = partials + (size_t)split * kTerms * cells_padded + c;
#pragma unroll
      for (int t = 0; t < kTerms; ++t) dst[(size_t)t * cells_padded] = sm.acc[t][tid];
    }
  } else if (warp == 4) {
    // ================= MMA issuer (one elected thread; the warp stays converged around it) ====
    // The issue thread shares a scheduler with a producer warp of every resident CTA, so its
    // instruction count matters: descriptors are built once per stage and advanced by adds.
    const uint32_t idesc64 = idesc_tf32(2 * kNT), idesc32 = idesc_tf32(kNT);
    const uint64_t desc_hi = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(1024 >> 4) << 16);
    int step = 0, ts = 0;
    uint32_t a_phase = 0;
    for (int st = 0; st < n_stage; ++st) {
      const int ss = st % kSmemStages;
      mbar_wait_role(&sm.smem_full[ss], (st / kSmemStages) & 1);
      // B tile of a k-block: rows 0..31 = Ph, rows 32..63 = Pl (K-major: LBO 1024 B, SBO 128 B)
      uint64_t b = desc_hi | (uint64_t)((smem_u32(sm.stage[ss]) & 0x3FFFFu) >> 4);
#pragma unroll
      for (int g = 0; g < kStageKb / kStepKb; ++g, ++step) {
        if (lane == 0) TRACE(1, step, 0);
        mbar_wait_role(&sm.a_full[ts], a_phase);
        if (lane == 0) TRACE(1, step, 1);
        const int in_seg = step & (kSegSteps - 1);
#if !APAP_TC_NODRAIN
        if (in_seg == 0 && step > 0) mbar_wait_role(&sm.d_empty, ((step / kSegSteps) - 1) & 1);   // previous segment drained
#endif
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int e = 0; e < kStepKb; ++e, b += kKbBytes >> 4) {
            const uint32_t a_hi = tmem_a + ts * 32 + e * 16, a_lo = a_hi + 8;
            const uint32_t fresh = (in_seg == 0 && e == 0) ? 0u : 1u;
            // D[0:32] (+)= hi . Ph and D[32:64] (+)= hi . Pl in one N = 64 MMA, then D[32:64] += lo . Ph:
            // the cross terms never touch the columns of the big sums
#if !APAP_TC_NOMMA
            mma_tf32_ts(tmem_d, a_hi, b, idesc64, fresh);
            mma_tf32_ts(tmem_d + kNT, a_lo, b, idesc32, 1u);
#endif
          }
          mma_commit(&sm.a_empty[ts]);             // the A columns are free when these retire
#if APAP_TC_NODRAIN
          if (step == n_step - 1) mma_commit(&sm.d_full);
#else
          if (in_seg == kSegSteps - 1 || step == n_step - 1) mma_commit(&sm.d_full);
#endif
          if (g == kStageKb / kStepKb - 1) mma_commit(&sm.smem_empty[ss]);   // ... and so is the shared-memory stage
        } else {
          b += (uint64_t)kStepKb * (kKbBytes >> 4);
        }
        __syncwarp();
        if (lane == 0) TRACE(1, step, 2);
        if (++ts == kTmemStages) { ts = 0; a_phase ^= 1; }
      }
    }
  } else {
    // ================= TMA producer (one thread) =============================================
    if (lane == 0) {
      const char *src = reinterpret_cast<const char *>(kp_blocks) + (size_t)kb0 * kKbBytes;
      for (int st = 0; st < n_stage; ++st) {
        const int ss = st % kSmemStages;
        TRACE(2, st, 0);
#ifdef APAP_TC_TRACE
        if (blockIdx.x == APAP_TC_TRACE && blockIdx.y == 0 && (st == 0 || st == n_stage - 1)) {
          g_trace[3][150 + (st ? 1 : 0)][0] = clock64(); g_trace[3][150 + (st ? 1 : 0)][1] = gtimer();
        }
#endif
        mbar_wait_role(&sm.smem_empty[ss], ((st / kSmemStages) & 1) ^ 1);   // first pass: free
        TRACE(2, st, 1);
        mbar_arrive_expect_tx(&sm.smem_full[ss], kStageBytesTc);
        bulk_g2s(sm.stage[ss], src + (size_t)st * kStageBytesTc, kStageBytesTc, &sm.smem_full[ss]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
#ifdef APAP_TC_TRACE
  if (tid == 0 && cta_lin < 8192) g_cta[cta_lin][3] = gtimer();
#endif
}

int launch_gram_tc(const float *kp_blocks, const float *anchors, int batch, int cells, int n_kp_padded, float gamma_sq,
                   float *partials, cudaStream_t st) {
  const GramPlan p = make_gram_plan(cells, n_kp_padded, sm_count_cached());
  const int n_kb = n_kp_padded / kKB;
  const int kb_per_split = p.chunks_per_split * (kChunk / kKB);
  dim3 grid((cells + 127) / 128, p.k_splits, batch);
  if (p.k_splits > 65535 || batch > 65535) return fail(APAP_E_TOOBIG, "gram: grid.y/z exceeds 65535");
  if (APAP_TC_SMEM_PAD > 0 || sizeof(TcSmem) > 48 * 1024)
    cudaFuncSetAttribute(k_gram_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TcSmem) + APAP_TC_SMEM_PAD);
  k_gram_tc<<<grid, kTcThreads, sizeof(TcSmem) + APAP_TC_SMEM_PAD, st>>>(kp_blocks, anchors, cells, p.cells_padded, n_kb, kb_per_split,
                                                      p.k_splits, gamma_sq, partials);
  return check_cuda(cudaGetLastError(), "k_gram_tc launch");
}

#ifdef APAP_TC_TRACE
extern "C" int apap_lab_trace(long long *out) {
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(g_trace));
}
extern "C" int apap_lab_cta_trace(long long *out) {
  return (int)cudaMemcpyFromSymbol(out, g_cta, sizeof(g_cta));
}
#endif

}  // namespace apap
