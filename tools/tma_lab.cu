// LAB: which 2-D tensor-map TMA load configurations run on this box (sm_100a)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tma_lab tools/tma_lab.cu ; run: tools/tma_lab
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct Params { CUtensorMap maps[3]; int c0, c1, bytes, which; unsigned *out; };

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kVariant>
__global__ void k(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap single) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 32768);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const CUtensorMap *m = kVariant == 0 ? &single : &p.maps[p.which];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(p.bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(p.c0), "r"(p.c1), "r"(s32(bar))
                 : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(s32(bar)) : "memory");
  }
  unsigned sum = 0;
  for (int i = threadIdx.x; i < p.bytes; i += blockDim.x) sum += sm[i];
  atomicAdd(p.out, sum);
}

// latency of one box load on an otherwise idle SM / GPU: globaltimer around issue -> barrier flip, `reps` times
__global__ void k_lat(const __grid_constant__ CUtensorMap m, int bytes, int c0, int c1, int reps, long long *out) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 32768);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long best = 1 << 30, sum = 0;
    for (int r = 0; r < reps; ++r) {
      unsigned long long t0, t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm)),
                   "l"(reinterpret_cast<uint64_t>(&m)), "r"(c0 + 128 * ((r * 7 + blockIdx.x) % 5)), "r"(c1 + 41 * ((r + blockIdx.x) % 13)), "r"(s32(bar))
                   : "memory");
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(r & 1) : "memory");
      }
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      const long long d = (long long)(t1 - t0);
      best = d < best ? d : best; sum += d;
    }
    out[2 * blockIdx.x] = best; out[2 * blockIdx.x + 1] = sum / reps;
  }
}

// The tile engine's load pattern: every CTA keeps two box loads in flight (ring of two stages) over never-touched
// addresses (cold) or a small set (warm), optionally with 8 small tensor stores per box like the engine's output.
__global__ void k_ring(const __grid_constant__ CUtensorMap m, const __grid_constant__ CUtensorMap om, int bytes, int tiles_x,
                       int n_tiles, int warm, int stores, long long *out) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 2 * 20480 + 12288);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + 1)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    long long sum = 0, cnt = 0, worst = 0;
    unsigned long long t_issue[2];
    auto issue = [&](int k, int tile) {
      const int s = k & 1;
      const int t = warm ? (tile % 64) : tile;
      const int c0 = (t % tiles_x) * 96, c1 = (t / tiles_x) * 31;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_issue[s]));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + s)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(sm + s * 20480)),
                   "l"(reinterpret_cast<uint64_t>(&m)), "r"(c0), "r"(c1), "r"(s32(bar + s))
                   : "memory");
    };
    int tile = blockIdx.x, k = 0;
    issue(0, tile);
    if (tile + (int)gridDim.x < n_tiles) issue(1, tile + gridDim.x);
    for (; tile < n_tiles; tile += gridDim.x, ++k) {
      const int s = k & 1;
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(ok) : "r"(s32(bar + s)), "r"((k >> 1) & 1) : "memory");
      }
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      const long long d = (long long)(t1 - t_issue[s]);
      sum += d; ++cnt; worst = d > worst ? d : worst;
      if (stores) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int w = 0; w < 8; ++w)
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&om)),
                       "r"(s32(sm + 2 * 20480 + w * 1536)), "r"((tile % tiles_x) * 96 + (w & 3) * 24), "r"((tile / tiles_x) * 32 + (w >> 2) * 16)
                       : "memory");
        asm volatile("cp.async.bulk.commit_group;\n\tcp.async.bulk.wait_group.read 1;" ::: "memory");
      }
      if (tile + 2 * (int)gridDim.x < n_tiles) issue(k + 2, tile + 2 * gridDim.x);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    out[3 * blockIdx.x] = sum / (cnt ? cnt : 1); out[3 * blockIdx.x + 1] = worst; out[3 * blockIdx.x + 2] = cnt;
  }
}

typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1, only_variant = argc > 2 ? atoi(argv[2]) : -1;
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  printf("entry point: %s q=%d ptr=%p\n", cudaGetErrorName(e), (int)q, fp);
  Enc enc = (Enc)fp;
  const int W = 1024, H = 768, pitch = W * 3;
  std::vector<uint8_t> img((size_t)pitch * H);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(1 + i % 251);
  uint8_t *d; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  unsigned *out; cudaMalloc(&out, 4);
  struct Cfg { const char *name; CUtensorMapDataType dt; int esz; int bw, bh, c0, c1; } cfgs[] = {
      {"u32 box 112x45 c0=32 (16B aligned)", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 112, 45, 32, 40},
      {"u32 box 64x80 c0=33 (4B aligned)", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 64, 80, 33, 40},
      {"u32 box 112x45 c0=-8 c1=-5", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 112, 45, -8, -5},
      {"u32 box 32x160 past the end", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 32, 160, 760, 700},
      {"u8 box 256x40 c0=48", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 40, 48, 40},
      {"u8 box 256x40 c0=33", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 40, 33, 40},
  };
  if (argc > 1 && atoi(argv[1]) == 98) {                     // the tile engine's ring of box loads, cold / warm, with / without stores
    const int W2 = 7680, H2 = 4320, pitch2 = W2 * 3, CW = 8896, CH = 4664, cpitch = CW * 3;
    uint8_t *d2, *o2; cudaMalloc(&d2, (size_t)pitch2 * H2); cudaMemset(d2, 7, (size_t)pitch2 * H2);
    cudaMalloc(&o2, (size_t)cpitch * CH);
    uint8_t *flush; cudaMalloc(&flush, 256u << 20);
    long long *lo; cudaMalloc(&lo, 8 * 3 * 1024);
    auto mk = [&](CUtensorMap *m, void *base, int pitch, int rows, int bw, int bh) {
      const cuuint64_t dims[2] = {(cuuint64_t)(pitch / 4), (cuuint64_t)rows};
      const cuuint64_t strides[1] = {(cuuint64_t)pitch};
      const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
      const cuuint32_t ones[2] = {1, 1};
      return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUtensorMap m, om;
    mk(&m, d2, pitch2, H2, 128, 40); mk(&om, o2, cpitch, CH, 24, 16);
    const int tiles_x = 60, n_tiles = 60 * 139;               // 96 u32 = 128 px per tile across, 31 rows down: the whole image once
    cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
    for (int warm = 0; warm < 2; ++warm)
      for (int stores = 0; stores < 2; ++stores) {
        cudaMemset(flush, warm + stores, 256u << 20);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_ring<<<444, 32, 60000>>>(m, om, 20480, tiles_x, n_tiles, warm, stores, lo);
        cudaEventRecord(e1);
        cudaError_t e2 = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        long long h[3 * 444]; cudaMemcpy(h, lo, sizeof(h), cudaMemcpyDeviceToHost);
        long long mean = 0, worst = 0;
        for (int i = 0; i < 444; ++i) { mean += h[3 * i]; worst = h[3 * i + 1] > worst ? h[3 * i + 1] : worst; }
        printf("ring: %s boxes, %s: %s kernel %.1f us, issue->landed mean %lld ns, worst %lld ns (%d tiles, 444 CTAs x 2 in flight)\n",
               warm ? "warm (64 positions)" : "cold (whole 8K image once)", stores ? "with 8 stores per box" : "loads only",
               cudaGetErrorName(e2), ms * 1e3, mean / 444, worst, n_tiles);
      }
    return 0;
  }
  if (argc > 1 && atoi(argv[1]) == 99) {                     // latency of box loads, alone and with every SM loading
    const int W2 = 7680, H2 = 4320, pitch2 = W2 * 3;
    uint8_t *d2; cudaMalloc(&d2, (size_t)pitch2 * H2); cudaMemset(d2, 7, (size_t)pitch2 * H2);
    long long *lo; cudaMalloc(&lo, 8 * 2 * 1024);
    struct { int bw, bh; } shapes[] = {{128, 40}, {64, 40}, {32, 40}, {128, 10}, {128, 1}};
    for (auto &sh : shapes)
      for (int grid : {1, 148, 444}) {
        const cuuint64_t dims[2] = {(cuuint64_t)(pitch2 / 4), (cuuint64_t)H2};
        const cuuint64_t strides[1] = {(cuuint64_t)pitch2};
        const cuuint32_t box[2] = {(cuuint32_t)sh.bw, (cuuint32_t)sh.bh};
        const cuuint32_t ones[2] = {1, 1};
        CUtensorMap m;
        enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d2, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaFuncSetAttribute(k_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
        k_lat<<<grid, 32, 40000>>>(m, sh.bw * 4 * sh.bh, 256, 100, 64, lo);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long h[2 * 444]; cudaMemcpy(h, lo, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
        long long mb = 1 << 30, ma = 0;
        for (int i = 0; i < grid; ++i) { mb = h[2 * i] < mb ? h[2 * i] : mb; ma += h[2 * i + 1]; }
        printf("box %4d B x %2d rows (%5d B) grid %3d: %s best %lld ns mean %lld ns\n", sh.bw * 4, sh.bh, sh.bw * 4 * sh.bh, grid,
               cudaGetErrorName(e2), mb, ma / grid);
      }
    return 0;
  }
  for (int variant = 0; variant < 2; ++variant)
    for (int ci = 0; ci < (int)(sizeof(cfgs) / sizeof(cfgs[0])); ++ci) {
      if ((only >= 0 && ci != only) || (only_variant >= 0 && variant != only_variant)) continue;
      auto &c = cfgs[ci];
      Params p; memset(&p, 0, sizeof(p));
      const cuuint64_t dims[2] = {(cuuint64_t)(pitch / c.esz), (cuuint64_t)H};
      const cuuint64_t strides[1] = {(cuuint64_t)pitch};
      const cuuint32_t box[2] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh};
      const cuuint32_t ones[2] = {1, 1};
      CUtensorMap m;
      CUresult r = enc(&m, c.dt, 2, d, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      p.maps[0] = p.maps[1] = p.maps[2] = m; p.which = 2;
      p.c0 = c.c0; p.c1 = c.c1; p.bytes = c.bw * c.esz * c.bh; p.out = out;
      cudaMemset(out, 0, 4);
      if (variant == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000); k<0><<<1, 128, 40000>>>(p, m); }
      else { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000); k<1><<<1, 128, 40000>>>(p, m); }
      e = cudaDeviceSynchronize();
      unsigned got = 0; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
      // expected sum on the host (zero outside the image)
      unsigned want = 0;
      for (int r2 = 0; r2 < c.bh; ++r2)
        for (int b = 0; b < c.bw * c.esz; ++b) {
          long long y = c.c1 + r2, xb = (long long)c.c0 * c.esz + b;
          if (y >= 0 && y < H && xb >= 0 && xb < pitch) want += img[(size_t)y * pitch + xb];
        }
      printf("variant %d (%s) %-34s encode=%d run=%s sum %u want %u %s\n", variant, variant ? "map inside struct" : "map as own param",
             c.name, (int)r, cudaGetErrorName(e), got, want, got == want ? "OK" : "MISMATCH");
      if (e != cudaSuccess) { printf("sticky error, stop\n"); return 1; }
    }
  return 0;
}
