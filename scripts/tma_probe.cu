// TMA load-shape probe (B200): how fast can 148 CTAs stream a [B, N, H*D] fp16 tensor into shared memory with
//   mode 0  per-head boxes   (64 x 1 x 128 x 1) on the 4-D view (d, head, row, image): 128 rows of D*2 bytes, zero fill
//   mode 1  mode 0 + one cp.async.bulk.prefetch.L2 of the whole next row tile (128 rows x H*D*2 contiguous bytes)
//   mode 2  full-row boxes: ceil(H*D/64) boxes of (64 x 128) on the 2-D view (col, row): 128-byte aligned lines
//   mode 3  per-head boxes, 1-D bulk copies instead of tensor boxes are not possible (strided) -- skipped
// Each CTA walks (image, row tile) pairs in the kernel's order; per tile it loads all H heads; a ring of S stages with
// full/empty mbarriers; a consumer warp just waits for each stage and releases it (no math).  Prints us and GB/s.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lcuda -o scripts/bin/tma_probe scripts/tma_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    if (!ok && clock64() - t0 > 4000000000LL) { printf("probe: mbarrier timeout block %d\n", blockIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tma_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(p)), "r"(bytes) : "memory");
}

struct P {
  const __half* q;
  int B, N, H, D, tiles, mode, stages, atoms;   // atoms = ceil(H*D/64) (mode 2)
  uint32_t stage_bytes;
  long long* cycles;                            // per-CTA elapsed cycles
};

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm4, const __grid_constant__ CUtensorMap tm2, P p) {
  extern __shared__ unsigned char raw[];
  const uint32_t s0 = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bars = s0 + p.stages * p.stage_bytes;      // full[s] at +8s, empty[s] at +8(stages+s)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (p.stages + s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm4)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm2)) : "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  const int ntile = p.B * p.tiles;
  const int u0 = (int)((long long)blockIdx.x * ntile / gridDim.x), u1 = (int)((long long)(blockIdx.x + 1) * ntile / gridDim.x);
  // jobs: mode 0/1: one job per (tile, head) -- stage = one [128 x 128 B] atom; mode 2: one job per tile -- stage = atoms atoms
  const int jobs_per_tile = (p.mode == 2) ? 1 : p.H;
  const int njobs = (u1 - u0) * jobs_per_tile;
  if (warp == 0 && lane == 0) {
    const int64_t tile_elems = (int64_t)128 * p.H * p.D;
    if (p.mode == 1 && u0 < u1) {
      const int b = u0 / p.tiles, t = u0 % p.tiles;
      l2_prefetch(p.q + ((int64_t)b * p.N + (int64_t)t * 128) * p.H * p.D, (uint32_t)(tile_elems * 2));
    }
    for (int j = 0; j < njobs; ++j) {
      const int st = j % p.stages;
      mbar_wait(bars + 8 * (p.stages + st), (uint32_t)(((j / p.stages) & 1) ^ 1));
      const int u = u0 + j / jobs_per_tile, h = j % jobs_per_tile;
      const int b = u / p.tiles, t = u % p.tiles;
      const uint32_t dst = s0 + st * p.stage_bytes, full = bars + 8 * st;
      if (p.mode == 2) {
        mbar_expect(full, (uint32_t)p.atoms * 16384u);
        for (int a = 0; a < p.atoms; ++a) tma_2d(dst + a * 16384, &tm2, full, a * 64, b * p.N + t * 128);
      } else {
        if (p.mode == 1 && h == 0 && u + 1 < u1) {     // next row tile of this CTA -> L2, one contiguous request
          const int b2 = (u + 1) / p.tiles, t2 = (u + 1) % p.tiles;
          l2_prefetch(p.q + ((int64_t)b2 * p.N + (int64_t)t2 * 128) * p.H * p.D, (uint32_t)(tile_elems * 2));
        }
        mbar_expect(full, 16384u);
        tma_4d(dst, &tm4, full, 0, h, t * 128, b);
      }
    }
  } else if (warp == 1 && lane == 0) {
    for (int j = 0; j < njobs; ++j) {
      const int st = j % p.stages;
      mbar_wait(bars + 8 * st, (uint32_t)((j / p.stages) & 1));
      mbar_arrive(bars + 8 * (p.stages + st));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 16, N = argc > 2 ? atoi(argv[2]) : 4096, H = argc > 3 ? atoi(argv[3]) : 8,
            D = argc > 4 ? atoi(argv[4]) : 40;
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qr));
  EncodeFn enc = (EncodeFn)fnp;
  const size_t elems = (size_t)B * N * H * D;
  __half* q;
  CK(cudaMalloc(&q, elems * 2));
  CK(cudaMemset(q, 0, elems * 2));
  char* flush;
  CK(cudaMalloc(&flush, 256u << 20));
  long long* cyc;
  CK(cudaMalloc(&cyc, 148 * 8));
  CUtensorMap tm4, tm2;
  {
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)D * 2, (cuuint64_t)H * D * 2, (cuuint64_t)N * H * D * 2};
    cuuint32_t box[4] = {64, 1, 128, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm4, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, q, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode 4d failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)H * D, (cuuint64_t)B * N};
    cuuint64_t str[1] = {(cuuint64_t)H * D * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    CUresult r = enc(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, q, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode 2d failed %d\n", (int)r); return 1; }
  }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int atoms = (H * D + 63) / 64;
  printf("tensor [%d, %d, %d*%d] fp16 = %.1f MB, 148 CTAs\n", B, N, H, D, elems * 2 / 1e6);
  struct Cfg { int mode, stages; };
  std::vector<Cfg> cfgs = {{0, 2}, {0, 3}, {0, 4}, {0, 6}, {0, 8}, {0, 11}, {1, 3}, {1, 4}, {1, 8}, {2, 1}, {2, 2}};
  for (auto c : cfgs) {
    P p;
    p.q = q; p.B = B; p.N = N; p.H = H; p.D = D; p.tiles = (N + 127) / 128; p.mode = c.mode; p.stages = c.stages; p.atoms = atoms;
    p.stage_bytes = c.mode == 2 ? atoms * 16384u : 16384u;
    p.cycles = cyc;
    const size_t smem = (size_t)c.stages * p.stage_bytes + 1024 + 256;
    if (smem > 200 * 1024) continue;
    float best = 1e9f;
    for (int rep = 0; rep < 4; ++rep) {
      CK(cudaMemsetAsync(flush, rep, 256u << 20));
      CK(cudaEventRecord(e0));
      probe<<<148, 128, smem>>>(tm4, tm2, p);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    long long hc[148];
    CK(cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = hc[i] > mx ? hc[i] : mx;
    const int jobs_per_cta = (B * p.tiles * (c.mode == 2 ? 1 : H) + 147) / 148;
    printf("mode %d stages %2d: %8.2f us  %7.1f GB/s  (max CTA cycles %lld, %.0f cycles per %s job)\n", c.mode, c.stages,
           best * 1e3, elems * 2 / (best * 1e-3) / 1e9, mx, (double)mx / jobs_per_cta, c.mode == 2 ? "tile" : "head");
  }
  return 0;
}
