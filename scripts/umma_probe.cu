// Hardware probe for two questions round 1 left open (profiles/r02_plan.md).  Standalone: one CTA, operands written
// to shared memory by the threads themselves (no TMA), results checked on the device against integer arithmetic.
//
//   A. May a tcgen05.mma accumulator (and a TS-form A operand) start at ANY 16-column aligned TMEM column, whatever N?
//      The verified kernels keep accumulators at multiples of 32 (O) or of 16 with N = 80 (S); the experimental
//      four-group kernel first had N = 48 accumulators at 368 and 464 and faulted.
//   B. May TWO warps issue tcgen05.mma chains concurrently -- each warp both an SS-form chain (A from shared memory)
//      and a TS-form chain (A from tensor memory)?  The shipped kernels issue all SS chains from one warp and all TS
//      chains from another (fine); a variant where each of two warps issued both gave wrong results.
//   C. The other two things the four-group kernel does that no verified kernel does: a 20-warp CTA whose warps 16-19 use
//      tcgen05.ld / tcgen05.st, and an fp32 tcgen05.st write-back into an accumulator that is then re-loaded.
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I paint_with_words_sd_b200/csrc \
//              -o scripts/bin/umma_probe scripts/umma_probe.cu
// run:    scripts/bin/umma_probe            (prints one line per case: ok / MISMATCH count)
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx_sm100.cuh"

using namespace pww;

constexpr int kRows = 128;     // UMMA M
constexpr int kKS = 4;         // k-steps of 16 -> K = 64 (one 128-byte swizzled atom row)
constexpr uint32_t kATile = 128 * 128;   // [128 rows x 64 fp16], 128-byte swizzle
constexpr uint32_t kBTile = 128 * 128;   // up to 128 rows (N) x 64 fp16

__host__ __device__ inline int a_val(int m, int k, int salt) { return ((m + 3 * k + salt) % 7) - 3; }
__host__ __device__ inline int b_val(int n, int k, int salt) { return ((2 * n + k + salt) % 5) - 2; }

// byte offset of element (r, c) in a K-major [rows x 64] fp16 tile, SWIZZLE_128B, 8-row groups of 1024 bytes
__device__ inline uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 3) ^ (r & 7))) << 4) + (c & 7) * 2);
}

struct Case {
  int mode;        // 0: single issuer, SS chain into column d0 | 1: single issuer, TS chain (A = fp16 pairs stored at a0)
                   // 2: two warps, each issues an SS chain then a TS chain, concurrently | 3: the same work from ONE
                   // warp (control) | 4: the shipped pattern: one warp issues both SS chains, another both TS chains
                   // 5: mode 0 in a 640-thread CTA, checked by warps 16-19 after a tcgen05.st write-back (+1) and re-load
  int n;           // UMMA N (multiple of 16)
  int d0, a0;      // accumulator column / TS-form A column of chain 0
  int d1, a1;      // same for chain 1 (modes 2, 3)
  int rounds;
};

__global__ void __launch_bounds__(640, 1) probe_kernel(Case cs, int* errors) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
  const uint32_t offA[2] = {0, kATile}, offB[2] = {2 * kATile, 2 * kATile + kBTile};
  const uint32_t bar0 = smem0 + 2 * kATile + 2 * kBTile;
  auto BAR = [&](int i) { return bar0 + 8u * i; };      // 0,1: chain done (per issuer) | 2: TMEM pointer
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(BAR(0), 1);
    ptx::mbar_init(BAR(1), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 4) ptx::tmem_alloc<512>(BAR(2));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 2 * kATile + 2 * kBTile + 16);
  const int chains = cs.mode >= 2 ? 2 : 1;
  const int K = 16 * kKS;
  int bad = 0;
  for (int round = 0; round < cs.rounds; ++round) {
    // ---- operands: chain c uses salt = 11*c + round
    for (int c = 0; c < chains; ++c)
      for (int i = threadIdx.x; i < kRows * K; i += blockDim.x) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<__half*>(smem + offA[c] + sw128_off(r, k)) = __int2half_rn(a_val(r, k, 11 * c + round));
        if (r < cs.n) *reinterpret_cast<__half*>(smem + offB[c] + sw128_off(r, k)) = __int2half_rn(b_val(r, k, 11 * c + round));
      }
    // TS-form A operand: row = lane, 32-bit column j holds K elements 2j, 2j+1 (K/2 columns)
    if (warp < 4 && cs.mode >= 1 && cs.mode <= 4) {
      const int row = warp * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
      for (int c = 0; c < chains; ++c) {
        uint32_t pk[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const __half2 h = __floats2half2_rn((float)a_val(row, 2 * j, 11 * c + round + 5), (float)a_val(row, 2 * j + 1, 11 * c + round + 5));
          pk[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        ptx::tmem_st32_u32(tmem + lane_addr + (c ? cs.a1 : cs.a0), pk);
      }
      ptx::tmem_st_wait();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    // ---- issue
    const uint32_t idesc = ptx::make_idesc_f16(128, cs.n, false, false);
    auto ss_chain = [&](int c, uint32_t dcol) {
      for (int ks = 0; ks < kKS; ++ks)
        ptx::umma_ss(tmem + dcol, ptx::make_sw128_desc(smem0 + offA[c] + ks * 32, 16, 1024),
                     ptx::make_sw128_desc(smem0 + offB[c] + ks * 32, 16, 1024), idesc, ks > 0);
    };
    auto ts_chain = [&](int c, uint32_t dcol, uint32_t acol) {
      for (int ks = 0; ks < kKS; ++ks)
        ptx::umma_ts(tmem + dcol, tmem + acol + ks * 8, ptx::make_sw128_desc(smem0 + offB[c] + ks * 32, 16, 1024), idesc, ks > 0);
    };
    if ((cs.mode == 0 || cs.mode == 5) && warp == 4 && lane == 0) { ss_chain(0, cs.d0); ptx::umma_commit(BAR(0)); }
    if (cs.mode == 1 && warp == 4 && lane == 0) { ts_chain(0, cs.d0, cs.a0); ptx::umma_commit(BAR(0)); }
    if (cs.mode == 2 && (warp == 4 || warp == 5) && lane == 0) {
      // each warp: SS chain into its first accumulator, TS chain into its second one (d + n rounded up to 16 apart)
      const int c = warp - 4;
      const uint32_t d = c ? cs.d1 : cs.d0, a = c ? cs.a1 : cs.a0;
      ss_chain(c, d);
      ts_chain(c, d + 128, a);
      ptx::umma_commit(BAR(c));
    }
    if (cs.mode == 4 && warp == 4 && lane == 0) { ss_chain(0, cs.d0); ss_chain(1, cs.d1); ptx::umma_commit(BAR(0)); }
    if (cs.mode == 4 && warp == 5 && lane == 0) {
      ts_chain(0, cs.d0 + 128, cs.a0); ts_chain(1, cs.d1 + 128, cs.a1);
      ptx::umma_commit(BAR(1));
    }
    if (cs.mode == 3 && warp == 4 && lane == 0) {           // same work, one issuer: the control for mode 2
      ss_chain(0, cs.d0); ts_chain(0, cs.d0 + 128, cs.a0);
      ss_chain(1, cs.d1); ts_chain(1, cs.d1 + 128, cs.a1);
      ptx::umma_commit(BAR(0));
      ptx::umma_commit(BAR(1));
    }
    // ---- check
    const int rbase = cs.mode == 5 ? 16 : 0;       // which four warps read the accumulators back
    if (warp >= rbase && warp < rbase + 4) {
      const int row = (warp - rbase) * 32 + lane;
      const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
      for (int c = 0; c < chains; ++c) ptx::mbar_wait(BAR(c), (uint32_t)(round & 1));
      ptx::tc_fence_after();
      for (int c = 0; c < chains; ++c) {
        const int nacc = cs.mode >= 2 ? 2 : 1;
        for (int which = 0; which < nacc; ++which) {
          const bool ts = (cs.mode == 1) || (cs.mode >= 2 && which == 1);
          const uint32_t dcol = (c ? cs.d1 : cs.d0) + (which ? 128 : 0);
          for (int n0 = 0; n0 < cs.n; n0 += 16) {
            float v[16];
            ptx::tmem_ld16_sync(tmem + lane_addr + dcol + n0, v);
            if (cs.mode == 5) {                      // write-back and re-load, like pass 1 -> pass 2 of the softmax
              for (int j = 0; j < 16; ++j) v[j] += 1.0f;
              ptx::tmem_st16(tmem + lane_addr + dcol + n0, v);
              ptx::tmem_st_wait();
              ptx::tmem_ld16_sync(tmem + lane_addr + dcol + n0, v);
              for (int j = 0; j < 16; ++j) v[j] -= 1.0f;
            }
            for (int j = 0; j < 16; ++j) {
              int ref = 0;
              for (int k = 0; k < K; ++k)
                ref += a_val(row, k, 11 * c + round + (ts ? 5 : 0)) * b_val(n0 + j, k, 11 * c + round);
              if (v[j] != (float)ref) ++bad;
            }
          }
        }
      }
      ptx::tc_fence_before();
    }
    __syncthreads();
  }
  if (bad) atomicAdd(errors, bad);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) ptx::tmem_dealloc<512>(tmem);
}

int main() {
  int* d_err;
  cudaMalloc(&d_err, sizeof(int));
  const size_t smem = 2 * kATile + 2 * kBTile + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  Case cases[] = {
      // A: accumulator / A-operand placement
      {0, 80, 0, 0, 0, 0, 2},    {0, 80, 80, 0, 0, 0, 2},   {0, 80, 48, 0, 0, 0, 2},   {0, 80, 176, 0, 0, 0, 2},
      {0, 48, 0, 0, 0, 0, 2},    {0, 48, 128, 0, 0, 0, 2},  {0, 48, 320, 0, 0, 0, 2},  {0, 48, 368, 0, 0, 0, 2},
      {0, 48, 464, 0, 0, 0, 2},  {0, 48, 16, 0, 0, 0, 2},   {0, 64, 384, 0, 0, 0, 2},  {0, 96, 416, 0, 0, 0, 2},
      {1, 48, 0, 128, 0, 0, 2},  {1, 48, 368, 80, 0, 0, 2}, {1, 48, 128, 176, 0, 0, 2}, {1, 80, 240, 48, 0, 0, 2},
      // B: one issuer (control) vs two concurrent issuers, each doing an SS chain and a TS chain
      // (accumulators of chain c: SS at d, TS at d + 128; TS-form A operands in the gaps)
      {3, 80, 0, 208, 256, 464, 64}, {4, 80, 0, 208, 256, 464, 64}, {2, 80, 0, 208, 256, 464, 64},
      {2, 48, 0, 208, 256, 464, 64},
      // C: 20-warp CTA, accumulators read / written back / re-read by warps 16-19
      {5, 80, 48, 0, 0, 0, 4}, {5, 80, 432, 0, 0, 0, 4}, {5, 48, 384, 0, 0, 0, 4},
  };
  for (const Case& c : cases) {
    cudaMemset(d_err, 0, sizeof(int));
    probe_kernel<<<1, c.mode == 5 ? 640 : 192, smem>>>(c, d_err);
    cudaError_t e = cudaDeviceSynchronize();
    int h = -1;
    if (e == cudaSuccess) cudaMemcpy(&h, d_err, sizeof(int), cudaMemcpyDeviceToHost);
    printf("mode %d N %3d d0 %3d a0 %3d d1 %3d a1 %3d rounds %2d : %s", c.mode, c.n, c.d0, c.a0, c.d1, c.a1, c.rounds,
           e != cudaSuccess ? cudaGetErrorString(e) : (h == 0 ? "ok" : "MISMATCH"));
    if (e == cudaSuccess && h) printf(" (%d wrong values)", h);
    printf("\n");
    if (e != cudaSuccess) {      // sticky error: the context is gone, stop here
      printf("stopping: device error\n");
      return 1;
    }
  }
  return 0;
}
