// Cost of the control-path primitives of the warp-specialised kernels, measured on one warp of one CTA (B200):
// mbarrier probe of a completed phase (1 lane / 32 lanes), fence.proxy.async, tcgen05 fences, elect.sync, tcgen05.mma
// issue + commit (M=128 N=80 K=16, SS form), mbarrier arrive, named barrier, clock64 itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I paint_with_words_sd_b200/csrc -o scripts/bin/ctrl_probe scripts/ctrl_probe.cu
#include <cstdio>
#include "ptx_sm100.cuh"
using namespace pww;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ unsigned char raw[];
  const uint32_t s0 = (ptx::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t bar = s0 + 65536, bar2 = bar + 8, bar3 = bar + 16, tptr = bar + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1); ptx::mbar_init(bar2, 1); ptx::mbar_init(bar3, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(raw + (tptr - ptx::smem_u32(raw)));
  if (threadIdx.x == 0) { ptx::mbar_arrive(bar); }      // phase 0 of `bar` complete
  __syncthreads();
  constexpr int R = 256;
  int k = 0;
  auto T0 = [&]() { __syncwarp(); return clock64(); };
  if (warp == 0) {
    long long t;
    // 0: clock64 back to back
    t = T0(); for (int i = 0; i < R; ++i) { asm volatile("" ::: "memory"); (void)clock64(); } out[k++] = (clock64() - t);
    // 1: test_wait of a completed phase, all 32 lanes
    t = T0(); { unsigned acc = 0; for (int i = 0; i < R; ++i) acc += ptx::mbar_test(bar, 0); if (acc == 12345) out[63] = acc; } out[k++] = clock64() - t;
    // 2: try_wait (with suspend hint) of a completed phase, all 32 lanes
    t = T0(); { unsigned acc = 0; for (int i = 0; i < R; ++i) acc += ptx::mbar_try_wait(bar, 0); if (acc == 12345) out[63] = acc; } out[k++] = clock64() - t;
    // 3: test_wait, one lane only
    t = T0(); if (lane == 0) { unsigned acc = 0; for (int i = 0; i < R; ++i) acc += ptx::mbar_test(bar, 0); if (acc == 12345) out[63] = acc; } __syncwarp(); out[k++] = clock64() - t;
    // 4: fence.proxy.async (no outstanding stores)
    t = T0(); for (int i = 0; i < R; ++i) ptx::fence_proxy_async_smem(); out[k++] = clock64() - t;
    // 5: tcgen05.fence::after_thread_sync
    t = T0(); for (int i = 0; i < R; ++i) ptx::tc_fence_after(); out[k++] = clock64() - t;
    // 6: tcgen05.fence::before_thread_sync
    t = T0(); for (int i = 0; i < R; ++i) ptx::tc_fence_before(); out[k++] = clock64() - t;
    // 7: elect.sync
    t = T0(); { unsigned acc = 0; for (int i = 0; i < R; ++i) acc += elect_one(); if (acc == 12345) out[63] = acc; } out[k++] = clock64() - t;
    // 8: __syncwarp
    t = T0(); for (int i = 0; i < R; ++i) __syncwarp(); out[k++] = clock64() - t;
    // 9: mbarrier.arrive (lane 0) on a count-1 barrier (phases just keep completing)
    t = T0(); if (lane == 0) for (int i = 0; i < R; ++i) ptx::mbar_arrive(bar2); __syncwarp(); out[k++] = clock64() - t;
    // 10: shared-memory load (dependent chain)
    t = T0(); { unsigned a = 0; volatile unsigned* sp = reinterpret_cast<volatile unsigned*>(raw); for (int i = 0; i < R; ++i) a = sp[a & 15]; if (a == 12345) out[63] = a; } out[k++] = clock64() - t;
    // 11: fence.proxy.async with one outstanding global store before each
    t = T0(); for (int i = 0; i < R; ++i) { out[32 + lane] = i; ptx::fence_proxy_async_smem(); } out[k++] = clock64() - t;
    // 12: 3 x tcgen05.mma (SS, M128 N80 K16) + 2 commits per iteration, elected lane, descriptors uniform
    {
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, 80, false, false);
      t = T0();
      for (int i = 0; i < R; ++i) {
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 3; ++ks)
            ptx::umma_ss(tmem + (i & 3) * 80, ptx::make_sw128_desc(s0 + ks * 32, 16, 1024), ptx::make_sw128_desc(s0 + 16384 + ks * 32, 16, 1024), idesc, ks > 0);
          ptx::umma_commit(bar3);
          ptx::umma_commit(bar3);
        }
        __syncwarp();
      }
      out[k++] = clock64() - t;
    }
    // 13: the same plus a wait for the commit of the previous iteration (issue -> completion round trip)
    {
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, 80, false, false);
      // drain: bar3 got 2*R arrivals above -> parity of completed phases: (2R) phases completed, even -> current parity 0 pending
      t = T0();
      uint32_t ph = 0;
      for (int i = 0; i < R; ++i) {
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 3; ++ks)
            ptx::umma_ss(tmem + (i & 3) * 80, ptx::make_sw128_desc(s0 + ks * 32, 16, 1024), ptx::make_sw128_desc(s0 + 16384 + ks * 32, 16, 1024), idesc, ks > 0);
          ptx::umma_commit(bar3);
        }
        __syncwarp();
        ptx::mbar_wait(bar3, ph);
        ph ^= 1;
      }
      out[k++] = clock64() - t;
    }
    // 14: named barrier of 64 threads with warp 2 -- skipped (needs a partner loop)
    if (lane == 0) out[62] = k;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * 8);
  cudaMemset(d, 0, 64 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    probe<<<1, 128, 100 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  }
  long long h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"clock64", "mbar test_wait (done phase), 32 lanes", "mbar try_wait (done phase), 32 lanes", "mbar test_wait, 1 lane",
                         "fence.proxy.async", "tcgen05.fence::after", "tcgen05.fence::before", "elect.sync", "__syncwarp", "mbarrier.arrive (1 lane)",
                         "ld.shared dependent", "st.global + fence.proxy.async", "elect { 3 x tcgen05.mma + 2 commit }", "elect { 3 mma + commit } + wait completion"};
  for (int i = 0; i < (int)h[62] && i < 14; ++i) printf("%-48s %8.1f cycles per op\n", names[i], (double)h[i] / 256.0);
  return 0;
}
