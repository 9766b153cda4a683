// Launch floor of a one-CTA-per-SM cooperative kernel on B200: what a kernel shaped like xattn_fused2_kernel costs before it
// does any work.  Graph of 64 back-to-back launches, CUDA events.  Variants: empty body | + TMEM alloc/dealloc and one
// block-wide barrier | + one grid-wide barrier (release-add, acquire-poll) | + a 48 KB cold TMA-free global read per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I paint_with_words_sd_b200/csrc -o scripts/bin/launch_floor scripts/launch_floor.cu
#include <cstdio>
#include "ptx_sm100.cuh"
using namespace pww;

__global__ void __launch_bounds__(640, 1) k(int mode, unsigned* counter, const uint4* src, uint4* sink, unsigned epoch) {
  extern __shared__ unsigned char raw[];
  if (mode == 0) return;
  const uint32_t tptr = (ptx::smem_u32(raw) + 15u) & ~15u;
  if ((threadIdx.x >> 5) == 1) ptx::tmem_alloc<512>(tptr);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(raw + (tptr - ptx::smem_u32(raw)));
  if (mode >= 3) {                    // 48 KB per CTA, 16-byte loads, coalesced
    uint4 acc = make_uint4(0, 0, 0, 0);
    const uint4* s = src + (size_t)blockIdx.x * 3072;
    for (int i = threadIdx.x; i < 3072; i += 640) { uint4 v = __ldg(s + i); acc.x ^= v.x; acc.y ^= v.y; }
    if (acc.x == 0x12345678u) sink[threadIdx.x] = acc;
  }
  if (mode >= 2) {                    // grid barrier: everyone adds, everyone waits for gridDim.x * epoch
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
      unsigned v;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < gridDim.x * epoch);
    }
    __syncthreads();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 1) ptx::tmem_dealloc<512>(tmem);
}

int main() {
  unsigned* counter; uint4 *src, *sink;
  cudaMalloc(&counter, 4); cudaMalloc(&src, 148 * 3072 * 16 * 64); cudaMalloc(&sink, 640 * 16);
  cudaMemset(src, 1, 148 * 3072 * 16 * 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaStream_t s; cudaStreamCreate(&s);
  const char* names[] = {"empty body", "+ TMEM alloc/dealloc + 2 block barriers", "+ one grid barrier", "+ 48 KB cold read per CTA before the barrier"};
  for (int coop = 0; coop < 2; ++coop)
    for (int mode = 0; mode < 4; ++mode) {
      if (!coop && mode >= 2) continue;
      cudaMemsetAsync(counter, 0, 4, s);
      cudaGraph_t g; cudaGraphExec_t ge;
      cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
      for (int i = 0; i < 64; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(128); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = 220 * 1024; cfg.stream = s;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = coop;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k, mode, counter, (const uint4*)(src + (size_t)i * 148 * 3072), sink, (unsigned)(i + 1));
      }
      cudaStreamEndCapture(s, &g);
      cudaGraphInstantiate(&ge, g, 0);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemsetAsync(counter, 0, 4, s);
        cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
      }
      cudaError_t e = cudaGetLastError();
      printf("%-14s %-52s %6.2f us per launch%s\n", coop ? "cooperative" : "plain", names[mode], best * 1e3 / 64, e == cudaSuccess ? "" : " (CUDA error)");
    }
  return 0;
}
