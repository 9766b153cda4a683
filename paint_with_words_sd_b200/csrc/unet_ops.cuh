// Memory-bound fused ops of the UNet that CALLS the attention path: GroupNorm(+per-channel add)(+SiLU)
// and GEGLU on channels-last fp16 activations.  Each replaces 3-6 eager PyTorch launches and, together with
// channels-last convolutions, removes every NCHW<->NHWC conversion from the step.  Both are HBM-bound streaming
// kernels: 16-byte vector loads/stores, one pass for statistics + one pass to apply, deterministic reductions.
#pragma once
#include "pww_common.cuh"

namespace pww {
namespace uops {

struct GnParams {
  const __half* x;      // [B, HW, C] channels-last activations
  const __half* add;    // [B, C] or nullptr: per-image per-channel value added to x BEFORE normalisation
  const __half* gamma;  // [C]
  const __half* beta;   // [C]
  __half* y;            // [B, HW, C]
  float* partial;       // [B, chunks, G, 2] (sum, sumsq) scratch
  int B, HW, C, G, chunks, rows_per_chunk, silu;
  float eps;
};

// pass 1: per (image, row chunk) partial sums per group.  block = nvec * rpp threads (nvec = C/8 vectors per row)
__global__ void gn_stats_kernel(GnParams p) {
  extern __shared__ float sm[];                 // [rpp][C][2] staging for the cross-row reduction, then [G][2]
  const int nvec = p.C >> 3;
  const int rpp = blockDim.x / nvec;
  const int v = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int r0 = chunk * p.rows_per_chunk, r1 = min(p.HW, r0 + p.rows_per_chunk);
  float s[8], q[8], a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; a[i] = 0.f; }
  if (p.add) {
    const uint4 av = *reinterpret_cast<const uint4*>(p.add + (size_t)b * p.C + v * 8);
    const __half2* ah = reinterpret_cast<const __half2*>(&av);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(ah[i]); a[2 * i] = f.x; a[2 * i + 1] = f.y; }
  }
  const __half* xb = p.x + (size_t)b * p.HW * p.C;
  for (int r = r0 + rl; r < r1; r += rpp) {
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)r * p.C) + v);
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(xh[i]);
      const float x0 = f.x + a[2 * i], x1 = f.y + a[2 * i + 1];
      s[2 * i] += x0; q[2 * i] = fmaf(x0, x0, q[2 * i]);
      s[2 * i + 1] += x1; q[2 * i + 1] = fmaf(x1, x1, q[2 * i + 1]);
    }
  }
  // per-thread channel sums -> shared [rl][c]
  float* ss = sm;
  float* sq = sm + rpp * p.C;
#pragma unroll
  for (int i = 0; i < 8; ++i) { ss[rl * p.C + v * 8 + i] = s[i]; sq[rl * p.C + v * 8 + i] = q[i]; }
  __syncthreads();
  // one thread per group reduces its channels over all row lanes in a fixed order
  const int cg = p.C / p.G;
  for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
    float ts = 0.f, tq = 0.f;
    for (int r = 0; r < rpp; ++r)
      for (int c = g * cg; c < (g + 1) * cg; ++c) { ts += ss[r * p.C + c]; tq += sq[r * p.C + c]; }
    float* out = p.partial + (((size_t)b * p.chunks + chunk) * p.G + g) * 2;
    out[0] = ts; out[1] = tq;
  }
}

// pass 2: y = act((x + add - mean) * rstd * gamma + beta).  grid (row blocks, B); block = 256
__global__ void gn_apply_kernel(GnParams p, int rows_per_block) {
  extern __shared__ float sm[];                 // scale[C], shift[C]
  float* scale = sm;
  float* shift = sm + p.C;
  __shared__ float gmean[64], grstd[64];
  const int b = blockIdx.y;
  const int cg = p.C / p.G;
  for (int g = threadIdx.x; g < p.G; g += blockDim.x) {
    double ts = 0.0, tq = 0.0;
    const float* pp = p.partial + ((size_t)b * p.chunks * p.G + g) * 2;
    for (int c = 0; c < p.chunks; ++c) { ts += pp[(size_t)c * p.G * 2]; tq += pp[(size_t)c * p.G * 2 + 1]; }
    const double n = (double)p.HW * cg;
    const double mean = ts / n;
    double var = tq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    gmean[g] = (float)mean;
    grstd[g] = (float)(1.0 / sqrt(var + (double)p.eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const int g = c / cg;
    const float sc = grstd[g] * __half2float(p.gamma[c]);
    const float ad = p.add ? __half2float(p.add[(size_t)b * p.C + c]) : 0.f;
    scale[c] = sc;
    shift[c] = __half2float(p.beta[c]) + (ad - gmean[g]) * sc;
  }
  __syncthreads();
  const int nvec = p.C >> 3;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(p.HW, r0 + rows_per_block);
  const __half* xb = p.x + (size_t)b * p.HW * p.C;
  __half* yb = p.y + (size_t)b * p.HW * p.C;
  const int total = (r1 - r0) * nvec;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int r = r0 + i / nvec, v = i % nvec;
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)r * p.C) + v);
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
    __align__(16) __half2 o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __half22float2(xh[k]);
      float y0 = fmaf(f.x, scale[v * 8 + 2 * k], shift[v * 8 + 2 * k]);
      float y1 = fmaf(f.y, scale[v * 8 + 2 * k + 1], shift[v * 8 + 2 * k + 1]);
      if (p.silu) {
        y0 = y0 / (1.f + __expf(-y0));
        y1 = y1 / (1.f + __expf(-y1));
      }
      o[k] = __floats2half2_rn(y0, y1);
    }
    reinterpret_cast<uint4*>(yb + (size_t)r * p.C)[v] = *reinterpret_cast<const uint4*>(o);
  }
}

// GEGLU: out[m, i] = in[m, i] * gelu(in[m, I + i])   (exact erf GELU, like torch.nn.functional.gelu)
__global__ void geglu_kernel(const __half* __restrict__ in, __half* __restrict__ out, long long M, int I) {
  const int nvec = I >> 3;
  const long long total = M * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / nvec;
    const int v = (int)(i % nvec);
    const uint4 av = __ldg(reinterpret_cast<const uint4*>(in + m * 2 * I) + v);
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(in + m * 2 * I + I) + v);
    const __half2* ah = reinterpret_cast<const __half2*>(&av);
    const __half2* gh = reinterpret_cast<const __half2*>(&gv);
    __align__(16) __half2 o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = __half22float2(ah[k]), g = __half22float2(gh[k]);
      const float g0 = 0.5f * g.x * (1.f + erff(g.x * 0.70710678118654752f));
      const float g1 = 0.5f * g.y * (1.f + erff(g.y * 0.70710678118654752f));
      o[k] = __floats2half2_rn(a.x * g0, a.y * g1);
    }
    reinterpret_cast<uint4*>(out + m * I)[v] = *reinterpret_cast<const uint4*>(o);
  }
}

inline int gn_chunks(int HW) {
  int rows = 64;
  int c = (HW + rows - 1) / rows;
  return c < 1 ? 1 : (c > 256 ? 256 : c);
}

}  // namespace uops
}  // namespace pww
