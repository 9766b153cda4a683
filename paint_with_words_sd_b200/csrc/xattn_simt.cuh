// Bring-up cross-attention kernels (CUDA cores, one thread per query row).
//
// First correct device path for the Paint-with-Words fused region (paint_with_words.py:87-118): used to
// get parity green on hardware and as the on-device comparator while the tcgen05 kernels in
// xattn_tc.cuh are brought up.  K_h/V_h of one head live in shared memory (broadcast reads), the query
// row lives in registers, logits of the row live in shared memory.
#pragma once
#include "pww_common.cuh"

namespace pww {
namespace simt {

constexpr int kRows = 128;  // query rows per CTA == threads per CTA

template <int D>
__device__ __forceinline__ void load_head_tile(__half* dst, const __half* src, int T, int64_t row_stride, int tid) {
  // dst: [T][D] dense; src rows are row_stride apart; D % 8 == 0 so a row is D/8 16-byte chunks.
  constexpr int kChunks = D / 8;
  for (int i = tid; i < T * kChunks; i += kRows) {
    int t = i / kChunks, c = i % kChunks;
    reinterpret_cast<uint4*>(dst)[t * kChunks + c] =
        __ldg(reinterpret_cast<const uint4*>(src + t * row_stride) + c);
  }
}

template <int D>
__device__ __forceinline__ float dot_row(const __half2 (&q)[D / 2], const __half* krow) {
  float acc = 0.f;
  const __half2* k2 = reinterpret_cast<const __half2*>(krow);
#pragma unroll
  for (int i = 0; i < D / 2; ++i) {
    float2 a = __half22float2(q[i]);
    float2 b = __half22float2(k2[i]);
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
  }
  return acc;
}

template <int D>
__device__ __forceinline__ void load_q_row(__half2 (&q)[D / 2], const __half* src, bool valid) {
  uint4* dst = reinterpret_cast<uint4*>(q);
#pragma unroll
  for (int c = 0; c < D / 8; ++c) dst[c] = valid ? __ldg(reinterpret_cast<const uint4*>(src) + c) : make_uint4(0, 0, 0, 0);
}

// grid (tiles, H, B); dynamic smem: K_h [T][D] fp16
template <int D>
__global__ void __launch_bounds__(kRows) xattn_stats_kernel(XattnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* ks = reinterpret_cast<__half*>(smem_raw);
  __shared__ double red[3][kRows / 32];
  __shared__ bool is_last;

  const int tid = threadIdx.x, tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  if (p.wmap_index != nullptr && p.wmap_index[b] < 0) {
    if (tile == 0 && h == 0 && tid == 0) p.stats_out[b] = 0.f;
    return;
  }
  load_head_tile<D>(ks, p.k + b * p.k_bs + h * D, p.T, p.k_rs, tid);
  const int n = tile * kRows + tid;
  const bool valid = n < p.N;
  __half2 q[D / 2];
  load_q_row<D>(q, p.q + b * p.q_bs + (int64_t)(valid ? n : 0) * p.q_rs + h * D, valid);
  __syncthreads();

  float vmax = -INFINITY, sum = 0.f, sumsq = 0.f;
  if (valid) {
    for (int t = 0; t < p.T; ++t) {
      float s = round_to_f16(dot_row<D>(q, ks + t * D));
      vmax = fmaxf(vmax, s);
      sum += s;
      sumsq = fmaf(s, s, sumsq);
    }
  }
  double dmax = vmax, dsum = sum, dsq = sumsq;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    dsq += __shfl_xor_sync(0xffffffffu, dsq, o);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = dmax; red[1][tid >> 5] = dsum; red[2][tid >> 5] = dsq; }
  __syncthreads();
  const int cta_in_image = h * gridDim.x + tile;
  if (tid == 0) {
    StatPartial sp;
    sp.vmax = fmax(fmax(red[0][0], red[0][1]), fmax(red[0][2], red[0][3]));
    sp.sum = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    sp.sumsq = (red[2][0] + red[2][1]) + (red[2][2] + red[2][3]);
    sp.pad = 0.0;
    p.partials[(int64_t)b * p.ctas_per_image + cta_in_image] = sp;
    __threadfence();
    unsigned prev = atomicAdd(&p.counters[b], 1u);
    is_last = (prev == (unsigned)p.ctas_per_image - 1u);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last CTA of this image: deterministic fixed-order reduction of all partials
  const StatPartial* pp = p.partials + (int64_t)b * p.ctas_per_image;
  double m = -INFINITY, s1 = 0.0, s2 = 0.0;
  for (int i = tid; i < p.ctas_per_image; i += kRows) {
    m = fmax(m, __ldcg(&pp[i].vmax));
    s1 += __ldcg(&pp[i].sum);
    s2 += __ldcg(&pp[i].sumsq);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  __syncthreads();
  if ((tid & 31) == 0) { red[0][tid >> 5] = m; red[1][tid >> 5] = s1; red[2][tid >> 5] = s2; }
  __syncthreads();
  if (tid == 0) {
    m = fmax(fmax(red[0][0], red[0][1]), fmax(red[0][2], red[0][3]));
    s1 = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    s2 = (red[2][0] + red[2][1]) + (red[2][2] + red[2][3]);
    double cnt = (double)p.H * (double)p.N * (double)p.T;
    double r;
    if (p.stat == PWW_STAT_MAX) {
      r = m;
    } else {
      double var = (s2 - s1 * s1 / cnt) / (cnt - 1.0);
      r = sqrt(var > 0.0 ? var : 0.0);
    }
    p.stats_out[b] = round_to_f16((float)r);
    p.counters[b] = 0u;
  }
}

// grid (tiles, H, B); dynamic smem: K_h [T][D], V_h [T][D] fp16, logits [kRows][T+1] fp32
template <int D>
__global__ void __launch_bounds__(kRows) xattn_fwd_kernel(XattnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int T = p.T;
  __half* ks = reinterpret_cast<__half*>(smem_raw);
  __half* vs = ks + T * D;
  float* sc = reinterpret_cast<float*>(vs + T * D);
  const int pitch = T + 1;

  const int tid = threadIdx.x, tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  load_head_tile<D>(ks, p.k + b * p.k_bs + h * D, T, p.k_rs, tid);
  load_head_tile<D>(vs, p.v + b * p.k_bs + h * D, T, p.k_rs, tid);
  const int n = tile * kRows + tid;
  const bool valid = n < p.N;
  __half2 q[D / 2];
  load_q_row<D>(q, p.q + b * p.q_bs + (int64_t)(valid ? n : 0) * p.q_rs + h * D, valid);
  __syncthreads();
  if (!valid) return;

  int widx = (p.wmap != nullptr) ? (p.wmap_index != nullptr ? p.wmap_index[b] : b) : -1;
  const float coef = (widx >= 0) ? p.g_sigma[0] * p.stats[b] : 0.f;
  const float* wrow = (widx >= 0) ? p.wmap + (int64_t)widx * p.wmap_bs + (int64_t)n * T : nullptr;
  float* my = sc + tid * pitch;

  float rmax = -INFINITY;
  for (int t = 0; t < T; ++t) {
    float s = round_to_f16(dot_row<D>(q, ks + t * D));
    float bias = wrow ? coef * __ldg(wrow + t) : 0.f;
    float l = (s + bias) * p.scale;
    my[t] = l;
    rmax = fmaxf(rmax, l);
  }
  float sum = 0.f;
  for (int t = 0; t < T; ++t) {
    float e = __expf(my[t] - rmax);
    my[t] = e;
    sum += e;
  }
  const float inv = 1.f / sum;
  float acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.f;
  for (int t = 0; t < T; ++t) {
    float pr = round_to_f16(my[t] * inv);  // reference casts P to fp16 before P@V
    const __half2* v2 = reinterpret_cast<const __half2*>(vs + t * D);
#pragma unroll
    for (int i = 0; i < D / 2; ++i) {
      float2 vv = __half22float2(v2[i]);
      acc[2 * i] = fmaf(pr, vv.x, acc[2 * i]);
      acc[2 * i + 1] = fmaf(pr, vv.y, acc[2 * i + 1]);
    }
  }
  __half* orow = p.out + b * p.o_bs + (int64_t)n * p.o_rs + h * D;
#pragma unroll
  for (int c = 0; c < D / 8; ++c) {
    __align__(16) __half2 pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pk[j] = __floats2half2_rn(acc[c * 8 + 2 * j], acc[c * 8 + 2 * j + 1]);
    reinterpret_cast<uint4*>(orow)[c] = *reinterpret_cast<const uint4*>(pk);
  }
}

inline size_t stats_smem(int T, int D) { return (size_t)T * D * sizeof(__half); }
inline size_t fwd_smem(int T, int D) { return (size_t)2 * T * D * sizeof(__half) + (size_t)kRows * (T + 1) * sizeof(float); }

}  // namespace simt
}  // namespace pww
