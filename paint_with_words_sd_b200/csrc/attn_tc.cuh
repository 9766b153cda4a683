// Self-attention (context=None in the reference's inj_forward, paint_with_words.py:71-72, 87-118) as a flash
// attention forward on tcgen05 tensor cores: softmax(scale * Q_h K_h^T) V_h with keys = the image's own N pixels.
//
// One CTA = (image, head, NQ query tiles of 128 rows).  K and V tiles of BN keys stream through two TMA rings and are
// shared by the NQ query tiles; softmax group g (4 warps, one thread per query row) owns tile g's S, P and O buffers
// in tensor memory:
//     S_g = Q_g K_j^T   (UMMA M=128 N=BN)      -> TMEM, double-buffered: S_g(j+1) is issued while softmax(j) runs
//     online softmax with LAZY rescaling: the running max only moves (and O_g is rescaled in TMEM) when the tile max
//       exceeds it by more than 2^8, so the correction pass is rare; P = 2^(s - m) <= 256 fits fp16
//     P_g (packed fp16)                         -> TMEM (double-buffered where it fits), the A operand of
//     O_g += P_g V_j    (TS-form UMMA M=128 N=D K=BN, V MN-major) accumulating in TMEM over all key tiles;
//       for D = 40 / 80 a ones column in V makes accumulator column D the row sum of the fp16 P that was multiplied
// Warp roles (384 threads): warp 0 Q+K TMA producer | warp 1 S-UMMA issuer + TMEM owner | warp 2 V TMA producer |
// warp 3 PV-UMMA issuer | warps 4-7 and 8-11 the two softmax groups.  Every barrier that takes several arrivals per
// phase is per buffer, so a warp's next arrival on it is causally behind the completion of the current phase.
// Padding (d >= D, key >= N, row >= N) comes from TMA out-of-bounds zero fill; padded keys of the last tile are masked
// to -inf.
#pragma once
#include "ptx_sm100.cuh"
#include "pww_common.cuh"
#include "xattn_tc.cuh"   // make_tmap, num_sms, tc_error_buf

namespace pww {
namespace fa {

constexpr int kThreads = 384;

template <int D>
struct Cfg {
  static constexpr int NA = (D + 63) / 64;
  static constexpr int DP = (D + 15) / 16 * 16;
  static constexpr int BN = 64;                            // keys per tile
  static constexpr int NQ = (D <= 80) ? 2 : 1;             // query tiles per CTA
  static constexpr int NK = (D <= 80) ? 4 : 3;             // K ring depth
  static constexpr int NV = (D <= 80) ? 3 : 2;             // V ring depth
  // Row sums ride on the P.V UMMA where the last V atom has a spare column (see xattn_tc.cuh): accumulator column D.
  static constexpr bool ONES = (D == 40 || D == 80);
  static constexpr int DPV = ONES ? (D + 16) / 16 * 16 : DP;   // UMMA N of P.V: 48, 64, 96, 160
  static constexpr uint32_t QATOM = 128 * 128, KATOM = BN * 128, PATOM = 128 * 128;
  static constexpr uint32_t QBYTES = NA * QATOM;           // one query tile
  static constexpr uint32_t KSTAGE = NA * KATOM, VSTAGE = NA * KATOM;
  static constexpr uint32_t OFF_K = NQ * QBYTES;
  static constexpr uint32_t OFF_V = OFF_K + NK * KSTAGE;
  static constexpr uint32_t OFF_BAR = OFF_V + NV * VSTAGE;
  static constexpr uint32_t SMEM = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int OSTRIDE = (DPV + 31) / 32 * 32;
  static constexpr int NP = (D == 80) ? 1 : 2;             // P buffers per query tile (TMEM is full at D = 80)
  static_assert(NQ * 2 * BN + NQ * NP * 32 + NQ * OSTRIDE <= 512, "TMEM budget");
  static_assert(SMEM <= 232448, "shared memory budget");
  // S is double-buffered per query tile so the next S = Q K^T is issued while the softmax of the current one runs
  __host__ __device__ static constexpr uint32_t col_s(int g, int buf) { return (g * 2 + buf) * BN; }
  // P (64 fp16 per row = 32 columns) is the A operand of the P.V UMMA, read straight from tensor memory
  // double-buffered where TMEM allows, so the softmax of tile j+1 never waits for the P.V of tile j
  __host__ __device__ static constexpr uint32_t col_p(int g, int buf) { return NQ * 2 * BN + (g * NP + buf) * 32; }
  __host__ __device__ static constexpr uint32_t col_o(int g) { return NQ * 2 * BN + NQ * NP * 32 + g * OSTRIDE; }
};

struct Params {
  __half* out;
  int B, H, N;
  int64_t o_bs, o_rs;
  float scale;
};

template <int N>
__device__ __forceinline__ void tmem_ld_row(uint32_t ta, float* v) {
  if constexpr (N == 128) { ptx::tmem_ld64_sync(ta, v); ptx::tmem_ld64_sync(ta + 64, v + 64); }
  else if constexpr (N == 64) { ptx::tmem_ld64_sync(ta, v); }
  else if constexpr (N == 48) { ptx::tmem_ld32_sync(ta, v); ptx::tmem_ld16_sync(ta + 32, v + 32); }
  else if constexpr (N == 80) { ptx::tmem_ld64_sync(ta, v); ptx::tmem_ld16_sync(ta + 64, v + 64); }
  else if constexpr (N == 96) { ptx::tmem_ld64_sync(ta, v); ptx::tmem_ld32_sync(ta + 64, v + 64); }
  else if constexpr (N == 160) { ptx::tmem_ld64_sync(ta, v); ptx::tmem_ld64_sync(ta + 64, v + 64); ptx::tmem_ld32_sync(ta + 128, v + 128); }
}
template <int N>
__device__ __forceinline__ void tmem_st_row(uint32_t ta, const float* v) {
  if constexpr (N == 64) { ptx::tmem_st64(ta, v); }
  else if constexpr (N == 48) { ptx::tmem_st32(ta, v); ptx::tmem_st16(ta + 32, v + 32); }
  else if constexpr (N == 80) { ptx::tmem_st64(ta, v); ptx::tmem_st16(ta + 64, v + 64); }
  else if constexpr (N == 96) { ptx::tmem_st64(ta, v); ptx::tmem_st32(ta + 64, v + 64); }
  else if constexpr (N == 160) { ptx::tmem_st64(ta, v); ptx::tmem_st64(ta + 64, v + 64); ptx::tmem_st32(ta + 128, v + 128); }
  ptx::tmem_st_wait();
}

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const Params p) {
  using C = Cfg<D>;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = smem0 + C::OFF_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_QFULL = 0, B_KFULL = 1, B_KEMPTY = 5, B_VFULL = 9, B_VEMPTY = 12, B_SREADY = 15, B_SFREE = 19,
                B_PREADY = 23, B_PFREE = 27, B_TMEMPTR = 31;   // SREADY/SFREE/PREADY/PFREE: [g*2 + buf]
  // PREADY is per P buffer like PFREE: a 4-arrival barrier shared by both buffers could be completed by one fast
  // warp arriving for tiles j and j+1 before a slow warp has written its rows of P(j).  With one barrier per buffer a
  // warp's next arrival on it (tile j+NP) is ordered behind PFREE, i.e. behind the P.V that consumed phase j.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // block -> (image, head, query super-tile)
  const int qtiles = (p.N + 128 * C::NQ - 1) / (128 * C::NQ);
  const int qt = blockIdx.x % qtiles;
  const int h = (blockIdx.x / qtiles) % p.H;
  const int b = blockIdx.x / (qtiles * p.H);
  const int row0 = qt * 128 * C::NQ;
  const int n_kv = (p.N + C::BN - 1) / C::BN;
  // a query tile that starts beyond N has no rows: its softmax group idles and no UMMA is issued for it
  const int nq_live = (row0 + 128 < p.N && C::NQ == 2) ? 2 : 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
    ptx::mbar_init(BAR(B_QFULL), 1);
    for (int s = 0; s < C::NK; ++s) { ptx::mbar_init(BAR(B_KFULL + s), 1); ptx::mbar_init(BAR(B_KEMPTY + s), 1); }
    for (int s = 0; s < C::NV; ++s) { ptx::mbar_init(BAR(B_VFULL + s), 1); ptx::mbar_init(BAR(B_VEMPTY + s), 1); }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(BAR(B_SREADY + i), 1);
      ptx::mbar_init(BAR(B_SFREE + i), 4);
      ptx::mbar_init(BAR(B_PFREE + i), 1);
      ptx::mbar_init(BAR(B_PREADY + i), 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<C::TMEM_COLS>(BAR(B_TMEMPTR));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + C::OFF_BAR + 8 * B_TMEMPTR);
  // (no setmaxnreg re-distribution: the two-pass softmax keeps every role under the 168-register budget)

  if (warp == 0) {
    // ------------------------------------------------ producer: Q tiles once, then the K ring
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(BAR(B_QFULL), nq_live * C::QBYTES);
      for (int g = 0; g < nq_live; ++g)
#pragma unroll
        for (int a = 0; a < C::NA; ++a)
          ptx::tma_load_4d(smem0 + g * C::QBYTES + a * C::QATOM, &tmq, BAR(B_QFULL), a * 64, h, row0 + g * 128, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % C::NK;
        ptx::mbar_wait(BAR(B_KEMPTY + st), (uint32_t)(((j / C::NK) & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(BAR(B_KFULL + st), C::KSTAGE);
#pragma unroll
        for (int a = 0; a < C::NA; ++a)
          ptx::tma_load_4d(smem0 + C::OFF_K + st * C::KSTAGE + a * C::KATOM, &tmk, BAR(B_KFULL + st), a * 64, h,
                           j * C::BN, b);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ------------------------------------------------ producer: V ring
    if (lane == 0) {
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % C::NV;
        ptx::mbar_wait(BAR(B_VEMPTY + st), (uint32_t)(((j / C::NV) & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(BAR(B_VFULL + st), C::VSTAGE);
#pragma unroll
        for (int a = 0; a < C::NA; ++a)
          ptx::tma_load_4d(smem0 + C::OFF_V + st * C::VSTAGE + a * C::KATOM, &tmv, BAR(B_VFULL + st), a * 64, h,
                           j * C::BN, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------ UMMA issuer: S_g = Q_g K_j^T
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, C::BN, false, false);
      ptx::mbar_wait(BAR(B_QFULL), 0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % C::NK;
        ptx::mbar_wait(BAR(B_KFULL + st), (uint32_t)((j / C::NK) & 1));
        const int buf = j & 1;
        for (int g = 0; g < nq_live; ++g) {
          ptx::mbar_wait(BAR(B_SFREE + g * 2 + buf), (uint32_t)(((j >> 1) & 1) ^ 1));   // softmax done with this S buffer
          ptx::tc_fence_after();
          const uint32_t qb = smem0 + g * C::QBYTES, kb = smem0 + C::OFF_K + st * C::KSTAGE;
#pragma unroll
          for (int ks = 0; ks < C::DP / 16; ++ks)
            ptx::umma_ss(tmem_base + C::col_s(g, buf),
                         ptx::make_sw128_desc(qb + (ks / 4) * C::QATOM + (ks % 4) * 32, 16, 1024),
                         ptx::make_sw128_desc(kb + (ks / 4) * C::KATOM + (ks % 4) * 32, 16, 1024), idesc, ks > 0);
          ptx::umma_commit(BAR(B_SREADY + g * 2 + buf));
        }
        ptx::umma_commit(BAR(B_KEMPTY + st));
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------------------------ UMMA issuer: O_g += P_g V_j
    // The whole warp waits for the V tile (every VFULL phase is observed by the same threads, in order) and, for
    // D = 40 / 80, sets the spare column of the last V atom to 1.0 for every real key: accumulator column D of the
    // P.V UMMA is then the row sum of the fp16 P that was multiplied.
    constexpr uint32_t idesc = ptx::make_idesc_f16(128, C::DPV, false, true);
    for (int j = 0; j < n_kv; ++j) {
      const int st = j % C::NV;
      ptx::mbar_wait(BAR(B_VFULL + st), (uint32_t)((j / C::NV) & 1));
      if constexpr (C::ONES) {
        const int valid = p.N - j * C::BN;          // keys of this tile that exist
        unsigned char* vlast = smem_gen + C::OFF_V + st * C::VSTAGE + (C::NA - 1) * C::KATOM;
        constexpr int cc = D % 64;
        for (int r = lane; r < C::BN && r < valid; r += 32)
          *reinterpret_cast<__half*>(vlast + r * 128 + ((((cc >> 3) ^ (r & 7))) << 4) + (cc & 7) * 2) = __float2half(1.0f);
        ptx::fence_proxy_async_smem();
        __syncwarp();
      }
      if (lane == 0) {
        for (int g = 0; g < nq_live; ++g) {
          const int pb = j % C::NP;
          ptx::mbar_wait(BAR(B_PREADY + g * 2 + pb), (uint32_t)((j / C::NP) & 1));
          ptx::tc_fence_after();
          const uint32_t vb = smem0 + C::OFF_V + st * C::VSTAGE;
#pragma unroll
          for (int ks = 0; ks < C::BN / 16; ++ks)
            ptx::umma_ts(tmem_base + C::col_o(g), tmem_base + C::col_p(g, pb) + ks * 8,
                         ptx::make_sw128_desc(vb + ks * 16 * 128, C::KATOM, 1024), idesc, (j > 0) || (ks > 0));
          ptx::umma_commit(BAR(B_PFREE + g * 2 + pb));    // P buffer consumed == O_g holds tiles 0..j
        }
        ptx::umma_commit(BAR(B_VEMPTY + st));
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------ softmax groups
    const int g = (warp >> 2) - 1;
    if (g < nq_live) {
      const int row = ((warp & 3) << 5) | lane;
      const uint32_t lane_addr = (uint32_t)((warp & 3) << 5) << 16;
      const float sl2 = p.scale * 1.4426950408889634f;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_kv; ++j) {
        const int buf = j & 1;
        ptx::mbar_wait(BAR(B_SREADY + g * 2 + buf), (uint32_t)((j >> 1) & 1));
        ptx::tc_fence_after();
        const uint32_t ts = tmem_base + lane_addr + C::col_s(g, buf);
        const int valid = p.N - j * C::BN;          // keys of this tile that exist (>= BN except in the last tile)
        // pass 1 over S (TMEM reads are cheap): row max of the tile
        float m_tile;
        {
          float s[64];
          ptx::tmem_ld64_sync(ts, s);
          if (valid < C::BN) {
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (c >= valid) s[c] = -INFINITY;
          }
          float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
          for (int c = 4; c < 64; c += 4) {
            m0 = fmaxf(m0, s[c]); m1 = fmaxf(m1, s[c + 1]); m2 = fmaxf(m2, s[c + 2]); m3 = fmaxf(m3, s[c + 3]);
          }
          m_tile = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        }
        // lazy rescale: move the reference max only when the tile exceeds it by more than 2^8
        const bool need = (j == 0) || ((m_tile - m_run) * sl2 > 8.0f);
        float factor = 1.f;
        if (need) {
          factor = (j == 0) ? 0.f : ptx::ex2((m_run - m_tile) * sl2);
          m_run = m_tile;
        }
        const float nm = -m_run * sl2;
        // this tile's P buffer must have been consumed by its previous P.V ...
        const int pbuf = j % C::NP;
        ptx::mbar_wait(BAR(B_PFREE + g * 2 + pbuf), (uint32_t)(((j / C::NP) & 1) ^ 1));
        // ... and O may only be rescaled once the P.V of the previous tile has landed (rare path): that is the PFREE
        // phase of the previous tile's buffer, which this warp last waited on one phase earlier (no parity aliasing)
        if (j >= 1 && __any_sync(0xffffffffu, need)) {
          ptx::mbar_wait(BAR(B_PFREE + g * 2 + (j - 1) % C::NP), (uint32_t)((((j - 1) / C::NP)) & 1));
          ptx::tc_fence_after();
          float o[C::DPV];
          tmem_ld_row<C::DPV>(tmem_base + lane_addr + C::col_o(g), o);
#pragma unroll
          for (int c = 0; c < C::DPV; ++c) o[c] *= factor;
          tmem_st_row<C::DPV>(tmem_base + lane_addr + C::col_o(g), o);
        }
        // pass 2: p = 2^((s - m) * scale * log2 e) -> fp16 -> swizzled P tile, 32 columns at a time
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < C::BN; c0 += 32) {
          float s[32];
          ptx::tmem_ld32_sync(ts + c0, s);
          if (c0 + 32 == C::BN) {                   // last read of this S buffer: hand it back to the S-UMMA issuer
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(BAR(B_SFREE + g * 2 + buf));
          }
          if (valid < C::BN) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c0 + c >= valid) s[c] = -INFINITY;
          }
          uint32_t pk[16];
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float e0 = ptx::ex2(fmaf(s[c], sl2, nm)), e1 = ptx::ex2(fmaf(s[c + 1], sl2, nm));
            const float e2 = ptx::ex2(fmaf(s[c + 2], sl2, nm)), e3 = ptx::ex2(fmaf(s[c + 3], sl2, nm));
            if constexpr (!C::ONES) { a0 += e0; a1 += e1; a2 += e2; a3 += e3; }
            const __half2 h01 = __floats2half2_rn(e0, e1), h23 = __floats2half2_rn(e2, e3);
            pk[c / 2] = *reinterpret_cast<const uint32_t*>(&h01);
            pk[c / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
          }
          ptx::tmem_st16_u32(tmem_base + lane_addr + C::col_p(g, pbuf) + c0 / 2, pk);   // 32 fp16 = 16 columns of P
        }
        ptx::tmem_st_wait();
        if constexpr (!C::ONES) l_run = l_run * factor + ((a0 + a1) + (a2 + a3));
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(BAR(B_PREADY + g * 2 + pbuf));
      }
      // epilogue: O / l -> fp16 -> global
      ptx::mbar_wait(BAR(B_PFREE + g * 2 + (n_kv - 1) % C::NP), (uint32_t)(((n_kv - 1) / C::NP) & 1));
      ptx::tc_fence_after();
      float o[C::DPV];
      tmem_ld_row<C::DPV>(tmem_base + lane_addr + C::col_o(g), o);
      const float inv = 1.f / (C::ONES ? o[D] : l_run);
      const int n = row0 + g * 128 + row;
      if (n < p.N) {
        __half* orow = p.out + (int64_t)b * p.o_bs + (int64_t)n * p.o_rs + h * D;
#pragma unroll
        for (int c = 0; c < D / 8; ++c) {
          __align__(16) __half2 hk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) hk[q] = __floats2half2_rn(o[c * 8 + 2 * q] * inv, o[c * 8 + 2 * q + 1] * inv);
          reinterpret_cast<uint4*>(orow)[c] = *reinterpret_cast<const uint4*>(hk);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int D>
cudaError_t launch(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int64_t bs, int64_t rs,
                   int64_t o_bs, int64_t o_rs, float scale, cudaStream_t s) {
  using C = Cfg<D>;
  CUtensorMap tq, tk, tv;
  if (!tc::make_tmap(&tq, q, D, H, N, B, rs, bs, 128) || !tc::make_tmap(&tk, k, D, H, N, B, rs, bs, C::BN) ||
      !tc::make_tmap(&tv, v, D, H, N, B, rs, bs, C::BN))
    return cudaErrorInvalidValue;
  Params p;
  p.out = (__half*)out; p.B = B; p.H = H; p.N = N; p.o_bs = o_bs; p.o_rs = o_rs; p.scale = scale;
  static bool attr_set[tc::kMaxDevices] = {false};
  if (!attr_set[tc::cur_device()]) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[tc::cur_device()] = true;
  }
  const int qtiles = (N + 128 * C::NQ - 1) / (128 * C::NQ);
  attn_fwd_tc_kernel<D><<<B * H * qtiles, kThreads, C::SMEM, s>>>(tq, tk, tv, p);
  return cudaGetLastError();
}

}  // namespace fa
}  // namespace pww
