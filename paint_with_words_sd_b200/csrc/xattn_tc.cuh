// Paint-with-Words cross-attention on tcgen05 tensor cores (sm_100a), keys T <= 80.
//
// Fused region of the reference's inj_forward (paint_with_words.py:87-118), per image b / head h / 128-row tile:
//     S = Q_h K_h^T            UMMA M=128 N=80 K=D(+pad)   operands staged by TMA, accumulator in TMEM
//     P = softmax(scale*(fp16(S) + g*M_b*w[b]))             one thread per query row, mask tile in shared memory
//     O = P V_h                UMMA M=128 N=D K=80           P written to swizzled shared memory, V MN-major
// Work unit = (image, row tile, head), flattened head-minor; every CTA owns a contiguous range of units so the
// [128 x T] fp32 mask tile of a (image, tile) group is fetched once and reused by all heads of the group.
//
// Warp roles (320 threads): warp 0 TMA producer | warp 1 UMMA issuer + TMEM owner | warps 2-5 softmax group 0 |
// warps 6-9 softmax group 1.  Group g handles iterations it % 2 == g with its own S/O TMEM buffers and P buffer.
//
// Shared-memory tiles are 128-byte-swizzled "atoms" of 64 fp16 columns: Q [128 x 64], K/V [80 x 64] rows of 128 B.
// TMA's out-of-bounds zero fill supplies every padding the UMMAs need (d >= D, token >= T, row >= N).
#pragma once
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "ptx_sm100.cuh"
#include "pww_common.cuh"

namespace pww {
namespace tc {

constexpr int kBM = 128;         // query rows per tile
constexpr int kTP = 80;          // padded key count
constexpr int kThreads = 320;        // stats kernel
constexpr int kFwdThreads = 384;     // forward kernel: 12 warps
constexpr uint32_t kQAtom = 128 * 128;   // bytes
constexpr uint32_t kKAtom = kTP * 128;
constexpr uint32_t kPAtom = 128 * 128;
constexpr uint32_t kPBuf = 2 * kPAtom;   // token columns 0-63 | 64-79
constexpr uint32_t kMaskBytes = kBM * kTP * 4;
constexpr int kTmemCols = 512;
constexpr int kMaxBatch = 256;   // images per launch (the C ABI splits larger batches)
constexpr int kMaxLocal = 4;     // images one CTA's unit range can touch: B/148 + 2 <= 4 for B <= 256
// TMEM column map (512 columns): S buffers 80 fp32 columns, P buffers 40 columns (80 packed fp16: the A operand of
// the P.V UMMA is read straight from tensor memory), O buffers up to 160 columns.  For D = 160 there is no room for
// separate P buffers: P aliases the first 40 columns of its S buffer.
template <int D> __host__ __device__ constexpr uint32_t col_s(int g) { return g ? 96u : 0u; }
template <int D> __host__ __device__ constexpr uint32_t col_p(int g) { return D == 160 ? (g ? 96u : 0u) : (g ? 232u : 192u); }
template <int D> __host__ __device__ constexpr uint32_t col_o(int g) { return D == 160 ? (g ? 352u : 192u) : (g ? 384u : 288u); }

template <int D>
struct Cfg {
  static constexpr int NA = (D + 63) / 64;          // 64-column atoms along the head dim
  static constexpr int DP = (D + 15) / 16 * 16;     // UMMA extent of the head dim
  static constexpr int KSTEPS = DP / 16;
  // Q/K tiles and V tiles live in separate rings: Q/K of a unit are released as soon as S = Q K^T is done (long
  // before P.V), so the next units' Q/K loads -- ~2 us of TMA latency for 200+ short rows -- are issued early.
  static constexpr int NQK = (D <= 64) ? 4 : (D <= 80 ? 2 : 1);
  static constexpr int NV = (D <= 64) ? 3 : 2;
  static constexpr uint32_t QKSTAGE = NA * (kQAtom + kKAtom);
  static constexpr uint32_t VSTAGE = NA * kKAtom;
  // Row sums ride on the P.V UMMA: when the last 64-column V atom has a spare column (D = 40, 80) that column is
  // set to 1.0 for every real token, so accumulator column D holds sum_j fp16(P_j) -- exactly the normaliser of
  // the P that was multiplied -- and the softmax threads never add the 80 exponentials themselves.
  static constexpr bool ONES = (D == 40 || D == 80);
  static constexpr int DPV = ONES ? (D + 16) / 16 * 16 : DP;   // UMMA N of P.V: 48, 64, 96, 160
  static constexpr uint32_t OFF_V = NQK * QKSTAGE;
  static constexpr uint32_t OFF_MASK = OFF_V + NV * VSTAGE;
  static constexpr bool P_ALIAS = (D == 160);      // P overwrites S: the next S = Q K^T must wait for P.V, not for P
  static constexpr uint32_t OFF_BAR = OFF_MASK + kMaskBytes;
  static constexpr uint32_t SMEM = OFF_BAR + 256 + 1024;   // + alignment slack
  // stats kernel: Q and K only
  static constexpr uint32_t SSTAGE = NA * (kQAtom + kKAtom);
  static constexpr int S_NSTAGE = (D <= 80) ? 3 : 2;
  static constexpr uint32_t S_OFF_BAR = S_NSTAGE * SSTAGE;
  static constexpr uint32_t S_SMEM = S_OFF_BAR + 256 + 1024;
};

// Debug timeline (test infrastructure): when TcParams::timeline is non-null, CTA 0 records clock64 per
// (tag, iteration) in shared memory and dumps the table to global memory when the kernel ends.
constexpr int kTlTags = 12, kTlIts = 40;
__shared__ long long tl_buf[kTlTags * kTlIts];
#define PWW_TL(tag, it)                                                                      \
  do {                                                                                       \
    if (tp.timeline != nullptr && blockIdx.x == 0 && (it) < kTlIts) tl_buf[(tag) * kTlIts + (it)] = clock64(); \
  } while (0)

struct TcParams {
  XattnParams x;
  int tiles;        // row tiles per image
  int units;        // B * tiles * H
  int k_batched;    // 0 when k/v have batch stride 0 (shared context)
  long long* timeline;   // debug only (see PWW_TL)
};

__device__ __forceinline__ void tl_init(const TcParams& tp) {
  if (tp.timeline != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < kTlTags * kTlIts; i += blockDim.x) tl_buf[i] = 0;
}
__device__ __forceinline__ void tl_dump(const TcParams& tp) {   // call after a __syncthreads at kernel end
  if (tp.timeline != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < kTlTags * kTlIts; i += blockDim.x) tp.timeline[i] = tl_buf[i];
}

// Walks consecutive units without a div/mod per step.
struct UnitIter {
  int b, tile, h, tiles, H;
  __device__ __forceinline__ UnitIter(int u, int tiles_, int H_) : tiles(tiles_), H(H_) {
    h = u % H_;
    const int t = u / H_;
    tile = t % tiles_;
    b = t / tiles_;
  }
  __device__ __forceinline__ void next() {
    if (++h == H) {
      h = 0;
      if (++tile == tiles) { tile = 0; ++b; }
    }
  }
};
// Forward-kernel unit order: tile-major, image-minor, head-minor.  Consecutive (image, tile) groups alternate between
// images, so every CTA's contiguous range mixes biased (mask + bias work) and unbiased units instead of some CTAs
// getting only the expensive kind (the b-major order left ~10 % of the kernel as tail imbalance).
struct UnitIterTM {
  int b, tile, h, B, H;
  __device__ __forceinline__ UnitIterTM(int u, int B_, int H_) : B(B_), H(H_) {
    h = u % H_;
    const int t = u / H_;
    b = t % B_;
    tile = t / B_;
  }
  __device__ __forceinline__ void next() {
    if (++h == H) {
      h = 0;
      if (++b == B) { b = 0; ++tile; }
    }
  }
};
__device__ __forceinline__ void cta_range(int units, int& u0, int& u1) {
  u0 = (int)((long long)blockIdx.x * units / gridDim.x);
  u1 = (int)((long long)(blockIdx.x + 1) * units / gridDim.x);
}
__device__ __forceinline__ int image_widx(const XattnParams& p, int b) {
  if (p.wmap == nullptr) return -1;
  return p.wmap_index ? p.wmap_index[b] : b;
}

// ---------------------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------------------
// TT = 77 compiles the key-length checks away (the Stable Diffusion case); TT = 0 keeps them for any T <= 80.
template <int D, int TT>
__global__ void __launch_bounds__(kFwdThreads, 1)
xattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                    const __grid_constant__ CUtensorMap tmv, const TcParams tp) {
  using C = Cfg<D>;
  const XattnParams& p = tp.x;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = smem0 + C::OFF_BAR;
  // barrier slots (8 bytes each)
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_QFULL = 0, B_QEMPTY = 4, B_VFULL = 8, B_VEMPTY = 11, B_MFULL = 14, B_MEMPTY = 15, B_SREADY = 16,
                B_PREADY = 18, B_OREADY = 20, B_OFREE = 22, B_TMEMPTR = 24;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u0, u1;
  cta_range(tp.units, u0, u1);
  const int n_it = u1 - u0;
  // per-image weight-map index and bias coefficient g(sigma)*M_b, staged once (no global loads in the loops)
  __shared__ int s_widx[kMaxBatch];
  __shared__ float s_coef[kMaxBatch];
  tl_init(tp);
  for (int b = threadIdx.x; b < p.B; b += kFwdThreads) {
    const int wi = image_widx(p, b);
    s_widx[b] = wi;
    s_coef[b] = wi >= 0 ? __ldg(p.g_sigma) * __ldg(p.stats + b) : 0.f;
  }

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
    for (int s = 0; s < C::NQK; ++s) {
      ptx::mbar_init(BAR(B_QFULL + s), 1);
      ptx::mbar_init(BAR(B_QEMPTY + s), 1);
    }
    for (int s = 0; s < C::NV; ++s) {
      ptx::mbar_init(BAR(B_VFULL + s), 1);
      ptx::mbar_init(BAR(B_VEMPTY + s), 1);
    }
    ptx::mbar_init(BAR(B_MFULL), 1);
    ptx::mbar_init(BAR(B_MEMPTY), 8);          // one elected arrive per softmax warp
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(BAR(B_SREADY + g), 1);
      ptx::mbar_init(BAR(B_PREADY + g), 4);
      ptx::mbar_init(BAR(B_OREADY + g), 1);
      ptx::mbar_init(BAR(B_OFREE + g), 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(BAR(B_TMEMPTR));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + C::OFF_BAR + 8 * B_TMEMPTR);

  if (warp == 0) {
    // ===================================== TMA producer: Q and K tiles =====================================
    if (lane == 0) {
      UnitIterTM uq(u0, p.B, p.H);
      for (int it = 0; it < n_it; ++it, uq.next()) {
        const int st = it % C::NQK;
        ptx::mbar_wait(BAR(B_QEMPTY + st), (uint32_t)(((it / C::NQK) & 1) ^ 1));
        PWW_TL(1, it);
        const uint32_t sb = smem0 + st * C::QKSTAGE;
        ptx::mbar_arrive_expect_tx(BAR(B_QFULL + st), C::QKSTAGE);
        const int kb = tp.k_batched ? uq.b : 0;
#pragma unroll
        for (int a = 0; a < C::NA; ++a) {
          ptx::tma_load_4d(sb + a * kQAtom, &tmq, BAR(B_QFULL + st), a * 64, uq.h, uq.tile * kBM, uq.b);
          ptx::tma_load_4d(sb + C::NA * kQAtom + a * kKAtom, &tmk, BAR(B_QFULL + st), a * 64, uq.h, 0, kb);
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ===================================== TMA producer: mask tiles and V tiles =====================================
    UnitIterTM uv(u0, p.B, p.H);
    int grp = -1;
    for (int it = 0; it < n_it; ++it, uv.next()) {
      if (it == 0 || uv.h == 0) {                 // first unit of an (image, tile) group: stage its mask tile
        ++grp;
        ptx::mbar_wait(BAR(B_MEMPTY), (uint32_t)((grp & 1) ^ 1));
        const int widx = s_widx[uv.b];
        if (widx >= 0) {
          const int rows = min(kBM, p.N - uv.tile * kBM);
          const uint32_t bytes = (uint32_t)rows * p.T * 4u;
          const float* src = p.wmap + (int64_t)widx * p.wmap_bs + (int64_t)uv.tile * kBM * p.T;
          if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            if (lane == 0) {
              ptx::mbar_arrive_expect_tx(BAR(B_MFULL), bytes);
              ptx::bulk_load_1d(smem0 + C::OFF_MASK, src, bytes, BAR(B_MFULL));
            }
          } else {                                // ragged tail tile: plain loads by the whole warp
            float* dst = reinterpret_cast<float*>(smem_gen + C::OFF_MASK);
            for (int i = lane; i < rows * p.T; i += 32) dst[i] = __ldg(src + i);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(BAR(B_MFULL));
          }
        } else if (lane == 0) {
          ptx::mbar_arrive(BAR(B_MFULL));         // keep the phases in lock-step for unbiased images
        }
      }
      const int st = it % C::NV;
      ptx::mbar_wait(BAR(B_VEMPTY + st), (uint32_t)(((it / C::NV) & 1) ^ 1));
      if (lane == 0) {
        const uint32_t sb = smem0 + C::OFF_V + st * C::VSTAGE;
        ptx::mbar_arrive_expect_tx(BAR(B_VFULL + st), C::VSTAGE);
        const int kb = tp.k_batched ? uv.b : 0;
#pragma unroll
        for (int a = 0; a < C::NA; ++a) ptx::tma_load_4d(sb + a * kKAtom, &tmv, BAR(B_VFULL + st), a * 64, uv.h, 0, kb);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== UMMA issuer: S = Q K^T =====================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_f16(128, kTP, false, false);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % C::NQK, g = it & 1, local = it >> 1;
        ptx::mbar_wait(BAR(B_QFULL + st), (uint32_t)((it / C::NQK) & 1));
        if (local >= 1)                                  // S[g] consumed (and, when P aliases S, P.V done with it)
          ptx::mbar_wait(BAR((C::P_ALIAS ? B_OREADY : B_PREADY) + g), (uint32_t)((local - 1) & 1));
        PWW_TL(2, it);
        ptx::tc_fence_after();
        const uint32_t sb = smem0 + st * C::QKSTAGE;
#pragma unroll
        for (int ks = 0; ks < C::KSTEPS; ++ks) {
          const uint32_t qa = sb + (ks / 4) * kQAtom + (ks % 4) * 32;
          const uint32_t ka = sb + C::NA * kQAtom + (ks / 4) * kKAtom + (ks % 4) * 32;
          ptx::umma_ss(tmem_base + col_s<D>(g), ptx::make_sw128_desc(qa, 16, 1024), ptx::make_sw128_desc(ka, 16, 1024),
                       idesc_qk, ks > 0);
        }
        ptx::umma_commit(BAR(B_SREADY + g));
        ptx::umma_commit(BAR(B_QEMPTY + st));      // Q/K tiles are dead once S exists
        PWW_TL(4, it);
      }
    }
    __syncwarp();
  } else if (warp == 11) {
    // ===================================== UMMA issuer: O = P V =====================================
    if (lane == 0) {
      constexpr uint32_t idesc_pv = ptx::make_idesc_f16(128, C::DPV, false, true);
      for (int j = 0; j < n_it; ++j) {
        const int st = j % C::NV, g = j & 1, local = j >> 1;
        ptx::mbar_wait(BAR(B_PREADY + g), (uint32_t)(local & 1));
        if (local >= 1) ptx::mbar_wait(BAR(B_OFREE + g), (uint32_t)((local - 1) & 1));
        ptx::mbar_wait(BAR(B_VFULL + st), (uint32_t)((j / C::NV) & 1));
        PWW_TL(3, j);
        ptx::tc_fence_after();
        const uint32_t vb = smem0 + C::OFF_V + st * C::VSTAGE;
#pragma unroll
        for (int ks = 0; ks < kTP / 16; ++ks)          // A = P from tensor memory: 8 columns (16 fp16) per k-step
          ptx::umma_ts(tmem_base + col_o<D>(g), tmem_base + col_p<D>(g) + ks * 8,
                       ptx::make_sw128_desc(vb + ks * 16 * 128, kKAtom, 1024), idesc_pv, ks > 0);
        ptx::umma_commit(BAR(B_OREADY + g));
        ptx::umma_commit(BAR(B_VEMPTY + st));
        PWW_TL(8, j);
      }
    }
    __syncwarp();
  } else {
    // ===================================== softmax / epilogue groups =====================================
    const int g = (warp - 2) >> 2;
    const int row = ((warp & 3) << 5) | lane;                  // TMEM lane == tile row
    const uint32_t lane_addr = (uint32_t)((warp & 3) << 5) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const float* mask_row = reinterpret_cast<const float*>(smem_gen + C::OFF_MASK) + row * (TT ? TT : p.T);
    UnitIterTM ui(u0, p.B, p.H);
    int grp = -1;
    // deferred epilogue state: iteration `pend` of this group has its P.V in flight / finished
    int pend_local = -1, pend_n = 0;
    __half* pend_out = nullptr;
    float pend_inv = 0.f;

    auto warp_arrive = [&](uint32_t bar) {       // one arrive per warp (barrier counts are per warp)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar);
    };
    auto epilogue = [&]() {                      // O[g] (fp32, TMEM) -> * 1/rowsum -> fp16 -> global
      ptx::tc_fence_after();
      const uint32_t ta = tmem_base + lane_addr + col_o<D>(g);
      float o[C::DPV];
      // whole accumulator row in as few TMEM loads as possible
      if constexpr (C::DPV == 48) { ptx::tmem_ld32_sync(ta, o); ptx::tmem_ld16_sync(ta + 32, o + 32); }
      else if constexpr (C::DPV == 64) { ptx::tmem_ld64_sync(ta, o); }
      else if constexpr (C::DPV == 96) { ptx::tmem_ld64_sync(ta, o); ptx::tmem_ld32_sync(ta + 64, o + 64); }
      else { ptx::tmem_ld64_sync(ta, o); ptx::tmem_ld64_sync(ta + 64, o + 64); ptx::tmem_ld32_sync(ta + 128, o + 128); }
      ptx::tc_fence_before();
      warp_arrive(BAR(B_OFREE + g));             // O[g] is in registers: the next P.V may overwrite it
      const float inv = C::ONES ? 1.f / o[D] : pend_inv;
      if (pend_n < p.N) {
#pragma unroll
        for (int c = 0; c < D / 8; ++c) {
          __align__(16) __half2 pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) pk[q] = __floats2half2_rn(o[c * 8 + 2 * q] * inv, o[c * 8 + 2 * q + 1] * inv);
          reinterpret_cast<uint4*>(pend_out)[c] = *reinterpret_cast<const uint4*>(pk);
        }
      }
    };

    for (int it = 0; it < n_it; ++it, ui.next()) {
      if (ui.h == 0 || it == 0) ++grp;
      const bool last_of_group = (it == n_it - 1) || (ui.h == p.H - 1);
      if ((it & 1) == g) {
        const int local = it >> 1;
        const int widx = s_widx[ui.b];
        const float coef = s_coef[ui.b];
        ptx::mbar_wait(BAR(B_SREADY + g), (uint32_t)(local & 1));
        if ((threadIdx.x & 127) == 64) PWW_TL(5, it);
        ptx::tc_fence_after();
        float s[kTP];
        ptx::tmem_ld64_sync(tmem_base + lane_addr + col_s<D>(g), s);
        ptx::tmem_ld16_sync(tmem_base + lane_addr + col_s<D>(g) + 64, s + 64);
        if ((threadIdx.x & 127) == 64) PWW_TL(0, it);
        // logits t_j = S_j + coef*w_j (unscaled), row max with 4 independent chains
        if (widx >= 0) {
          ptx::mbar_wait(BAR(B_MFULL), (uint32_t)(grp & 1));

          if constexpr (TT == 77) {
#pragma unroll
            for (int j = 0; j < 77; ++j) s[j] = fmaf(coef, mask_row[j], s[j]);
          } else {
#pragma unroll
            for (int j = 0; j < kTP; ++j)
              if (j < p.T) s[j] = fmaf(coef, mask_row[j], s[j]);
          }
        }
        if constexpr (TT == 77) {
          s[77] = s[78] = s[79] = -INFINITY;
        } else {
#pragma unroll
          for (int j = 0; j < kTP; ++j)
            if (j >= p.T) s[j] = -INFINITY;
        }
        float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
        for (int j = 4; j < kTP; j += 4) {
          m0 = fmaxf(m0, s[j]); m1 = fmaxf(m1, s[j + 1]); m2 = fmaxf(m2, s[j + 2]); m3 = fmaxf(m3, s[j + 3]);
        }
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        const float nm = -mx * sl2;
        // p_j = 2^(t_j*sl2 - mx*sl2), UNNORMALISED, packed to fp16; O is scaled by 1/rowsum in the epilogue (fp32).
        // (ex2.approx.f16x2 was tried: on sm_100a it lowers to two MUFU.EX2.F16 plus a PRMT -- no MUFU saving and
        //  40 extra instructions per row -- so the exponentials stay fp32.)
        float a0 = 0.f, a1 = 0.f;
        uint32_t pk[kTP / 2];
#pragma unroll
        for (int j = 0; j < kTP; j += 2) {
          const float e0 = ptx::ex2(fmaf(s[j], sl2, nm)), e1 = ptx::ex2(fmaf(s[j + 1], sl2, nm));
          const __half2 h = __floats2half2_rn(e0, e1);
          pk[j / 2] = *reinterpret_cast<const uint32_t*>(&h);
          if constexpr (!C::ONES) {
            const float2 f = __half22float2(h);       // sum exactly what the UMMA will multiply
            a0 += f.x; a1 += f.y;
          }
        }
        const float sum = a0 + a1;
        if constexpr (C::ONES) {
          // token r = this thread's row index: V[r][D] = 1.0 in the last V atom of this iteration's stage
          ptx::mbar_wait(BAR(B_VFULL + it % C::NV), (uint32_t)((it / C::NV) & 1));   // V tile has landed
          if (row < (TT ? TT : p.T)) {
            unsigned char* vlast = smem_gen + C::OFF_V + (it % C::NV) * C::VSTAGE + (C::NA - 1) * kKAtom;
            constexpr int cc = D % 64;             // spare column inside the last atom
            *reinterpret_cast<__half*>(vlast + row * 128 + ((((cc >> 3) ^ (row & 7))) << 4) + (cc & 7) * 2) =
                __float2half(1.0f);
          }
        }
        // the previous iteration's P.V must be done before its P buffer is overwritten
        if (lane == 0) { const int qd = warp & 3; PWW_TL((qd == 2 ? 7 : (qd == 3 ? 9 : (qd == 0 ? 10 : 11))), it); }
        if (pend_local >= 0) ptx::mbar_wait(BAR(B_OREADY + g), (uint32_t)(pend_local & 1));
        ptx::tmem_st32_u32(tmem_base + lane_addr + col_p<D>(g), pk);          // P row -> tensor memory (packed fp16)
        ptx::tmem_st8_u32(tmem_base + lane_addr + col_p<D>(g) + 32, pk + 32);
        ptx::tmem_st_wait();
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        warp_arrive(BAR(B_PREADY + g));
        if ((threadIdx.x & 127) == 64) PWW_TL(6, it);
        if (last_of_group) warp_arrive(BAR(B_MEMPTY));
        if (pend_local >= 0) epilogue();           // overlaps with this iteration's P.V

        pend_local = local;
        pend_n = ui.tile * kBM + row;
        pend_out = p.out + (int64_t)ui.b * p.o_bs + (int64_t)pend_n * p.o_rs + ui.h * D;
        pend_inv = 1.f / sum;
      } else if (last_of_group) {
        warp_arrive(BAR(B_MEMPTY));
      }
    }
    if (pend_local >= 0) {
      ptx::mbar_wait(BAR(B_OREADY + g), (uint32_t)(pend_local & 1));
      epilogue();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  tl_dump(tp);
  if (warp == 1) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// statistics kernel: per-image max / (sum, sumsq) of fp16(S) over all heads, rows and tokens
// ---------------------------------------------------------------------------------------------------------
constexpr int kStatTmemCols = 256;

template <int D, int TT>
__global__ void __launch_bounds__(kThreads, 1)
xattn_stats_tc_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                      const TcParams tp) {
  using C = Cfg<D>;
  const XattnParams& p = tp.x;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = smem0 + C::S_OFF_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_FULL = 0, B_EMPTY = 3, B_SREADY = 6, B_SFREE = 8, B_TMEMPTR = 10;
  __shared__ StatPartial s_part[8][kMaxLocal];                  // [reducer warp][local image]
  __shared__ unsigned char s_skip[kMaxBatch];
  __shared__ int is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u0, u1;
  cta_range(tp.units, u0, u1);
  const int n_it = u1 - u0;
  const int upi = tp.tiles * p.H;                               // units per image
  const int b_first = n_it > 0 ? u0 / upi : 0;
  tl_init(tp);
  for (int b = threadIdx.x; b < p.B; b += kThreads) s_skip[b] = (p.wmap_index != nullptr && p.wmap_index[b] < 0) ? 1 : 0;
  if (threadIdx.x < 8 * kMaxLocal) {
    StatPartial sp;
    sp.vmax = -INFINITY; sp.sum = 0.0; sp.sumsq = 0.0; sp.pad = 0.0;
    s_part[threadIdx.x / kMaxLocal][threadIdx.x % kMaxLocal] = sp;
  }

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    for (int s = 0; s < C::S_NSTAGE; ++s) {
      ptx::mbar_init(BAR(B_FULL + s), 1);
      ptx::mbar_init(BAR(B_EMPTY + s), 1);
    }
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(BAR(B_SREADY + g), 1);
      ptx::mbar_init(BAR(B_SFREE + g), 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kStatTmemCols>(BAR(B_TMEMPTR));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + C::S_OFF_BAR + 8 * B_TMEMPTR);

  auto skip = [&](int b) { return s_skip[b] != 0; };

  if (warp == 0) {
    int k = 0;                                                  // index among non-skipped iterations
    UnitIter un(u0, tp.tiles, p.H);
    for (int it = 0; it < n_it; ++it, un.next()) {
      if (skip(un.b)) continue;
      const int st = k % C::S_NSTAGE;
      ptx::mbar_wait(BAR(B_EMPTY + st), (uint32_t)(((k / C::S_NSTAGE) & 1) ^ 1));
      if (lane == 0) {
        PWW_TL(1, k);
        const uint32_t sb = smem0 + st * C::SSTAGE;
        ptx::mbar_arrive_expect_tx(BAR(B_FULL + st), C::SSTAGE);
        const int kb = tp.k_batched ? un.b : 0;
#pragma unroll
        for (int a = 0; a < C::NA; ++a) {
          ptx::tma_load_4d(sb + a * kQAtom, &tmq, BAR(B_FULL + st), a * 64, un.h, un.tile * kBM, un.b);
          ptx::tma_load_4d(sb + C::NA * kQAtom + a * kKAtom, &tmk, BAR(B_FULL + st), a * 64, un.h, 0, kb);
        }
      }
      __syncwarp();
      ++k;
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_f16(128, kTP, false, false);
      int k = 0;
      UnitIter un(u0, tp.tiles, p.H);
      for (int it = 0; it < n_it; ++it, un.next()) {
        if (skip(un.b)) continue;
        const int st = k % C::S_NSTAGE, g = k & 1, local = k >> 1;
        ptx::mbar_wait(BAR(B_FULL + st), (uint32_t)((k / C::S_NSTAGE) & 1));
        PWW_TL(2, k);
        if (local >= 1) ptx::mbar_wait(BAR(B_SFREE + g), (uint32_t)((local - 1) & 1));
        PWW_TL(3, k);
        ptx::tc_fence_after();
        const uint32_t sb = smem0 + st * C::SSTAGE;
#pragma unroll
        for (int ks = 0; ks < C::KSTEPS; ++ks) {
          const uint32_t qa = sb + (ks / 4) * kQAtom + (ks % 4) * 32;
          const uint32_t ka = sb + C::NA * kQAtom + (ks / 4) * kKAtom + (ks % 4) * 32;
          ptx::umma_ss(tmem_base + g * 128, ptx::make_sw128_desc(qa, 16, 1024), ptx::make_sw128_desc(ka, 16, 1024),
                       idesc_qk, ks > 0);
        }
        ptx::umma_commit(BAR(B_SREADY + g));
        ptx::umma_commit(BAR(B_EMPTY + st));
        PWW_TL(4, k);
        ++k;
      }
    }
    __syncwarp();
  } else {
    const int g = (warp - 2) >> 2;
    const int row = ((warp & 3) << 5) | lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) << 5) << 16;
    float vmax = -INFINITY, sum = 0.f, sumsq = 0.f;
    double dsum = 0.0, dsq = 0.0;                               // per-image accumulators across units
    int cur_b = -1, k = 0;
    auto flush = [&]() {
      if (cur_b < 0) return;
      double m = vmax, a = dsum + (double)sum, q = dsq + (double)sumsq;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane == 0) {
        StatPartial sp;
        sp.vmax = m; sp.sum = a; sp.sumsq = q; sp.pad = 1.0;
        s_part[warp - 2][cur_b - b_first] = sp;
      }
      vmax = -INFINITY; sum = 0.f; sumsq = 0.f; dsum = 0.0; dsq = 0.0;
    };
    UnitIter un(u0, tp.tiles, p.H);
    for (int it = 0; it < n_it; ++it, un.next()) {
      if (skip(un.b)) continue;
      if ((k & 1) == g) {
        if (un.b != cur_b) { flush(); cur_b = un.b; }
        const int local = k >> 1;
        ptx::mbar_wait(BAR(B_SREADY + g), (uint32_t)(local & 1));
        if (threadIdx.x == 64 || threadIdx.x == 192) PWW_TL(5, k);
        ptx::tc_fence_after();
        float s[kTP];
        ptx::tmem_ld64_sync(tmem_base + lane_addr + g * 128, s);
        ptx::tmem_ld16_sync(tmem_base + lane_addr + g * 128 + 64, s + 64);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(BAR(B_SFREE + g));
        if (threadIdx.x == 64 || threadIdx.x == 192) PWW_TL(6, k);
        if (un.tile * kBM + row < p.N) {
          if (p.stat == PWW_STAT_MAX) {
            // max(fp16(s)) == fp16(max(s)): rounding is monotonic, so round once at the very end
            if constexpr (TT == 77) {
              s[77] = s[78] = s[79] = -INFINITY;
            } else {
#pragma unroll
              for (int j = 0; j < kTP; ++j)
                if (j >= p.T) s[j] = -INFINITY;
            }
            float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
            for (int j = 4; j < kTP; j += 4) {
              m0 = fmaxf(m0, s[j]); m1 = fmaxf(m1, s[j + 1]); m2 = fmaxf(m2, s[j + 2]); m3 = fmaxf(m3, s[j + 3]);
            }
            vmax = fmaxf(vmax, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
          } else {
            float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int j = 0; j < kTP; j += 2) {
              // padded columns hold exact zeros (K rows >= T are zero-filled), so they add nothing
              const __half2 h = __floats2half2_rn(s[j], s[j + 1]);
              const float2 f = __half22float2(h);
              a0 += f.x; a1 += f.y;
              q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
            }
            sum = a0 + a1; sumsq = q0 + q1;
          }
        }
        dsum += (double)sum; dsq += (double)sumsq; sum = 0.f; sumsq = 0.f;
      }
      if ((threadIdx.x == 64 || threadIdx.x == 192) && ((k & 1) == g)) PWW_TL(7, k);
      ++k;
    }
    flush();
  }
  // ---- per-CTA partials (fixed warp order), then the last CTA to arrive finalises every image ----
  ptx::tc_fence_before();
  __syncthreads();
  tl_dump(tp);
  if (warp == 1) ptx::tmem_dealloc<kStatTmemCols>(tmem_base);
  const int G = gridDim.x;
  if (threadIdx.x < kMaxLocal && n_it > 0) {
    const int b = b_first + threadIdx.x;
    if (b < p.B && (int64_t)b * upi < u1 && !skip(b)) {
      StatPartial sp = s_part[0][threadIdx.x];
      for (int w = 1; w < 8; ++w) {
        sp.vmax = fmax(sp.vmax, s_part[w][threadIdx.x].vmax);
        sp.sum += s_part[w][threadIdx.x].sum;
        sp.sumsq += s_part[w][threadIdx.x].sumsq;
      }
      p.partials[(int64_t)b * G + blockIdx.x] = sp;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(p.counters, 1u);
    is_last = (prev == (unsigned)G - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // warp w finalises images w, w+10, ...: lanes stride over the CTAs whose unit range intersects the image
  for (int b = warp; b < p.B; b += kThreads / 32) {
    if (skip(b)) {
      if (lane == 0) p.stats_out[b] = 0.f;
      continue;
    }
    const long long lo = (long long)b * upi, hi = lo + upi;
    double m = -INFINITY, a = 0.0, q = 0.0;
    for (int c = lane; c < G; c += 32) {
      const long long c0 = (long long)c * tp.units / G, c1 = (long long)(c + 1) * tp.units / G;
      if (c1 > c0 && c0 < hi && c1 > lo) {
        const StatPartial* pp = p.partials + (int64_t)b * G + c;
        m = fmax(m, __ldcg(&pp->vmax));
        a += __ldcg(&pp->sum);
        q += __ldcg(&pp->sumsq);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      a += __shfl_xor_sync(0xffffffffu, a, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      const double cnt = (double)p.H * (double)p.N * (double)p.T;
      double r;
      if (p.stat == PWW_STAT_MAX) {
        r = m;
      } else {
        const double var = (q - a * a / cnt) / (cnt - 1.0);
        r = sqrt(var > 0.0 ? var : 0.0);
      }
      p.stats_out[b] = round_to_f16((float)r);
    }
  }
  if (threadIdx.x == 0) p.counters[0] = 0u;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// [B, L, H*D] fp16 viewed as (d, head, row, batch); box = 64 d x 1 head x box_rows rows x 1, 128-byte swizzle.
// Diagnostics of the last host-side failure in this file (read by pww_abi.cu into pww_last_cuda_error()).
inline char* tc_error_buf() {
  static thread_local char buf[256] = "";
  return buf;
}

inline bool make_tmap(CUtensorMap* m, const void* base, int D, int H, int L, int B, int64_t row_stride,
                      int64_t batch_stride, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    snprintf(tc_error_buf(), 256, "cuTensorMapEncodeTiled entry point not found");
    return false;
  }
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)H, (cuuint64_t)L, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)D * 2, (cuuint64_t)row_stride * 2,
                           (cuuint64_t)(batch_stride > 0 ? batch_stride : (int64_t)L * row_stride) * 2};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    snprintf(tc_error_buf(), 256,
             "cuTensorMapEncodeTiled failed (CUresult %d): base %p dims {%d,%d,%d,%d} strides {%lld,%lld,%lld} box rows %d",
             (int)r, base, D, H, L, B, (long long)strides[0], (long long)strides[1], (long long)strides[2], box_rows);
  return r == CUDA_SUCCESS;
}

inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

inline int stats_grid(int units) { return units < num_sms() ? units : num_sms(); }

inline long long*& debug_timeline() {
  static long long* ptr = nullptr;
  return ptr;
}

template <int D>
cudaError_t launch_fwd(const XattnParams& x, cudaStream_t s) {
  using C = Cfg<D>;
  CUtensorMap tq, tk, tv;
  const int kB = x.k_bs > 0 ? x.B : 1;
  if (!make_tmap(&tq, x.q, D, x.H, x.N, x.B, x.q_rs, x.q_bs, kBM) ||
      !make_tmap(&tk, x.k, D, x.H, x.T, kB, x.k_rs, x.k_bs, kTP) ||
      !make_tmap(&tv, x.v, D, x.H, x.T, kB, x.k_rs, x.k_bs, kTP))
    return cudaErrorInvalidValue;
  TcParams tp;
  tp.x = x;
  tp.tiles = ceil_div(x.N, kBM);
  tp.units = x.B * tp.tiles * x.H;
  tp.k_batched = x.k_bs > 0 ? 1 : 0;
  tp.timeline = debug_timeline();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xattn_fwd_tc_kernel<D, 77>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(xattn_fwd_tc_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = tp.units < num_sms() ? tp.units : num_sms();
  if (x.T == 77)
    xattn_fwd_tc_kernel<D, 77><<<grid, kFwdThreads, C::SMEM, s>>>(tq, tk, tv, tp);
  else
    xattn_fwd_tc_kernel<D, 0><<<grid, kFwdThreads, C::SMEM, s>>>(tq, tk, tv, tp);
  return cudaGetLastError();
}

template <int D>
cudaError_t launch_stats(const XattnParams& x, cudaStream_t s) {
  using C = Cfg<D>;
  CUtensorMap tq, tk;
  const int kB = x.k_bs > 0 ? x.B : 1;
  if (!make_tmap(&tq, x.q, D, x.H, x.N, x.B, x.q_rs, x.q_bs, kBM) ||
      !make_tmap(&tk, x.k, D, x.H, x.T, kB, x.k_rs, x.k_bs, kTP))
    return cudaErrorInvalidValue;
  TcParams tp;
  tp.x = x;
  tp.tiles = ceil_div(x.N, kBM);
  tp.units = x.B * tp.tiles * x.H;
  tp.k_batched = x.k_bs > 0 ? 1 : 0;
  tp.timeline = debug_timeline();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xattn_stats_tc_kernel<D, 77>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::S_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(xattn_stats_tc_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::S_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (x.T == 77)
    xattn_stats_tc_kernel<D, 77><<<stats_grid(tp.units), kThreads, C::S_SMEM, s>>>(tq, tk, tp);
  else
    xattn_stats_tc_kernel<D, 0><<<stats_grid(tp.units), kThreads, C::S_SMEM, s>>>(tq, tk, tp);
  return cudaGetLastError();
}

// partial slots the stats kernel writes per image
inline int stats_slots() { return num_sms(); }

}  // namespace tc
}  // namespace pww
