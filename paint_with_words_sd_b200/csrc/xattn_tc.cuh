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

#include "ptx_sm100.cuh"
#include "pww_common.cuh"

namespace pww {
namespace tc {

constexpr int kBM = 128;         // query rows per tile
constexpr int kTP = 80;          // padded key count
constexpr int kThreads = 320;
constexpr uint32_t kQAtom = 128 * 128;   // bytes
constexpr uint32_t kKAtom = kTP * 128;
constexpr uint32_t kPAtom = 128 * 128;
constexpr uint32_t kPBuf = 2 * kPAtom;   // token columns 0-63 | 64-79
constexpr uint32_t kMaskBytes = kBM * kTP * 4;
constexpr int kTmemCols = 512;
// TMEM column map (512 columns): S buffers 80 wide, O buffers up to 160 wide
__host__ __device__ constexpr uint32_t col_s(int g) { return g ? 96u : 0u; }
__host__ __device__ constexpr uint32_t col_o(int g) { return g ? 352u : 192u; }

template <int D>
struct Cfg {
  static constexpr int NA = (D + 63) / 64;          // 64-column atoms along the head dim
  static constexpr int DP = (D + 15) / 16 * 16;     // UMMA extent of the head dim
  static constexpr int KSTEPS = DP / 16;
  static constexpr int NSTAGE = (D <= 64) ? 3 : 1;
  static constexpr uint32_t STAGE = NA * (kQAtom + 2 * kKAtom);
  static constexpr uint32_t OFF_P = NSTAGE * STAGE;
  static constexpr uint32_t OFF_MASK = OFF_P + 2 * kPBuf;
  static constexpr uint32_t OFF_BAR = OFF_MASK + kMaskBytes;
  static constexpr uint32_t SMEM = OFF_BAR + 256 + 1024;   // + alignment slack
  // stats kernel: Q and K only
  static constexpr uint32_t SSTAGE = NA * (kQAtom + kKAtom);
  static constexpr int S_NSTAGE = (D <= 80) ? 3 : 2;
  static constexpr uint32_t S_OFF_BAR = S_NSTAGE * SSTAGE;
  static constexpr uint32_t S_SMEM = S_OFF_BAR + 256 + 1024;
};

struct TcParams {
  XattnParams x;
  int tiles;        // row tiles per image
  int units;        // B * tiles * H
  int k_batched;    // 0 when k/v have batch stride 0 (shared context)
};

struct Unit {
  int b, tile, h;
};
__device__ __forceinline__ Unit decode_unit(int u, int tiles, int H) {
  Unit r;
  r.h = u % H;
  int t = u / H;
  r.tile = t % tiles;
  r.b = t / tiles;
  return r;
}
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
template <int D>
__global__ void __launch_bounds__(kThreads, 1)
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
  constexpr int B_FULL = 0, B_EMPTY = 3, B_MFULL = 6, B_MEMPTY = 7, B_SREADY = 8, B_PREADY = 10, B_OREADY = 12,
                B_OFREE = 14, B_TMEMPTR = 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u0, u1;
  cta_range(tp.units, u0, u1);
  const int n_it = u1 - u0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
    for (int s = 0; s < C::NSTAGE; ++s) {
      ptx::mbar_init(BAR(B_FULL + s), 1);
      ptx::mbar_init(BAR(B_EMPTY + s), 1);
    }
    ptx::mbar_init(BAR(B_MFULL), 1);
    ptx::mbar_init(BAR(B_MEMPTY), 256);
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(BAR(B_SREADY + g), 1);
      ptx::mbar_init(BAR(B_PREADY + g), 128);
      ptx::mbar_init(BAR(B_OREADY + g), 1);
      ptx::mbar_init(BAR(B_OFREE + g), 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(BAR(B_TMEMPTR));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + C::OFF_BAR + 8 * B_TMEMPTR);

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    int grp = -1, prev_bt = -1;
    for (int it = 0; it < n_it; ++it) {
      const Unit un = decode_unit(u0 + it, tp.tiles, p.H);
      const int bt = un.b * tp.tiles + un.tile;
      // Q/K/V first: the UMMA warp needs them before any softmax group can release the previous mask tile
      const int st = it % C::NSTAGE;
      ptx::mbar_wait(BAR(B_EMPTY + st), (uint32_t)(((it / C::NSTAGE) & 1) ^ 1));
      if (lane == 0) {
        const uint32_t sb = smem0 + st * C::STAGE;
        ptx::mbar_arrive_expect_tx(BAR(B_FULL + st), C::STAGE);
        const int kb = tp.k_batched ? un.b : 0;
#pragma unroll
        for (int a = 0; a < C::NA; ++a) {
          ptx::tma_load_4d(sb + a * kQAtom, &tmq, BAR(B_FULL + st), a * 64, un.h, un.tile * kBM, un.b);
          ptx::tma_load_4d(sb + C::NA * kQAtom + a * kKAtom, &tmk, BAR(B_FULL + st), a * 64, un.h, 0, kb);
          ptx::tma_load_4d(sb + C::NA * (kQAtom + kKAtom) + a * kKAtom, &tmv, BAR(B_FULL + st), a * 64, un.h, 0, kb);
        }
      }
      __syncwarp();
      if (bt != prev_bt) {                      // new (image, tile) group: stage its mask tile (single buffer)
        prev_bt = bt;
        ++grp;
        ptx::mbar_wait(BAR(B_MEMPTY), (uint32_t)((grp & 1) ^ 1));
        const int widx = image_widx(p, un.b);
        if (widx >= 0) {
          const int rows = min(kBM, p.N - un.tile * kBM);
          const uint32_t bytes = (uint32_t)rows * p.T * 4u;
          const float* src = p.wmap + (int64_t)widx * p.wmap_bs + (int64_t)un.tile * kBM * p.T;
          if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            if (lane == 0) {
              ptx::mbar_arrive_expect_tx(BAR(B_MFULL), bytes);
              ptx::bulk_load_1d(smem0 + C::OFF_MASK, src, bytes, BAR(B_MFULL));
            }
          } else {                              // ragged tail tile: plain loads by the whole warp
            float* dst = reinterpret_cast<float*>(smem_gen + C::OFF_MASK);
            for (int i = lane; i < rows * p.T; i += 32) dst[i] = __ldg(src + i);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(BAR(B_MFULL));
          }
        } else if (lane == 0) {
          ptx::mbar_arrive(BAR(B_MFULL));       // keep the phases in lock-step for unbiased images
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== UMMA issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_f16(128, kTP, false, false);
      constexpr uint32_t idesc_pv = ptx::make_idesc_f16(128, C::DP, false, true);
      auto issue_qk = [&](int it) {
        const int st = it % C::NSTAGE, g = it & 1;
        ptx::mbar_wait(BAR(B_FULL + st), (uint32_t)((it / C::NSTAGE) & 1));
        ptx::tc_fence_after();
        const uint32_t sb = smem0 + st * C::STAGE;
#pragma unroll
        for (int ks = 0; ks < C::KSTEPS; ++ks) {
          const uint32_t qa = sb + (ks / 4) * kQAtom + (ks % 4) * 32;
          const uint32_t ka = sb + C::NA * kQAtom + (ks / 4) * kKAtom + (ks % 4) * 32;
          ptx::umma_ss(tmem_base + col_s(g), ptx::make_sw128_desc(qa, 16, 1024), ptx::make_sw128_desc(ka, 16, 1024),
                       idesc_qk, ks > 0);
        }
        ptx::umma_commit(BAR(B_SREADY + g));
      };
      auto issue_pv = [&](int j) {
        const int st = j % C::NSTAGE, g = j & 1, local = j >> 1;
        ptx::mbar_wait(BAR(B_PREADY + g), (uint32_t)(local & 1));
        if (local >= 1) ptx::mbar_wait(BAR(B_OFREE + g), (uint32_t)((local - 1) & 1));
        ptx::tc_fence_after();
        const uint32_t sb = smem0 + st * C::STAGE;
        const uint32_t vb = sb + C::NA * (kQAtom + kKAtom);
        const uint32_t pb = smem0 + C::OFF_P + g * kPBuf;
#pragma unroll
        for (int ks = 0; ks < kTP / 16; ++ks) {
          const uint32_t pa = pb + (ks / 4) * kPAtom + (ks % 4) * 32;
          const uint32_t va = vb + ks * 16 * 128;
          ptx::umma_ss(tmem_base + col_o(g), ptx::make_sw128_desc(pa, 16, 1024),
                       ptx::make_sw128_desc(va, kKAtom, 1024), idesc_pv, ks > 0);
        }
        ptx::umma_commit(BAR(B_OREADY + g));
        ptx::umma_commit(BAR(B_EMPTY + st));
      };
      for (int it = 0; it <= n_it; ++it) {
        if (C::NSTAGE >= 2) {
          if (it < n_it) issue_qk(it);
          if (it >= 1) issue_pv(it - 1);
        } else {
          if (it >= 1) issue_pv(it - 1);
          if (it < n_it) issue_qk(it);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================== softmax / epilogue groups =====================================
    const int g = (warp - 2) >> 2;
    const int row = ((warp & 3) << 5) | lane;                  // TMEM lane == tile row
    const uint32_t lane_addr = (uint32_t)((warp & 3) << 5) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const float* mask_row = reinterpret_cast<const float*>(smem_gen + C::OFF_MASK) + row * p.T;
    int grp = -1, prev_bt = -1;
    for (int it = 0; it < n_it; ++it) {
      const Unit un = decode_unit(u0 + it, tp.tiles, p.H);
      const int bt = un.b * tp.tiles + un.tile;
      if (bt != prev_bt) { prev_bt = bt; ++grp; }
      const bool last_of_group =
          (it == n_it - 1) || (decode_unit(u0 + it + 1, tp.tiles, p.H).tile != un.tile) ||
          (decode_unit(u0 + it + 1, tp.tiles, p.H).b != un.b);
      if ((it & 1) == g) {
        const int local = it >> 1;
        const int widx = image_widx(p, un.b);
        float coef = 0.f;
        if (widx >= 0) coef = __ldg(p.g_sigma) * __ldg(p.stats + un.b);
        ptx::mbar_wait(BAR(B_SREADY + g), (uint32_t)(local & 1));
        ptx::tc_fence_after();
        float s[kTP];
        ptx::tmem_ld64_sync(tmem_base + lane_addr + col_s(g), s);
        ptx::tmem_ld16_sync(tmem_base + lane_addr + col_s(g) + 64, s + 64);
        if (widx >= 0) ptx::mbar_wait(BAR(B_MFULL), (uint32_t)(grp & 1));
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kTP; ++j) {
          float x = -INFINITY;
          if (j < p.T) {
            float sv = round_to_f16(s[j]);                     // the reference's matmul output is fp16
            if (widx >= 0) sv = fmaf(coef, mask_row[j], sv);
            x = sv * sl2;
          }
          s[j] = x;
          mx = fmaxf(mx, x);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kTP; ++j) {
          s[j] = ptx::ex2(s[j] - mx);
          sum += s[j];
        }
        const float inv = 1.f / sum;
        // P (fp16, normalised like the reference) -> swizzled K-major tile
        unsigned char* pbuf = smem_gen + C::OFF_P + g * kPBuf + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
        for (int c = 0; c < kTP / 8; ++c) {
          __align__(16) __half2 pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) pk[q] = __floats2half2_rn(s[c * 8 + 2 * q] * inv, s[c * 8 + 2 * q + 1] * inv);
          *reinterpret_cast<uint4*>(pbuf + (c >> 3) * kPAtom + (((c & 7) ^ (row & 7)) << 4)) =
              *reinterpret_cast<const uint4*>(pk);
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        ptx::mbar_arrive(BAR(B_PREADY + g));
        // epilogue of this iteration
        ptx::mbar_wait(BAR(B_OREADY + g), (uint32_t)(local & 1));
        ptx::tc_fence_after();
        const int n = un.tile * kBM + row;
        __half* orow = p.out + (int64_t)un.b * p.o_bs + (int64_t)n * p.o_rs + un.h * D;
#pragma unroll
        for (int c = 0; c < D / 8; ++c) {
          float o[8];
          ptx::tmem_ld8_sync(tmem_base + lane_addr + col_o(g) + c * 8, o);
          __align__(16) __half2 pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) pk[q] = __floats2half2_rn(o[2 * q], o[2 * q + 1]);
          if (n < p.N) reinterpret_cast<uint4*>(orow)[c] = *reinterpret_cast<const uint4*>(pk);
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(BAR(B_OFREE + g));
      }
      if (last_of_group) ptx::mbar_arrive(BAR(B_MEMPTY));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// statistics kernel: per-image max / (sum, sumsq) of fp16(S) over all heads, rows and tokens
// ---------------------------------------------------------------------------------------------------------
constexpr int kStatTmemCols = 256;

template <int D>
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
  __shared__ double red[3][kThreads / 32];
  __shared__ int is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u0, u1;
  cta_range(tp.units, u0, u1);
  const int n_it = u1 - u0;
  const int slots = gridDim.x * 8;                             // partial slots per image: (cta, softmax warp)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    for (int s = 0; s < C::S_NSTAGE; ++s) {
      ptx::mbar_init(BAR(B_FULL + s), 1);
      ptx::mbar_init(BAR(B_EMPTY + s), 1);
    }
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(BAR(B_SREADY + g), 1);
      ptx::mbar_init(BAR(B_SFREE + g), 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kStatTmemCols>(BAR(B_TMEMPTR));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + C::S_OFF_BAR + 8 * B_TMEMPTR);

  auto skip = [&](int b) { return p.wmap_index != nullptr && p.wmap_index[b] < 0; };

  if (warp == 0) {
    int k = 0;                                                  // index among non-skipped iterations
    for (int it = 0; it < n_it; ++it) {
      const Unit un = decode_unit(u0 + it, tp.tiles, p.H);
      if (skip(un.b)) continue;
      const int st = k % C::S_NSTAGE;
      ptx::mbar_wait(BAR(B_EMPTY + st), (uint32_t)(((k / C::S_NSTAGE) & 1) ^ 1));
      if (lane == 0) {
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
      for (int it = 0; it < n_it; ++it) {
        const Unit un = decode_unit(u0 + it, tp.tiles, p.H);
        if (skip(un.b)) continue;
        const int st = k % C::S_NSTAGE, g = k & 1, local = k >> 1;
        ptx::mbar_wait(BAR(B_FULL + st), (uint32_t)((k / C::S_NSTAGE) & 1));
        if (local >= 1) ptx::mbar_wait(BAR(B_SFREE + g), (uint32_t)((local - 1) & 1));
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
        ++k;
      }
    }
    __syncwarp();
  } else {
    const int g = (warp - 2) >> 2;
    const int row = ((warp & 3) << 5) | lane;
    const uint32_t lane_addr = (uint32_t)((warp & 3) << 5) << 16;
    const int slot = blockIdx.x * 8 + (warp - 2);
    // neutral partials for every image first; images this warp touches are overwritten below (same thread)
    if (lane == 0) {
      for (int b = 0; b < p.B; ++b) {
        StatPartial sp;
        sp.vmax = -INFINITY; sp.sum = 0.0; sp.sumsq = 0.0; sp.pad = 0.0;
        p.partials[(int64_t)b * slots + slot] = sp;
      }
    }
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
        p.partials[(int64_t)cur_b * slots + slot] = sp;
      }
      vmax = -INFINITY; sum = 0.f; sumsq = 0.f; dsum = 0.0; dsq = 0.0;
    };
    for (int it = 0; it < n_it; ++it) {
      const Unit un = decode_unit(u0 + it, tp.tiles, p.H);
      if (skip(un.b)) continue;
      if ((k & 1) == g) {
        if (un.b != cur_b) { flush(); cur_b = un.b; }
        const int local = k >> 1;
        ptx::mbar_wait(BAR(B_SREADY + g), (uint32_t)(local & 1));
        ptx::tc_fence_after();
        float s[kTP];
        ptx::tmem_ld64_sync(tmem_base + lane_addr + g * 128, s);
        ptx::tmem_ld16_sync(tmem_base + lane_addr + g * 128 + 64, s + 64);
        ptx::tc_fence_before();
        ptx::mbar_arrive(BAR(B_SFREE + g));
        if (un.tile * kBM + row < p.N) {
#pragma unroll
          for (int j = 0; j < kTP; ++j) {
            if (j < p.T) {
              const float sv = round_to_f16(s[j]);
              vmax = fmaxf(vmax, sv);
              sum += sv;
              sumsq = fmaf(sv, sv, sumsq);
            }
          }
        }
        dsum += (double)sum; dsq += (double)sumsq; sum = 0.f; sumsq = 0.f;
      }
      ++k;
    }
    flush();
  }
  // ---- arrival: the last CTA reduces every image's slots in a fixed order (deterministic) ----
  __threadfence();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<kStatTmemCols>(tmem_base);
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(p.counters, 1u);
    is_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int b = 0; b < p.B; ++b) {
    if (skip(b)) {
      if (threadIdx.x == 0) p.stats_out[b] = 0.f;
      continue;
    }
    const StatPartial* pp = p.partials + (int64_t)b * slots;
    double m = -INFINITY, a = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < slots; i += kThreads) {
      m = fmax(m, __ldcg(&pp[i].vmax));
      a += __ldcg(&pp[i].sum);
      q += __ldcg(&pp[i].sumsq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      a += __shfl_xor_sync(0xffffffffu, a, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    __syncthreads();
    if (lane == 0) { red[0][warp] = m; red[1][warp] = a; red[2][warp] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
      m = -INFINITY; a = 0.0; q = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) { m = fmax(m, red[0][w]); a += red[1][w]; q += red[2][w]; }
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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xattn_fwd_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = tp.units < num_sms() ? tp.units : num_sms();
  xattn_fwd_tc_kernel<D><<<grid, kThreads, C::SMEM, s>>>(tq, tk, tv, tp);
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
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xattn_stats_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::S_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  xattn_stats_tc_kernel<D><<<stats_grid(tp.units), kThreads, C::S_SMEM, s>>>(tq, tk, tp);
  return cudaGetLastError();
}

// partial slots the stats kernel writes per image
inline int stats_slots() { return num_sms() * 8; }

}  // namespace tc
}  // namespace pww
