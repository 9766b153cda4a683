// Paint-with-Words cross-attention forward, FOUR softmax groups per CTA (head dim 40 only).
//
// EXPERIMENTAL -- built, selectable with pww_debug_set_variant(2), NOT the default path and NOT working yet: its
// first (and, for lack of GPU budget, only) hardware run in round 1 ended in cudaErrorLaunchFailure at B=2, N=4096
// (scripts/g4_try.py; a fault, not a barrier time-out).  The accumulators have since been moved to 128-column
// aligned TMEM addresses (they sat at 320 + 48 g); next step is compute-sanitizer on a GPU box
// (see profiles/r01_xattn_notes.md, "Next steps").  Same maths, operands, unit order and
// mask protocol as xattn_fwd_tc_kernel (xattn_tc.cuh); what changes is how much independent softmax work one SM holds:
//
//   * 640 threads: warps 0-3 are the helpers (0 Q/K TMA producer, 1 S = Q K^T issuer + TMEM owner, 2 V + mask TMA
//     producer, 3 O = P V issuer), warps 4-19 are four softmax groups of 128 rows.  Every SM sub-partition hosts one
//     helper and one warp of each group, so while one group waits for its UMMA hand-off the other three keep the
//     MUFU / FMA pipes busy (the two-group kernel leaves ~60 % of the issue slots idle).
//   * the softmax is a chunked two-pass (32 + 32 + 16 columns) so a thread never holds more than one chunk: < 100
//     registers per thread.  Pass 1 adds the bias (written back to tensor memory for biased units) and takes the row
//     max, pass 2 exponentiates and stores the packed fp16 P over the consumed part of the S row.
//   * TMEM (512 columns): group g owns O at columns [128 g, 128 g + 48) and S/P at [128 g + 48, 128 g + 128).  One score
//     buffer per group: the buffer is S, then P, then free when the P.V that read it completes; that completion also
//     gates the (non-deferred) epilogue, which goes through a per-warp staging tile and a TMA store.
//   * at the workload's own B=2 launch (3.5 units per CTA) every unit of a CTA runs in the first and only round.
#pragma once
#include "xattn_tc.cuh"

namespace pww {
namespace tc {
namespace g4 {

constexpr int kD = 40;
constexpr int kGMax = 4;                   // softmax groups (the kernel is also built with 2, as a bisection point)
constexpr int NQK = 3, NV = 4;
constexpr uint32_t QKSTAGE = kQAtom + kKAtom;       // one 64-column atom each (D = 40)
constexpr uint32_t VSTAGE = kKAtom;
constexpr uint32_t OFF_V = NQK * QKSTAGE;
constexpr uint32_t OFF_MASK = OFF_V + NV * VSTAGE;
constexpr uint32_t OFF_BAR = OFF_MASK + kMaskBytes;
constexpr uint32_t OFF_STG = OFF_BAR + 256;
constexpr uint32_t STG_WARP = 32 * kD * 2;
constexpr uint32_t SMEM = OFF_STG + kGMax * 4 * STG_WARP + 1024;
static_assert(SMEM + 8192 <= 232448, "shared memory budget (dynamic + static tables)");
static_assert(OFF_STG % 128 == 0 && QKSTAGE % 1024 == 0 && OFF_V % 1024 == 0, "TMA / swizzle alignment");
constexpr int DPV = 48;                    // UMMA N of P.V: 40 value columns + the ones column, padded to 16
// 128 columns per group: O first (128-column aligned, like every accumulator the verified kernels use), S / P behind it
// at a 16-column aligned offset (the verified D = 80 kernel keeps S at columns 80 and 240).
__host__ __device__ constexpr uint32_t col_o(int g) { return 128u * g; }
__host__ __device__ constexpr uint32_t col_s(int g) { return 128u * g + 48u; }
static_assert(col_s(kGMax - 1) + kTP <= 512, "TMEM budget");

template <int TT, int kG>
__global__ void __launch_bounds__(128 + kG * 128, 1)
xattn_fwd_g4_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                    const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmo,
                    const TcParams tp) {
  const XattnParams& p = tp.x;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = smem0 + OFF_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  // SREADY / PREADY / PVDONE are per group.  PREADY is a 4-arrival barrier: a warp's next arrival on it (next unit of
  // the group) is behind SREADY of that unit, which is behind PVDONE of this one, which is behind this phase.
  constexpr int B_QFULL = 0, B_QEMPTY = 3, B_VFULL = 6, B_VEMPTY = 10, B_MFULL = 14, B_MEMPTY = 15, B_SREADY = 16,
                B_PREADY = 20, B_PVDONE = 24, B_TMEMPTR = 28;
  constexpr int kThreads4 = 128 + kG * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int u0, u1;
  cta_range(tp.units, u0, u1);
  const int n_it = u1 - u0;
  __shared__ int s_widx[kMaxBatch];
  __shared__ float s_coef[kMaxBatch];
  __shared__ int s_img[kMaxBatch];             // biased images first, then unbiased (see FwdWalk)
  __shared__ int s_nb;
  for (int b = threadIdx.x; b < p.B; b += kThreads4) {
    const int wi = image_widx(p, b);
    s_widx[b] = wi;
    s_coef[b] = wi >= 0 ? __ldg(p.g_sigma) * __ldg(p.stats + b) : 0.f;
  }
  if (warp == 4) {                             // stable partition of the images by "has a weight map"
    int nb_total = 0;
    for (int base = 0; base < p.B; base += 32) {
      const int b = base + lane;
      nb_total += __popc(__ballot_sync(0xffffffffu, b < p.B && image_widx(p, b) >= 0));
    }
    int cb = 0, cu = 0;
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < p.B; base += 32) {
      const int b = base + lane;
      const bool valid = b < p.B, bi = valid && image_widx(p, b) >= 0;
      const unsigned mb = __ballot_sync(0xffffffffu, bi), mu = __ballot_sync(0xffffffffu, valid && !bi);
      if (bi) s_img[cb + __popc(mb & lt)] = b;
      else if (valid) s_img[nb_total + cu + __popc(mu & lt)] = b;
      cb += __popc(mb);
      cu += __popc(mu);
    }
    if (lane == 0) s_nb = nb_total;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
    ptx::prefetch_tmap(&tmo);
    for (int s = 0; s < NQK; ++s) {
      ptx::mbar_init(BAR(B_QFULL + s), 1);
      ptx::mbar_init(BAR(B_QEMPTY + s), 1);
    }
    for (int s = 0; s < NV; ++s) {
      ptx::mbar_init(BAR(B_VFULL + s), 1);
      ptx::mbar_init(BAR(B_VEMPTY + s), 1);
    }
    ptx::mbar_init(BAR(B_MFULL), 1);
    ptx::mbar_init(BAR(B_MEMPTY), kG * 4);     // one elected arrive per softmax warp
    for (int g = 0; g < kG; ++g) {
      ptx::mbar_init(BAR(B_SREADY + g), 1);
      ptx::mbar_init(BAR(B_PREADY + g), 4);
      ptx::mbar_init(BAR(B_PVDONE + g), 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(BAR(B_TMEMPTR));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + OFF_BAR + 8 * B_TMEMPTR);

  if (warp == 0) {
    // ===================================== TMA producer: Q and K tiles =====================================
    if (lane == 0) {
      FwdWalk wq(u0, p.B, p.H, s_nb, s_img);
      for (int it = 0; it < n_it; ++it, wq.next()) {
        const FwdUnit uq = wq.get();
        const int st = it % NQK;
        ptx::mbar_wait(BAR(B_QEMPTY + st), (uint32_t)(((it / NQK) & 1) ^ 1));
        const uint32_t sb = smem0 + st * QKSTAGE;
        ptx::mbar_arrive_expect_tx(BAR(B_QFULL + st), QKSTAGE);
        const int kb = tp.k_batched ? uq.b : 0;
        ptx::tma_load_4d(sb, &tmq, BAR(B_QFULL + st), 0, uq.h, uq.tile * kBM, uq.b);
        ptx::tma_load_4d(sb + kQAtom, &tmk, BAR(B_QFULL + st), 0, uq.h, 0, kb);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ===================================== TMA producer: mask tiles and V tiles =====================================
    FwdWalk wv(u0, p.B, p.H, s_nb, s_img);
    int grp = -1;
    for (int it = 0; it < n_it; ++it, wv.next()) {
      const FwdUnit uv = wv.get();
      if (it == 0 || uv.j == 0) {                 // first unit of a group in this CTA's range: stage its mask tile
        ++grp;
        ptx::mbar_wait(BAR(B_MEMPTY), (uint32_t)((grp & 1) ^ 1));
        const int widx = (uv.mask_b >= 0 && uv.j <= uv.jl) ? s_widx[uv.mask_b] : -1;
        if (widx >= 0) {
          const int rows = min(kBM, p.N - uv.tile * kBM);
          const uint32_t bytes = (uint32_t)rows * p.T * 4u;
          const float* src = p.wmap + (int64_t)widx * p.wmap_bs + (int64_t)uv.tile * kBM * p.T;
          if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            if (lane == 0) {
              ptx::mbar_arrive_expect_tx(BAR(B_MFULL), bytes);
              ptx::bulk_load_1d(smem0 + OFF_MASK, src, bytes, BAR(B_MFULL));
            }
          } else {                                // ragged tail tile: plain loads by the whole warp
            float* dst = reinterpret_cast<float*>(smem_gen + OFF_MASK);
            for (int i = lane; i < rows * p.T; i += 32) dst[i] = __ldg(src + i);
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(BAR(B_MFULL));
          }
        } else if (lane == 0) {
          ptx::mbar_arrive(BAR(B_MFULL));         // keep the phases in lock-step for groups without a mask
        }
      }
      const int st = it % NV;
      ptx::mbar_wait(BAR(B_VEMPTY + st), (uint32_t)(((it / NV) & 1) ^ 1));
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(BAR(B_VFULL + st), VSTAGE);
        const int kb = tp.k_batched ? uv.b : 0;
        ptx::tma_load_4d(smem0 + OFF_V + st * VSTAGE, &tmv, BAR(B_VFULL + st), 0, uv.h, 0, kb);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================== UMMA issuer: S[g] = Q K^T =====================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = ptx::make_idesc_f16(128, kTP, false, false);
      for (int it = 0; it < n_it; ++it) {
        const int st = it % NQK, g = it % kG, local = it / kG;
        ptx::mbar_wait(BAR(B_QFULL + st), (uint32_t)((it / NQK) & 1));
        if (local >= 1)                            // the P.V that read P out of this group's buffer has finished
          ptx::mbar_wait(BAR(B_PVDONE + g), (uint32_t)((local - 1) & 1));
        ptx::tc_fence_after();
        const uint32_t sb = smem0 + st * QKSTAGE;
#pragma unroll
        for (int ks = 0; ks < 3; ++ks)             // K = 48: 40 head-dim columns + zero fill
          ptx::umma_ss(tmem_base + col_s(g), ptx::make_sw128_desc(sb + ks * 32, 16, 1024),
                       ptx::make_sw128_desc(sb + kQAtom + ks * 32, 16, 1024), idesc_qk, ks > 0);
        ptx::umma_commit(BAR(B_SREADY + g));
        ptx::umma_commit(BAR(B_QEMPTY + st));      // Q/K tiles are dead once S exists
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================================== UMMA issuer: O[g] = P V =====================================
    if (lane == 0) {
      constexpr uint32_t idesc_pv = ptx::make_idesc_f16(128, DPV, false, true);
      for (int j = 0; j < n_it; ++j) {
        const int st = j % NV, g = j % kG, local = j / kG;
        ptx::mbar_wait(BAR(B_PREADY + g), (uint32_t)(local & 1));
        ptx::mbar_wait(BAR(B_VFULL + st), (uint32_t)((j / NV) & 1));
        ptx::tc_fence_after();
        const uint32_t vb = smem0 + OFF_V + st * VSTAGE;
#pragma unroll
        for (int ks = 0; ks < kTP / 16; ++ks)      // A = P from tensor memory: 8 columns (16 fp16) per k-step
          ptx::umma_ts(tmem_base + col_o(g), tmem_base + col_s(g) + ks * 8,
                       ptx::make_sw128_desc(vb + ks * 16 * 128, kKAtom, 1024), idesc_pv, ks > 0);
        ptx::umma_commit(BAR(B_PVDONE + g));
        ptx::umma_commit(BAR(B_VEMPTY + st));
      }
    }
    __syncwarp();
  } else {
    // ===================================== softmax / epilogue groups =====================================
    const int g = (warp - 4) >> 2;
    const int row = ((warp & 3) << 5) | lane;                  // TMEM lane == tile row
    const uint32_t lane_addr = (uint32_t)((warp & 3) << 5) << 16;
    const uint32_t ts = tmem_base + lane_addr + col_s(g);      // this row of S, later of P
    const uint32_t to = tmem_base + lane_addr + col_o(g);
    const float sl2 = p.scale * 1.4426950408889634f;
    const int T = TT ? TT : p.T;
    const float* mask_row = reinterpret_cast<const float*>(smem_gen + OFF_MASK) + row * T;
    const uint32_t stg_off = OFF_STG + (uint32_t)(warp - 4) * STG_WARP;
    FwdWalk ws(u0, p.B, p.H, s_nb, s_img);
    int grp = -1;
    auto warp_arrive = [&](uint32_t bar) {       // one arrive per warp (barrier counts are per warp)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar);
    };

    for (int it = 0; it < n_it; ++it, ws.next()) {
      const FwdUnit ui = ws.get();
      if (ui.j == 0 || it == 0) ++grp;
      const bool release_here = (ui.j == mask_release_pos(ui, it, n_it));
      if ((it % kG) != g) {
        if (release_here) warp_arrive(BAR(B_MEMPTY));   // another group's unit: this warp's mask reads are behind it
        continue;
      }
      const int local = it / kG;
      const int widx = s_widx[ui.b];
      const float coef = s_coef[ui.b];
      ptx::mbar_wait(BAR(B_SREADY + g), (uint32_t)(local & 1));
      ptx::tc_fence_after();
      // ---- pass 1: t_j = S_j + coef * w_j (written back for biased units), row max
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      if (widx >= 0) ptx::mbar_wait(BAR(B_MFULL), (uint32_t)(grp & 1));
      // (TT = 77: every column below 64 is a real token and the checks fold away; TT = 0: any T <= 80)
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float s[32];
        ptx::tmem_ld32_sync(ts + c0, s);
        if (widx >= 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (TT == 77 || c0 + j < T) s[j] = fmaf(coef, mask_row[c0 + j], s[j]);
          ptx::tmem_st32(ts + c0, s);
        }
        if constexpr (TT != 77) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j >= T) s[j] = -INFINITY;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          m0 = fmaxf(m0, s[j]); m1 = fmaxf(m1, s[j + 1]); m2 = fmaxf(m2, s[j + 2]); m3 = fmaxf(m3, s[j + 3]);
        }
      }
      {
        float s[16];
        ptx::tmem_ld16_sync(ts + 64, s);
        if (widx >= 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (64 + j < T) s[j] = fmaf(coef, mask_row[64 + j], s[j]);
          ptx::tmem_st16(ts + 64, s);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (64 + j < T) m0 = fmaxf(m0, s[j]);
      }
      if (widx >= 0) ptx::tmem_st_wait();               // pass 2 re-reads what pass 1 wrote
      if (release_here) warp_arrive(BAR(B_MEMPTY));     // the mask values are consumed: the next tile may land
      const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      const float nm = -mx * sl2;
      // ---- pass 2: p_j = 2^(t_j*sl2 - mx*sl2), UNNORMALISED fp16, stored over the consumed part of the row
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float s[32];
        ptx::tmem_ld32_sync(ts + c0, s);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = (TT == 77 || c0 + j < T) ? ptx::ex2(fmaf(s[j], sl2, nm)) : 0.f;
          const float e1 = (TT == 77 || c0 + j + 1 < T) ? ptx::ex2(fmaf(s[j + 1], sl2, nm)) : 0.f;
          const __half2 h = __floats2half2_rn(e0, e1);
          pk[j / 2] = *reinterpret_cast<const uint32_t*>(&h);
        }
        ptx::tmem_st16_u32(ts + c0 / 2, pk);            // columns [c0/2, c0/2 + 16): already read
      }
      {
        float s[16];
        ptx::tmem_ld16_sync(ts + 64, s);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float e0 = (64 + j < T) ? ptx::ex2(fmaf(s[j], sl2, nm)) : 0.f;
          const float e1 = (64 + j + 1 < T) ? ptx::ex2(fmaf(s[j + 1], sl2, nm)) : 0.f;
          const __half2 h = __floats2half2_rn(e0, e1);
          pk[j / 2] = *reinterpret_cast<const uint32_t*>(&h);
        }
        ptx::tmem_st8_u32(ts + 32, pk);
      }
      ptx::tmem_st_wait();
      // row sums ride on the P.V UMMA: V[token r][D] = 1.0 for every real token r (see Cfg::ONES in xattn_tc.cuh)
      ptx::mbar_wait(BAR(B_VFULL + it % NV), (uint32_t)((it / NV) & 1));
      if (row < T) {
        unsigned char* vlast = smem_gen + OFF_V + (it % NV) * VSTAGE;
        constexpr int cc = kD;                   // spare column of the only V atom
        *reinterpret_cast<__half*>(vlast + row * 128 + ((((cc >> 3) ^ (row & 7))) << 4) + (cc & 7) * 2) = __float2half(1.0f);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      warp_arrive(BAR(B_PREADY + g));
      // ---- epilogue: O / rowsum -> fp16 -> staging tile -> TMA store
      ptx::mbar_wait(BAR(B_PVDONE + g), (uint32_t)(local & 1));
      ptx::tc_fence_after();
      float o[DPV];
      ptx::tmem_ld32_sync(to, o);
      ptx::tmem_ld16_sync(to + 32, o + 32);
      ptx::tc_fence_before();
      const float inv = 1.f / o[kD];
      if (lane == 0) ptx::bulk_wait_group_read0();      // this warp's previous store has been read out of the tile
      __syncwarp();
      unsigned char* dst = smem_gen + stg_off + lane * (kD * 2);
#pragma unroll
      for (int c = 0; c < kD / 8; ++c) {
        __align__(16) __half2 hk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) hk[q] = __floats2half2_rn(o[c * 8 + 2 * q] * inv, o[c * 8 + 2 * q + 1] * inv);
        reinterpret_cast<uint4*>(dst)[c] = *reinterpret_cast<const uint4*>(hk);
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      const int n0 = ui.tile * kBM + ((warp & 3) << 5);
      if (lane == 0 && n0 < p.N) {
        ptx::tma_store_4d(&tmo, smem0 + stg_off, 0, ui.h, n0, ui.b);
        ptx::bulk_commit_group();
      }
    }
    if (lane == 0) ptx::bulk_wait_group0();             // the staging tile must outlive the last store
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace g4

template <int G>
cudaError_t launch_fwd_g4_impl(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                               const CUtensorMap& to, const TcParams& tp, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(g4::xattn_fwd_g4_kernel<77, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, g4::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(g4::xattn_fwd_g4_kernel<0, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, g4::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int grid = tp.units < num_sms() ? tp.units : num_sms();
  if (tp.x.T == 77)
    g4::xattn_fwd_g4_kernel<77, G><<<grid, 128 + G * 128, g4::SMEM, s>>>(tq, tk, tv, to, tp);
  else
    g4::xattn_fwd_g4_kernel<0, G><<<grid, 128 + G * 128, g4::SMEM, s>>>(tq, tk, tv, to, tp);
  return cudaGetLastError();
}
// groups = 4: the design point; groups = 2: the same code with the thread count of the shipped kernel (bisection)
inline cudaError_t launch_fwd_g4(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                                 const CUtensorMap& to, const TcParams& tp, cudaStream_t s, int groups) {
  return groups == 2 ? launch_fwd_g4_impl<2>(tq, tk, tv, to, tp, s) : launch_fwd_g4_impl<4>(tq, tk, tv, to, tp, s);
}

}  // namespace tc
}  // namespace pww
