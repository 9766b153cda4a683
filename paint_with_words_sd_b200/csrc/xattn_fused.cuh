// Paint-with-Words cross-attention, ONE launch: statistic + bias + softmax + P.V on tcgen05 tensor cores (sm_100a).
//
// Fused region of the reference's inj_forward (paint_with_words.py:87-118) including the weight function's grid-wide
// statistic (paint_with_words.py:402-405: w * log(1 + sigma) * qk.max(), README variants with qk.std()):
//     S    = Q_h K_h^T                          UMMA M=128 N=80 K=D(+pad), operands by TMA, accumulator in TMEM
//     M_b  = max / unbiased std of fp16(S) over ALL heads, rows and tokens of image b   (grid-wide, per image)
//     S'   = S + (g * M_b) * W_b                the bias is two more k-steps of the SAME UMMA chain (see below)
//     P    = softmax(scale * S')                two threads per query row (40 columns each)
//     O    = P V_h                              TS-form UMMA, P (fp16) written over S in TMEM
//
// Packed weight map.  The reference's dense [N, 77] fp32 map has at most a handful of distinct non-zero
// columns (one per painted region, paint_with_words.py:255-272), so it is stored as a column dictionary:
//     W[n, t] = Mu[n, cidx[t]]      Mu [N, R] fp32 (R <= 10 distinct columns), cidx [77] (-1 = zero column)
// and Mu is split into fp16 hi/lo halves:  mpack[n] = [ hi(Mu[n,0..9]) | lo(Mu[n,0..9]) | hi(Mu[n,0..9]) | 0 0 ]  (32 fp16,
// 64 bytes per row instead of 308).  With x = g * M_b split the same way (xh, xl) the B operand row of token t is
// [ xh at cidx[t] | xh at 10 + cidx[t] | xl at 20 + cidx[t] ], so the two extra k-steps add
// hi*xh + lo*xh + hi*xl = W * x to 2^-22 relative -- the 77 loads + 77 FMAs per biased row of the dense kernel are gone.
//
// One launch.  Every CTA owns a contiguous range of (image, row tile, head) units in an order that alternates the units of
// a biased image and of an unbiased one (classifier-free guidance supplies both), and runs three job lists over it:
//     stat jobs   S of every BIASED unit -> per-thread max / sums -> per-CTA partial per image -> global, arrive
//     main jobs   the UNBIASED units: softmax + P.V need nothing from other CTAs and overlap the grid barrier
//     main jobs   the BIASED units: S is recomputed (Q/K/M tiles are prefetched during the wait; 3 k-steps are cheaper than
//                 parking S in TMEM), the S-issuer warp first waits until every CTA that owns units of the image has
//                 published, reduces the partials in a fixed order and builds the B-operand tile above.
// The launch is cooperative (all CTAs co-resident, grid <= number of SMs), partials are reduced in a fixed order by every
// consumer: the statistic is deterministic and identical on every CTA.
//
// Warp roles (640 threads): 0 TMA producer Q/K/M | 1 UMMA issuer S (+ TMEM owner, grid barrier, B-operand tile) |
// 2 TMA producer V | 3 UMMA issuer P.V (+ ones column) | 4-11 softmax group 0 | 12-19 softmax group 1.
// A softmax group is 8 warps = 128 rows x 2 column halves; the two threads of a row exchange their partial row maxima
// through shared memory (named barrier of 64 threads).  16 softmax warps instead of 8 is what the round-1 profile asked for
// (issue slots 39 % busy, stalls dominated by fixed-latency dependencies with 2 warps per scheduler).
//
// Barrier rule: every mbarrier is waited on, phase after phase, by the same threads; no waiter ever skips
// a phase, and every arrival of a multi-arrival barrier is causally behind the completion of the previous phase.
#pragma once
#include "ptx_sm100.cuh"
#include "pww_common.cuh"
#include "xattn_tc.cuh"   // make_tmap, make_tmap_out, encode_fn, num_sms, tc_error_buf

namespace pww {
namespace fx {

constexpr int kBM = 128;          // query rows per tile
constexpr int kTP = 80;           // padded key count
constexpr int kThreads = 640;     // 20 warps
constexpr int kMaxBatch = 32;     // images per launch (the C ABI splits larger batches)
constexpr int kMaxLocal = 4;      // biased images one CTA's unit range may touch (checked on the host)
constexpr int kMW = 32;           // packed-map columns per row = two 16-wide k-steps
constexpr int kRC = 10;           // dictionary capacity (distinct non-zero columns)
constexpr uint32_t kQAtom = 128 * 128, kKAtom = kTP * 128, kMAtom = 128 * 64, kCoefTile = kTP * 64;

template <int D>
struct Cfg {
  static constexpr int NA = (D + 63) / 64;          // 64-column atoms along the head dim
  static constexpr int DP = (D + 15) / 16 * 16;
  static constexpr int KSTEPS = DP / 16;
  static constexpr bool ONES = (D == 40 || D == 80);            // row sums ride on the P.V UMMA (spare V column = 1.0)
  static constexpr int DPV = ONES ? (D + 16) / 16 * 16 : DP;    // UMMA N of P.V: 48, 64, 96, 160
  // TMEM: NS score slots of 80 fp32 columns (P, 40 columns of packed fp16, overwrites the S it came from) and NO output
  // accumulators.  Job i uses slot i % NS and softmax group i % 2, so a slot always belongs to the same group.
  static constexpr int NS = (D <= 80) ? 4 : 2;
  static constexpr int NO = (D == 40) ? 4 : 2;
  static constexpr uint32_t O_STRIDE = DPV;
  __host__ __device__ static constexpr uint32_t col_s(int slot) { return (uint32_t)slot * 80u; }
  __host__ __device__ static constexpr uint32_t col_o(int os) { return (uint32_t)NS * 80u + (uint32_t)os * O_STRIDE; }
  static_assert(NS * 80 + NO * DPV <= 512, "TMEM budget");
  static constexpr int NQK = (D <= 64) ? 3 : (D == 80 ? 2 : 1);
  static constexpr int NV = (D <= 64) ? 3 : (D == 80 ? 2 : 1);
  static constexpr uint32_t QKBYTES = NA * (kQAtom + kKAtom);
  static constexpr uint32_t QKSTAGE = QKBYTES + kMAtom;        // Q atoms | K atoms | packed-map atom
  static constexpr uint32_t VSTAGE = NA * kKAtom;
  // output columns of the two threads of a row; every piece is a multiple of 8 columns (16-byte TMA boxes)
  static constexpr int C0 = (D == 40) ? 24 : D / 2;
  static constexpr int C1 = D - C0;
  static constexpr int EPI_W = (D == 160) ? 40 : C0;            // columns per epilogue pass (staging row width)
  static constexpr int EPI_NPASS = (D == 160) ? 2 : 1;
  static constexpr uint32_t STG_WARP = 32 * EPI_W * 2;
  static constexpr uint32_t OFF_V = NQK * QKSTAGE;
  static constexpr uint32_t OFF_COEF = OFF_V + NV * VSTAGE;     // 2 B-operand tiles [80 x 32 fp16], 64-byte-swizzled rows
  static constexpr uint32_t OFF_STG = OFF_COEF + 2 * kCoefTile;
  static constexpr uint32_t OFF_XCHG = OFF_STG + 16 * STG_WARP; // [group][buf][half][128] fp32 row maxima, then row sums
  static constexpr uint32_t OFF_BAR = OFF_XCHG + 2 * 2 * 2 * 128 * 4 * 2;
  static constexpr uint32_t SMEM = OFF_BAR + 512 + 1024;        // + alignment slack
  static_assert(SMEM <= 232448 - 6144, "shared memory budget (dynamic + static tables)");
};

struct FxParams {
  XattnParams x;            // q/k/v/out, strides, wmap_index, g_sigma, scale, stat, stats_out, counters, partials
  const int8_t* cidx;       // [Bw, 80] dictionary column per token, -1 = none
  int tiles, units, k_batched;
  int grid;                 // CTAs (== gridDim.x): partial slots per image
  long long* timeline;      // debug only: clock64 stamps [tag][job] of CTA `tl_cta` (see FX_TL)
  int tl_cta;
  int hg;                   // head groups per row tile (grouped-head kernel, xattn_fused2.cuh)
  unsigned* jobs_dump;      // debug only: [grid][2 + 2 * 512] = njobs, nstat, job table of every CTA (grouped-head kernel)
};
constexpr int kFxTlTags = 24, kFxTlIts = 64;
#define FX_TL(tag, it)                                                                                     \
  do {                                                                                                     \
    if constexpr (kTimeline) {                                                                             \
      if (fp.timeline != nullptr && (int)blockIdx.x == fp.tl_cta && (it) >= 0 && (it) < kFxTlIts)         \
        fp.timeline[(tag) * kFxTlIts + (it)] = clock64();                                                  \
    }                                                                                                      \
  } while (0)

// ------------------------------------------------------------------------------------------------------------------
// unit order (shared by the kernel and the host replay)
// ------------------------------------------------------------------------------------------------------------------
// `img` lists the biased images first (nb of them), then the unbiased ones.  Groups: np = min(nb, nu) PAIR groups (biased
// image img[g] + unbiased image img[nb + g], tiles * 2H units: tile-major, then head, biased unit before unbiased), then
// the SOLO groups of the images without a partner (tiles * H units).  A CTA takes a contiguous range of this order, so it
// gets the same number of biased and unbiased units (+-1) whatever the image order of the batch, and stays inside one or
// two images.
struct FxUnit {
  int b, h, tile, biased, gi;
};
struct FxWalk {
  int B, H, tiles, nb, np;
  const int* img;
  int gi, tile, j;
  __host__ __device__ __forceinline__ FxWalk() {}
  __host__ __device__ __forceinline__ FxWalk(int u, int B_, int H_, int tiles_, int nb_, const int* img_)
      : B(B_), H(H_), tiles(tiles_), nb(nb_), img(img_) {
    const int nu = B - nb;
    np = nb < nu ? nb : nu;
    const int per_pair = tiles * 2 * H;
    if (u < np * per_pair) {
      gi = u / per_pair;
      const int r = u - gi * per_pair;
      tile = r / (2 * H);
      j = r - tile * 2 * H;
    } else {
      u -= np * per_pair;
      const int per_solo = tiles * H;
      const int s = u / per_solo;
      gi = np + s;
      const int r = u - s * per_solo;
      tile = r / H;
      j = r - tile * H;
    }
  }
  __host__ __device__ __forceinline__ void next() {
    const int gsize = gi < np ? 2 * H : H;
    if (++j == gsize) {
      j = 0;
      if (++tile == tiles) { tile = 0; ++gi; }
    }
  }
  __host__ __device__ __forceinline__ FxUnit get() const {
    FxUnit r;
    r.tile = tile;
    r.gi = gi;
    if (gi < np) {
      r.h = j >> 1;
      r.biased = (j & 1) ^ 1;
      r.b = r.biased ? img[gi] : img[nb + gi];
    } else {
      r.h = j;
      r.biased = (2 * nb > B) ? 1 : 0;
      r.b = r.biased ? img[gi] : img[nb + gi];
    }
    return r;
  }
};
// 32-bit arithmetic on purpose: a 64-bit division is ~100 SASS instructions and this is inlined at every membership test
// (it was 28 % of the grouped-head kernel's code); the host checks units * grid < 2^32 (fused_units_ok).
__host__ __device__ __forceinline__ void fx_range(int cta, int grid, int units, int& u0, int& u1) {
  u0 = (int)((unsigned)cta * (unsigned)units / (unsigned)grid);
  u1 = (int)((unsigned)(cta + 1) * (unsigned)units / (unsigned)grid);
}
inline bool fused_units_ok(long long units, int grid) { return units > 0 && units * (long long)(grid + 1) < (1ll << 32); }
// Does CTA `cta` own at least one unit of the BIASED image at list position `pos` (== its group index)?
__host__ __device__ __forceinline__ bool fx_cta_has_image(int cta, int grid, int units, int pos, int H, int tiles, int np) {
  int lo, hi;
  fx_range(cta, grid, units, lo, hi);
  const int per_pair = tiles * 2 * H, per_solo = tiles * H;
  const bool pair = pos < np;
  const int base = pair ? pos * per_pair : np * per_pair + (pos - np) * per_solo;
  const int len = pair ? per_pair : per_solo;
  const int a = lo > base ? lo : base, b = hi < base + len ? hi : base + len;
  if (a >= b) return false;
  if (!pair) return true;
  return (b - a >= 2) || (((a - base) & 1) == 0);        // biased units sit at even offsets of a pair group
}

// Job lists of one CTA: stat jobs (biased units), main jobs of the unbiased units, main jobs of the biased units.
struct FxJob {
  int b, h, tile, biased, gi;
  int kind;     // 0 = stat, 1 = main
  int i;        // global job index (slot = i % NS, softmax group = i % 2, Q/K stage = i % NQK)
  int m;        // main-job index (V stage = m % NV), -1 for stat jobs
  int li;       // local index of the biased image inside this CTA's range (biased jobs)
};
struct FxJobs {
  FxWalk w;
  int u0, n_it, B, H, tiles, nb;
  const int* img;
  int it, phase, i, m, li, lastb;
  __host__ __device__ __forceinline__ FxJobs(int u0_, int n_it_, int B_, int H_, int tiles_, int nb_, const int* img_)
      : u0(u0_), n_it(n_it_), B(B_), H(H_), tiles(tiles_), nb(nb_), img(img_) {
    phase = nb > 0 ? 0 : 1;
    i = 0; m = 0;
    rewind();
  }
  __host__ __device__ __forceinline__ void rewind() {
    w = FxWalk(u0, B, H, tiles, nb, img);
    it = 0; li = -1; lastb = -1;
  }
  __host__ __device__ __forceinline__ bool next(FxJob& jb) {
    for (;;) {
      while (it < n_it) {
        const FxUnit u = w.get();
        w.next();
        ++it;
        if (u.biased && u.b != lastb) { ++li; lastb = u.b; }
        const bool want = (phase == 1) ? !u.biased : (u.biased != 0);
        if (!want) continue;
        jb.b = u.b; jb.h = u.h; jb.tile = u.tile; jb.biased = u.biased; jb.gi = u.gi;
        jb.kind = phase == 0 ? 0 : 1;
        jb.i = i++;
        jb.m = phase == 0 ? -1 : m++;
        jb.li = u.biased ? li : -1;
        return true;
      }
      if (phase == 2) return false;
      ++phase;
      rewind();
    }
  }
};

// ------------------------------------------------------------------------------------------------------------------
// small PTX helpers local to this kernel
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// Shared-memory matrix descriptor, K-major SWIZZLE_64B: rows of 64 bytes, 8-row groups of 512 bytes (layout type 4).
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(16u >> 4) << 16;
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// 40 consecutive TMEM columns of this thread's lane (x32 + x8), one wait.
__device__ __forceinline__ void tmem_ld40_sync(uint32_t taddr, float* v) {
  uint32_t r[40];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%40];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%32, %33, %34, %35, %36, %37, %38, %39}, [%41];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39])
      : "r"(taddr), "r"(taddr + 32)
      : "memory");
#pragma unroll
  for (int i = 0; i < 40; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld1_sync(uint32_t taddr, float* v) {
  uint32_t r;
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r);
}
__device__ __forceinline__ void tmem_st4_u32(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
template <int D, int TT>
__global__ void __launch_bounds__(kThreads, 1)
xattn_fused_kernel(const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
                   const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmm,
                   const __grid_constant__ CUtensorMap tmo0, const __grid_constant__ CUtensorMap tmo1,
                   const FxParams fp) {
  using C = Cfg<D>;
  constexpr bool kTimeline = true;                // clock64 stamps compiled in (debug timeline of this kernel)
  const XattnParams& p = fp.x;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem0 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem_gen = smem_raw + (smem0 - ptx::smem_u32(smem_raw));
  const uint32_t bar0 = smem0 + C::OFF_BAR;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_QFULL = 0, B_QEMPTY = 3, B_VFULL = 6, B_VEMPTY = 9, B_SREADY = 12, B_SFREE = 16, B_PREADY = 20,
                B_PVDONE = 24, B_OFREE = 28, B_COEF = 32, B_TMEMPTR = 34;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = TT ? TT : p.T;
  int u0, u1;
  fx_range(blockIdx.x, gridDim.x, fp.units, u0, u1);
  const int n_it = u1 - u0;

  __shared__ int s_widx[kMaxBatch];
  __shared__ int s_img[kMaxBatch];                // biased images first, then unbiased
  __shared__ int s_nb;
  __shared__ float s_coef[kMaxLocal];             // g(sigma) * statistic of the CTA's local biased images
  __shared__ StatPartial s_part[16][kMaxLocal];   // [softmax warp][local biased image]

  if (warp == 2) {                                // stable partition of the images by "has a weight map"
    const int b = lane;
    const int wi = (b < p.B && p.wmap != nullptr) ? (p.wmap_index ? p.wmap_index[b] : b) : -1;   // wmap == NULL: no maps at all
    const bool valid = b < p.B, bi = valid && wi >= 0;
    const unsigned mb = __ballot_sync(0xffffffffu, bi), mu = __ballot_sync(0xffffffffu, valid && !bi);
    const unsigned lt = (1u << lane) - 1u;
    const int nbt = __popc(mb);
    if (bi) s_img[__popc(mb & lt)] = b;
    else if (valid) s_img[nbt + __popc(mu & lt)] = b;
    if (valid) s_widx[b] = wi;
    if (lane == 0) s_nb = nbt;
  }
  for (int i = threadIdx.x; i < 16 * kMaxLocal; i += kThreads) {
    StatPartial sp;
    sp.vmax = -INFINITY; sp.sum = 0.0; sp.sumsq = 0.0; sp.pad = 0.0;
    s_part[i / kMaxLocal][i % kMaxLocal] = sp;
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmq);
    ptx::prefetch_tmap(&tmk);
    ptx::prefetch_tmap(&tmv);
    ptx::prefetch_tmap(&tmm);
    ptx::prefetch_tmap(&tmo0);
    ptx::prefetch_tmap(&tmo1);
    for (int s = 0; s < 3; ++s) {
      ptx::mbar_init(BAR(B_QFULL + s), 1);
      ptx::mbar_init(BAR(B_QEMPTY + s), 1);
      ptx::mbar_init(BAR(B_VFULL + s), 1);
      ptx::mbar_init(BAR(B_VEMPTY + s), 1);
    }
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(BAR(B_SREADY + s), 1);
      ptx::mbar_init(BAR(B_SFREE + s), 8);       // one elected arrive per warp of the slot's softmax group
      ptx::mbar_init(BAR(B_PREADY + s), 8);
      ptx::mbar_init(BAR(B_PVDONE + s), 1);
      ptx::mbar_init(BAR(B_OFREE + s), 8);
    }
    ptx::mbar_init(BAR(B_COEF + 0), 1);
    ptx::mbar_init(BAR(B_COEF + 1), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(BAR(B_TMEMPTR));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + C::OFF_BAR + 8 * B_TMEMPTR);
  const int nb = s_nb;
  if (threadIdx.x == 0) FX_TL(13, 0);
  const int nu_img = p.B - nb;
  const int np = nb < nu_img ? nb : nu_img;

  if (blockIdx.x == 0 && p.stats_out != nullptr)            // images without a weight map report statistic 0
    for (int b = threadIdx.x; b < p.B; b += kThreads)
      if (s_widx[b] < 0) p.stats_out[b] = 0.f;

  if (warp == 0) {
    // ============================== TMA producer: Q, K and packed-map tiles, one stage per job ==============================
    if (lane == 0) {
      FxJobs jobs(u0, n_it, p.B, p.H, fp.tiles, nb, s_img);
      FxJob jb;
      while (jobs.next(jb)) {
        const int st = jb.i % C::NQK;
        ptx::mbar_wait(BAR(B_QEMPTY + st), (uint32_t)(((jb.i / C::NQK) & 1) ^ 1));
        const uint32_t sb = smem0 + st * C::QKSTAGE;
        const bool with_map = jb.kind == 1 && jb.biased;
        ptx::mbar_arrive_expect_tx(BAR(B_QFULL + st), C::QKBYTES + (with_map ? kMAtom : 0u));
        const int kb = fp.k_batched ? jb.b : 0;
#pragma unroll
        for (int a = 0; a < C::NA; ++a) {
          ptx::tma_load_4d(sb + a * kQAtom, &tmq, BAR(B_QFULL + st), a * 64, jb.h, jb.tile * kBM, jb.b);
          ptx::tma_load_4d(sb + C::NA * kQAtom + a * kKAtom, &tmk, BAR(B_QFULL + st), a * 64, jb.h, 0, kb);
        }
        if (with_map) tma_load_3d(sb + C::QKBYTES, &tmm, BAR(B_QFULL + st), 0, jb.tile * kBM, s_widx[jb.b]);
        FX_TL(0, jb.i);
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // ============================== TMA producer: V tiles of the main jobs ==============================
    if (lane == 0) {
      FxJobs jobs(u0, n_it, p.B, p.H, fp.tiles, nb, s_img);
      FxJob jb;
      while (jobs.next(jb)) {
        if (jb.kind == 0) continue;
        const int st = jb.m % C::NV;
        ptx::mbar_wait(BAR(B_VEMPTY + st), (uint32_t)(((jb.m / C::NV) & 1) ^ 1));
        const uint32_t sb = smem0 + C::OFF_V + st * C::VSTAGE;
        ptx::mbar_arrive_expect_tx(BAR(B_VFULL + st), C::VSTAGE);
        const int kb = fp.k_batched ? jb.b : 0;
#pragma unroll
        for (int a = 0; a < C::NA; ++a) ptx::tma_load_4d(sb + a * kKAtom, &tmv, BAR(B_VFULL + st), a * 64, jb.h, 0, kb);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================== UMMA issuer: S of every job (+ grid barrier and the bias operand) ==============================
    constexpr uint32_t idesc_qk = ptx::make_idesc_f16(128, kTP, false, false);
    FxJobs jobs(u0, n_it, p.B, p.H, fp.tiles, nb, s_img);
    FxJob jb;
    int ns = -1;                                   // number of stat jobs (known once the first main job shows up)
    bool stats_ready = false;
    int cur_li = -1;
    while (jobs.next(jb)) {
      if (jb.kind == 1 && ns < 0) ns = jb.i;
      if (jb.kind == 1 && jb.biased) {
        if (!stats_ready) {
          // ---- grid barrier: every CTA owning units of my biased images has published its partial ----
          int lb[kMaxLocal], lp[kMaxLocal], nl = 0;
          {
            FxWalk w(u0, p.B, p.H, fp.tiles, nb, s_img);
            int last = -1;
            for (int it = 0; it < n_it; ++it, w.next()) {
              const FxUnit u = w.get();
              if (u.biased && u.b != last) {
                if (nl < kMaxLocal) { lb[nl] = u.b; lp[nl] = u.gi; }
                ++nl;
                last = u.b;
              }
            }
          }
          const int G = (int)gridDim.x;
          if (lane == 0) FX_TL(10, 0);
          for (int l = 0; l < nl && l < kMaxLocal; ++l) {
            const int b = lb[l];
            int expect = 0, first_c = 1 << 30;
            for (int c = lane; c < G; c += 32)
              if (fx_cta_has_image(c, G, fp.units, lp[l], p.H, fp.tiles, np)) { ++expect; if (c < first_c) first_c = c; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              expect += __shfl_xor_sync(0xffffffffu, expect, o);
              first_c = min(first_c, __shfl_xor_sync(0xffffffffu, first_c, o));
            }
            if (lane == 0) {
              const long long t0 = clock64();
              while (ld_acquire_gpu(p.counters + b) < (unsigned)expect) {
                __nanosleep(64);
                if (clock64() - t0 > 20000000000LL) {
                  printf("pww: grid barrier timeout block %d image %d have %u want %d\n", blockIdx.x, b,
                         ld_acquire_gpu(p.counters + b), expect);
                  __trap();
                }
              }
            }
            __syncwarp();
            (void)ld_acquire_gpu(p.counters + b);                  // every lane orders its partial loads behind the counter
            double m = -INFINITY, a = 0.0, q = 0.0;
            for (int c = lane; c < G; c += 32)
              if (fx_cta_has_image(c, G, fp.units, lp[l], p.H, fp.tiles, np)) {
                const StatPartial* pp = p.partials + (int64_t)b * G + c;
                m = fmax(m, __ldcg(&pp->vmax));
                a += __ldcg(&pp->sum);
                q += __ldcg(&pp->sumsq);
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
              const float st16 = round_to_f16((float)r);          // qk.max() / qk.std() return fp16 in the reference
              s_coef[l] = __ldg(p.g_sigma) * st16;
              if (first_c == (int)blockIdx.x && p.stats_out != nullptr) p.stats_out[b] = st16;
            }
          }
          __syncwarp();
          if (lane == 0) FX_TL(11, 0);
          stats_ready = true;
        }
        if (jb.li != cur_li) {
          // ---- B operand of the bias k-steps for this image: [80 tokens x 32] fp16, 64-byte-swizzled rows ----
          const int buf = jb.li & 1;
          if (cur_li >= 0 && lane == 0) ptx::umma_commit(BAR(B_COEF + (cur_li & 1)));   // UMMAs reading the old tile
          if (jb.li >= 2) ptx::mbar_wait(BAR(B_COEF + buf), (uint32_t)(((jb.li >> 1) - 1) & 1));
          const float x = s_coef[jb.li < kMaxLocal ? jb.li : 0];
          const __half xh = __float2half_rn(x);
          const __half xl = __float2half_rn(x - __half2float(xh));
          unsigned char* tile = smem_gen + C::OFF_COEF + buf * kCoefTile;
          const int8_t* ci = fp.cidx + (int64_t)s_widx[jb.b] * kTP;
          for (int t = lane; t < kTP; t += 32) {
            uint4* rowp = reinterpret_cast<uint4*>(tile + t * 64);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) rowp[ch] = make_uint4(0, 0, 0, 0);
            const int r = (t < T) ? (int)ci[t] : -1;
            if (r >= 0 && r < kRC) {
              auto put = [&](int k, __half v) {
                *reinterpret_cast<__half*>(tile + t * 64 + ((((k >> 3) ^ ((t >> 1) & 3))) << 4) + (k & 7) * 2) = v;
              };
              put(r, xh);
              put(kRC + r, xh);
              put(2 * kRC + r, xl);
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          cur_li = jb.li;
        }
      }
      if (lane == 0) {
        const int st = jb.i % C::NQK, slot = jb.i % C::NS;
        ptx::mbar_wait(BAR(B_QFULL + st), (uint32_t)((jb.i / C::NQK) & 1));
        FX_TL(1, jb.i);
        if (jb.i >= C::NS) {                           // the previous job on this slot is done with it
          const int prev = jb.i - C::NS;
          if (jb.kind == 0 || prev < ns) ptx::mbar_wait(BAR(B_SFREE + slot), (uint32_t)((prev / C::NS) & 1));
          else ptx::mbar_wait(BAR(B_PVDONE + slot), (uint32_t)(((prev - ns) / C::NS) & 1));
        }
        ptx::tc_fence_after();
        FX_TL(2, jb.i);
        const uint32_t sb = smem0 + st * C::QKSTAGE;
#pragma unroll
        for (int ks = 0; ks < C::KSTEPS; ++ks) {
          const uint32_t qa = sb + (ks / 4) * kQAtom + (ks % 4) * 32;
          const uint32_t ka = sb + C::NA * kQAtom + (ks / 4) * kKAtom + (ks % 4) * 32;
          ptx::umma_ss(tmem_base + C::col_s(slot), ptx::make_sw128_desc(qa, 16, 1024), ptx::make_sw128_desc(ka, 16, 1024),
                       idesc_qk, ks > 0);
        }
        if (jb.kind == 1 && jb.biased) {
          const uint32_t ma = sb + C::QKBYTES, ca = smem0 + C::OFF_COEF + (jb.li & 1) * kCoefTile;
#pragma unroll
          for (int ks = 0; ks < kMW / 16; ++ks)
            ptx::umma_ss(tmem_base + C::col_s(slot), make_sw64_desc(ma + ks * 32), make_sw64_desc(ca + ks * 32), idesc_qk, true);
        }
        ptx::umma_commit(BAR(B_SREADY + slot));
        ptx::umma_commit(BAR(B_QEMPTY + st));      // Q/K/map tiles are dead once S exists
        FX_TL(3, jb.i);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ============================== UMMA issuer: O = P V of every main job ==============================
    // The whole warp waits for the V tile (every VFULL phase observed by the same threads, in order) and sets the spare
    // column of the last V atom to 1.0 for every real token (D = 40 / 80): accumulator column D is the row sum.
    constexpr uint32_t idesc_pv = ptx::make_idesc_f16(128, C::DPV, false, true);
    FxJobs jobs(u0, n_it, p.B, p.H, fp.tiles, nb, s_img);
    FxJob jb;
    while (jobs.next(jb)) {
      if (jb.kind == 0) continue;
      const int st = jb.m % C::NV, slot = jb.i % C::NS, os = jb.i % C::NO;
      ptx::mbar_wait(BAR(B_VFULL + st), (uint32_t)((jb.m / C::NV) & 1));
      if (lane == 0) FX_TL(7, jb.i);
      if constexpr (C::ONES) {
        unsigned char* vlast = smem_gen + C::OFF_V + st * C::VSTAGE + (C::NA - 1) * kKAtom;
        constexpr int cc = D % 64;                 // spare column inside the last atom
        for (int r = lane; r < T; r += 32)
          *reinterpret_cast<__half*>(vlast + r * 128 + ((((cc >> 3) ^ (r & 7))) << 4) + (cc & 7) * 2) = __float2half(1.0f);
        ptx::fence_proxy_async_smem();
        __syncwarp();
      }
      if (lane == 0) {
        ptx::mbar_wait(BAR(B_PREADY + slot), (uint32_t)((jb.m / C::NS) & 1));
        FX_TL(8, jb.i);
        if (jb.m >= C::NO) ptx::mbar_wait(BAR(B_OFREE + os), (uint32_t)(((jb.m / C::NO) - 1) & 1));
        ptx::tc_fence_after();
        const uint32_t vb = smem0 + C::OFF_V + st * C::VSTAGE;
#pragma unroll
        for (int ks = 0; ks < kTP / 16; ++ks)          // A = P from tensor memory: 8 columns (16 fp16) per k-step
          ptx::umma_ts(tmem_base + C::col_o(os), tmem_base + C::col_s(slot) + ks * 8,
                       ptx::make_sw128_desc(vb + ks * 16 * 128, kKAtom, 1024), idesc_pv, ks > 0);
        ptx::umma_commit(BAR(B_PVDONE + slot));
        ptx::umma_commit(BAR(B_VEMPTY + st));
        FX_TL(9, jb.i);
      }
      __syncwarp();
    }
  } else {
    // ============================== softmax groups: 8 warps = 128 rows x 2 column halves ==============================
    const int sw = warp - 4;                       // 0..15
    const int g = sw >> 3;                         // softmax group
    const int c = (sw >> 2) & 1;                   // column half: S columns [40c, 40c + 40)
    const int qd = sw & 3;                         // TMEM lane quarter (== warp % 4)
    const int row = (qd << 5) | lane;
    const uint32_t lane_addr = (uint32_t)(qd << 5) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int pair_bar = 1 + g * 4 + qd;           // named barrier shared by the two warps of a row quarter
    float* xchg = reinterpret_cast<float*>(smem_gen + C::OFF_XCHG);          // [g][buf][half][128] maxima
    float* xsum = xchg + 2 * 2 * 2 * 128;                                     // [g][buf][half][128] row sums (!ONES)
    const uint32_t stg_off = C::OFF_STG + (uint32_t)sw * C::STG_WARP;
    const int oc0 = c ? C::C0 : 0;                 // first output column of this thread
    const int ocn = c ? C::C1 : C::C0;             // number of output columns of this thread

    auto warp_arrive = [&](uint32_t bar) {         // one arrive per warp (barrier counts are per warp)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar);
    };

    // ---- stat jobs: per-thread partial of the statistic over this thread's 40 columns ----
    float vmax = -INFINITY;
    double dsum = 0.0, dsq = 0.0;
    int cur_li = -1;
    auto flush = [&]() {
      if (cur_li < 0) return;
      double m = vmax, a = dsum, q = dsq;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane == 0 && cur_li < kMaxLocal) {
        StatPartial sp;
        sp.vmax = m; sp.sum = a; sp.sumsq = q; sp.pad = 1.0;
        s_part[sw][cur_li] = sp;
      }
      vmax = -INFINITY; dsum = 0.0; dsq = 0.0;
    };

    FxJobs jobs(u0, n_it, p.B, p.H, fp.tiles, nb, s_img);
    FxJob jb;
    bool more = jobs.next(jb);
    int n_stat = 0;
    while (more && jb.kind == 0) {
      ++n_stat;
      if ((jb.i & 1) == g) {
        const int slot = jb.i % C::NS;
        if (jb.li != cur_li) { flush(); cur_li = jb.li; }
        ptx::mbar_wait(BAR(B_SREADY + slot), (uint32_t)((jb.i / C::NS) & 1));
        ptx::tc_fence_after();
        if ((sw & 7) == 0 && lane == 0) FX_TL(4, jb.i);
        float s[40];
        tmem_ld40_sync(tmem_base + lane_addr + C::col_s(slot) + c * 40, s);
        ptx::tc_fence_before();
        warp_arrive(BAR(B_SFREE + slot));
        if ((sw & 7) == 0 && lane == 0) FX_TL(5, jb.i);
        if (jb.tile * kBM + row < p.N) {
          if (p.stat == PWW_STAT_MAX) {
            // max(fp16(s)) == fp16(max(s)): rounding is monotonic, so round once at the very end
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 40; j += 2) {
              if (c * 40 + j < T) m0 = fmaxf(m0, s[j]);
              if (c * 40 + j + 1 < T) m1 = fmaxf(m1, s[j + 1]);
            }
            vmax = fmaxf(vmax, fmaxf(m0, m1));
          } else {
            float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int j = 0; j < 40; j += 2) {
              // padded columns hold exact zeros (K rows >= T are zero-filled), so they add nothing
              const __half2 h = __floats2half2_rn(s[j], s[j + 1]);
              const float2 f = __half22float2(h);
              a0 += f.x; a1 += f.y;
              q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
            }
            dsum += (double)(a0 + a1);
            dsq += (double)(q0 + q1);
          }
        }
      }
      more = jobs.next(jb);
    }
    if (n_stat > 0) {
      flush();
      ptx::named_bar_sync(9, 512);               // all 16 softmax warps have written their partials
      if (sw == 0) {
        // publish: one lane per local image reduces the 16 warps in a fixed order, writes the CTA's slot, arrives
        int lbv = -1, nl = 0;
        {
          FxWalk w(u0, p.B, p.H, fp.tiles, nb, s_img);
          int last = -1;
          for (int it = 0; it < n_it; ++it, w.next()) {
            const FxUnit u = w.get();
            if (u.biased && u.b != last) {
              if (nl == lane) lbv = u.b;
              ++nl;
              last = u.b;
            }
          }
        }
        if (lane < nl && lane < kMaxLocal) {
          StatPartial sp = s_part[0][lane];
          for (int w2 = 1; w2 < 16; ++w2) {
            sp.vmax = fmax(sp.vmax, s_part[w2][lane].vmax);
            sp.sum += s_part[w2][lane].sum;
            sp.sumsq += s_part[w2][lane].sumsq;
          }
          p.partials[(int64_t)lbv * gridDim.x + blockIdx.x] = sp;
          __threadfence();
          atomicAdd(p.counters + lbv, 1u);
        }
        if (lane == 0) FX_TL(12, 0);
        __syncwarp();
      }
    }

    // ---- main jobs ----
    int pend = 0;                                  // 1 = a job of this group has its P.V in flight / finished
    int pend_slot = 0, pend_os = 0, pend_ph = 0, pend_n0 = 0, pend_b = 0, pend_h = 0, pend_xb = 0;
    float pend_sum = 0.f;
    int xb = 0;                                    // exchange buffer parity of this group's next job

    auto epilogue = [&]() {                        // O (fp32, TMEM) -> * 1/rowsum -> fp16 -> staging -> TMA store
      ptx::mbar_wait(BAR(B_PVDONE + pend_slot), (uint32_t)pend_ph);
      ptx::tc_fence_after();
      const uint32_t ta = tmem_base + lane_addr + C::col_o(pend_os);
#pragma unroll
      for (int ps = 0; ps < C::EPI_NPASS; ++ps) {
        float o[41];
        if constexpr (D == 40) {
          // columns [oc0, oc0 + 24) (the second half only uses 16 of them) and the row-sum column 40
          ptx::tmem_ld16_sync(ta + oc0, o);
          ptx::tmem_ld8_sync(ta + oc0 + 16, o + 16);
          tmem_ld1_sync(ta + 40, o + 40);
        } else if constexpr (D == 64) {
          ptx::tmem_ld32_sync(ta + oc0, o);
        } else if constexpr (D == 80) {
          tmem_ld40_sync(ta + oc0, o);
          tmem_ld1_sync(ta + 80, o + 40);
        } else {
          tmem_ld40_sync(ta + oc0 + ps * 40, o);
        }
        if (ps == C::EPI_NPASS - 1) {
          ptx::tc_fence_before();
          warp_arrive(BAR(B_OFREE + pend_os));     // O is in registers: the next P.V on this accumulator may start
        }
        float inv;
        if constexpr (C::ONES) inv = 1.f / o[40];
        else inv = 1.f / (pend_sum + xsum[((g * 2 + pend_xb) * 2 + (c ^ 1)) * 128 + row]);
        // the staging tile is free once the previous store of this warp has been read out (lane 0 owns the groups)
        if (lane == 0) bulk_wait_group_read0();
        __syncwarp();
        const int w8 = (D == 160) ? 5 : (ocn >> 3);      // 16-byte chunks this thread writes
#pragma unroll
        for (int ch = 0; ch < C::EPI_W / 8; ++ch) {
          if (ch < w8) {
            __align__(16) __half2 pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) pk[k] = __floats2half2_rn(o[ch * 8 + 2 * k] * inv, o[ch * 8 + 2 * k + 1] * inv);
            // staging rows are exactly ocn (or 40) columns wide for the store's box: pitch = box width
            *reinterpret_cast<uint4*>(smem_gen + stg_off + lane * (((D == 160) ? 40 : ocn) * 2) + ch * 16) =
                *reinterpret_cast<const uint4*>(pk);
          }
        }
        ptx::fence_proxy_async_smem();             // generic-proxy writes -> visible to the TMA (async proxy)
        __syncwarp();
        if (lane == 0 && pend_n0 < p.N) {          // rows >= N are clipped by the TMA
          ptx::tma_store_4d(c ? &tmo1 : &tmo0, smem0 + stg_off, oc0 + ps * 40, pend_h, pend_n0, pend_b);
          ptx::bulk_commit_group();
        }
      }
    };

    while (more) {
      if ((jb.i & 1) == g) {
        const int slot = jb.i % C::NS;
        const uint32_t ts = tmem_base + lane_addr + C::col_s(slot);
        ptx::mbar_wait(BAR(B_SREADY + slot), (uint32_t)((jb.i / C::NS) & 1));
        ptx::tc_fence_after();
        if ((sw & 7) == 0 && lane == 0) FX_TL(4, jb.i);
        float s[40];
        tmem_ld40_sync(ts + c * 40, s);
        // row max over this thread's columns (padded keys excluded), then over the row via the partner thread
        if constexpr (TT == 77) {
          if (c) { s[37] = -INFINITY; s[38] = -INFINITY; s[39] = -INFINITY; }
        } else {
#pragma unroll
          for (int j = 0; j < 40; ++j)
            if (c * 40 + j >= T) s[j] = -INFINITY;
        }
        float m0 = s[0], m1 = s[1], m2 = s[2], m3 = s[3];
#pragma unroll
        for (int j = 4; j < 40; j += 4) {
          m0 = fmaxf(m0, s[j]); m1 = fmaxf(m1, s[j + 1]); m2 = fmaxf(m2, s[j + 2]); m3 = fmaxf(m3, s[j + 3]);
        }
        float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        float* xm = xchg + ((g * 2 + xb) * 2) * 128;
        xm[c * 128 + row] = mx;
        ptx::named_bar_sync(pair_bar, 64);
        mx = fmaxf(mx, xm[(c ^ 1) * 128 + row]);
        const float nm = -mx * sl2;
        // p_j = 2^(s_j*sl2 - mx*sl2), UNNORMALISED, packed to fp16; O is scaled by 1/rowsum in the epilogue (fp32)
        float a0 = 0.f, a1 = 0.f;
        uint32_t pk[20];
#pragma unroll
        for (int j = 0; j < 40; j += 2) {
          const float e0 = ptx::ex2(fmaf(s[j], sl2, nm)), e1 = ptx::ex2(fmaf(s[j + 1], sl2, nm));
          const __half2 h = __floats2half2_rn(e0, e1);
          pk[j / 2] = *reinterpret_cast<const uint32_t*>(&h);
          if constexpr (!C::ONES) {
            const float2 f = __half22float2(h);       // sum exactly what the UMMA will multiply
            a0 += f.x; a1 += f.y;
          }
        }
        if constexpr (!C::ONES) xsum[((g * 2 + xb) * 2 + c) * 128 + row] = a0 + a1;
        // P (packed fp16) over the S columns it came from: this half owns P columns [20c, 20c + 20)
        ptx::tmem_st16_u32(ts + c * 20, pk);
        tmem_st4_u32(ts + c * 20 + 16, pk + 16);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        warp_arrive(BAR(B_PREADY + slot));
        if ((sw & 7) == 0 && lane == 0) FX_TL(5, jb.i);
        if (pend) epilogue();                      // overlaps with this job's P.V
        if ((sw & 7) == 0 && lane == 0) FX_TL(6, jb.i);
        pend = 1;
        pend_slot = slot;
        pend_os = jb.i % C::NO;
        pend_ph = (jb.m / C::NS) & 1;
        pend_n0 = jb.tile * kBM + (qd << 5);
        pend_b = jb.b;
        pend_h = jb.h;
        pend_xb = xb;
        pend_sum = a0 + a1;
        xb ^= 1;
      }
      more = jobs.next(jb);
    }
    if (pend) {
      if constexpr (!C::ONES) ptx::named_bar_sync(pair_bar, 64);   // the partner's row sum of the last job is written
      epilogue();
    }
    if (lane == 0) ptx::bulk_wait_group0();        // the staging tile must outlive the last store
    if ((sw & 7) == 0 && lane == 0) FX_TL(14, g);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FX_TL(15, 0);
  if (warp == 1) ptx::tmem_dealloc<512>(tmem_base);
  // the last CTA to leave resets the arrival counters for the next launch (every waiter has passed its barrier)
  if (threadIdx.x == 0 && nb > 0) {
    __threadfence();
    const unsigned prev = atomicAdd(p.counters + kMaxBatch, 1u);
    if (prev == gridDim.x - 1u) {
      for (int b = 0; b < kMaxBatch + 1; ++b) p.counters[b] = 0u;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
// Packed map [Bw, N, 32] fp16 viewed as (col, row, image); box = 32 x 128 x 1, 64-byte swizzle.
inline bool make_tmap_mpack(CUtensorMap* m, const void* base, int N, int Bw, int64_t batch_stride) {
  tc::EncodeTiledFn fn = tc::encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {(cuuint64_t)kMW, (cuuint64_t)N, (cuuint64_t)Bw};
  cuuint64_t strides[2] = {(cuuint64_t)kMW * 2, (cuuint64_t)batch_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)kMW, (cuuint32_t)kBM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    snprintf(tc::tc_error_buf(), 256, "cuTensorMapEncodeTiled(mpack) failed (CUresult %d): base %p N %d Bw %d stride %lld",
             (int)r, base, N, Bw, (long long)batch_stride);
  return r == CUDA_SUCCESS;
}

inline long long*& debug_timeline() {   // test infrastructure: device buffer [kFxTlTags][kFxTlIts] or null
  static long long* t = nullptr;
  return t;
}
inline int& debug_timeline_cta() {
  static int c = 0;
  return c;
}
inline unsigned*& debug_jobs_dump() {   // test infrastructure: device buffer the grouped-head kernel copies its job tables to
  static unsigned* p = nullptr;
  return p;
}
inline int& fused_variant() {        // 0 = grouped-head kernel at D = 40, 1 = per-head kernel everywhere (A/B timing)
  static int v = 0;
  return v;
}
inline int& debug_grid() {          // test infrastructure: cap the persistent grid (0 = number of SMs)
  static int g = 0;
  return g;
}
inline int fused_grid(int units) {
  int g = tc::num_sms();
  if (debug_grid() > 0 && debug_grid() < g) g = debug_grid();
  return units < g ? units : g;
}
// A CTA range may touch at most kMaxLocal biased images (partial slots in shared memory).
inline bool fused_range_ok(int B, int H, int tiles, int grid) {
  const long long units = (long long)B * H * tiles;
  const long long per_cta = (units + grid - 1) / grid;
  return per_cta <= (long long)(kMaxLocal - 1) * tiles * H;
}
inline size_t fused_workspace_bytes() {
  return 512 + (size_t)kMaxBatch * 2048 * sizeof(StatPartial) / 8;   // counters | per-image maxima | [32][256] partial slots
}

template <int D>
cudaError_t launch_fused(const XattnParams& x, const void* mpack, int64_t mpack_bs, int Bw, const int8_t* cidx,
                         cudaStream_t s) {
  using C = Cfg<D>;
  CUtensorMap tq, tk, tv, tm, to0, to1;
  const int kB = x.k_bs > 0 ? x.B : 1;
  if (!tc::make_tmap(&tq, x.q, D, x.H, x.N, x.B, x.q_rs, x.q_bs, kBM) ||
      !tc::make_tmap(&tk, x.k, D, x.H, x.T, kB, x.k_rs, x.k_bs, kTP) ||
      !tc::make_tmap(&tv, x.v, D, x.H, x.T, kB, x.k_rs, x.k_bs, kTP))
    return cudaErrorInvalidValue;
  if (mpack != nullptr) {
    if (!make_tmap_mpack(&tm, mpack, x.N, Bw, mpack_bs)) return cudaErrorInvalidValue;
  } else {
    tm = tq;                                       // never dereferenced: no image is biased
  }
  const int w0 = (D == 160) ? 40 : C::C0, w1 = (D == 160) ? 40 : C::C1;
  if (!tc::make_tmap_out(&to0, x.out, D, x.H, x.N, x.B, x.o_rs, x.o_bs, 32, w0, false) ||
      !tc::make_tmap_out(&to1, x.out, D, x.H, x.N, x.B, x.o_rs, x.o_bs, 32, w1, false))
    return cudaErrorInvalidValue;
  FxParams fp;
  fp.x = x;
  fp.cidx = cidx;
  fp.tiles = ceil_div(x.N, kBM);
  fp.units = x.B * fp.tiles * x.H;
  fp.k_batched = x.k_bs > 0 ? 1 : 0;
  fp.grid = fused_grid(fp.units);
  fp.timeline = debug_timeline();
  fp.tl_cta = debug_timeline_cta();
  fp.hg = x.H;
  fp.jobs_dump = nullptr;
  if (!fused_range_ok(x.B, x.H, fp.tiles, fp.grid) || !fused_units_ok(fp.units, fp.grid)) return cudaErrorInvalidConfiguration;
  static bool attr_set[tc::kMaxDevices] = {false};
  if (!attr_set[tc::cur_device()]) {
    cudaError_t e = cudaFuncSetAttribute(xattn_fused_kernel<D, 77>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(xattn_fused_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return e;
    attr_set[tc::cur_device()] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(fp.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;     // all CTAs co-resident: the in-kernel grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (x.T == 77) return cudaLaunchKernelEx(&cfg, xattn_fused_kernel<D, 77>, tq, tk, tv, tm, to0, to1, fp);
  return cudaLaunchKernelEx(&cfg, xattn_fused_kernel<D, 0>, tq, tk, tv, tm, to0, to1, fp);
}

// Host replay of the job lists (test infrastructure): out[job] = {cta, i, kind, m, b, h, tile, biased, li, gi} for every job
// of every CTA; returns the number of jobs written.
inline int fused_schedule_host(int B, int H, int tiles, int grid, const int* wmap_index, int* out, int max_jobs) {
  if (B <= 0 || B > kMaxBatch || H <= 0 || tiles <= 0 || grid <= 0) return -1;
  int img[kMaxBatch];
  int nb = 0;
  for (int b = 0; b < B; ++b) if (wmap_index[b] >= 0) img[nb++] = b;
  int nu = 0;
  for (int b = 0; b < B; ++b) if (wmap_index[b] < 0) img[nb + nu++] = b;
  const int units = B * H * tiles;
  int row = 0;
  for (int cta = 0; cta < grid; ++cta) {
    int u0, u1;
    fx_range(cta, grid, units, u0, u1);
    FxJobs jobs(u0, u1 - u0, B, H, tiles, nb, img);
    FxJob jb;
    while (jobs.next(jb)) {
      if (row >= max_jobs) return -2;
      int* o = out + 10 * (row++);
      o[0] = cta; o[1] = jb.i; o[2] = jb.kind; o[3] = jb.m; o[4] = jb.b; o[5] = jb.h; o[6] = jb.tile; o[7] = jb.biased;
      o[8] = jb.li; o[9] = jb.gi;
    }
  }
  return row;
}
inline int fused_cta_has_image_host(int cta, int grid, int B, int H, int tiles, const int* wmap_index, int b) {
  int nb = 0, pos = -1;
  for (int i = 0; i < B; ++i)
    if (wmap_index[i] >= 0) { if (i == b) pos = nb; ++nb; }
  if (pos < 0) return 0;
  const int nu = B - nb, np = nb < nu ? nb : nu;
  return fx_cta_has_image(cta, grid, B * H * tiles, pos, H, tiles, np) ? 1 : 0;
}

}  // namespace fx
}  // namespace pww
