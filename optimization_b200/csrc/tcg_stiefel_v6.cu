// Persistent fused truncated-CG for the Stiefel trace-minimisation Hessian, v6: warp-specialised roles with
// their own register budgets (setmaxnreg), every bulk operand of phase A staged by the TMA unit, the TMEM
// read-back done by the warps that consume it, p handed from the slicing role to the multiplying role through
// shared memory.  Same mathematics, reductions and scalar logic as tcg_stiefel_tc_kernel (v4, tcg_stiefel_tc.cu;
// reference line map in tcg.cuh): each CG iteration of IterativeSolvers.h:285-422 is two fused phases separated
// by an exact grid-wide (machine-wide) reduction.
//
// 20 warps = 5 warp groups:
//   S (warps 0-3, 24 registers)   : service lanes only -- warp 0 lane 0 = TMA producer (r / p_old tiles of the next
//                                   block into a 64 KB stage, the block's three int8 digit planes of A, L2 prefetch
//                                   of what follows); warp 1 lane 0 = MMA issuer (13 tcgen05.mma kind::i8 per block
//                                   into one of two 256-column TMEM accumulator sets, tcgen05.commit -> mbarrier).
//                                   S takes no part in the reductions, the scalar stage or phase B: it meets the
//                                   other roles at ONE CTA-wide barrier per iteration.
//   L (warps 4-7, 96 registers)   : r, p_old from the stage -> p = -r + beta p_old (l.420; written back to HBM for the
//                                   rows this CTA owns and, in place of p_old, to the stage), <p,p>, <p,r>, block
//                                   maximum, seven balanced int8 digit slices of p straight into the UMMA K-major
//                                   SWIZZLE_128B operand image -- one pass.
//   M (warps 8-15, 120 registers) : one 16-row group of one 64-row half block per warp: its elements of p from the
//                                   stage (then the half is released: the next block's p_old streams in), TMEM ->
//                                   registers in the mma.sync accumulator arrangement (tcgen05.ld 16x256b), integer
//                                   recombination = Z = A p; W = Z - p Lambda (the solve runs in the eigenbasis of S:
//                                   an elementwise shift), W written back and staged per 64-row half, <p,W>, <W,W>.
//   G (warps 16-19, 120 registers): projection Gram Y^T W of the staged half (fp64 tensor cores, exact fixed-point
//                                   accumulation), Y of the next half fetched with cp.async meanwhile.
// Hand-offs through mbarriers only.  Ownership is by 64-row HALF blocks (balanced to 1/11 instead of 1/6 of a CTA's
// work): a block shared by two CTAs is sliced and multiplied by both (the MMA needs all 128 rows of p as K),
// everything else -- p / W stores, the fp64 MMAs, the Gram, all partial sums -- is done for the owned half only.
// The A images are row-permuted (tc_row_of_lane) so that either half occupies 16 lanes of every TMEM lane quarter.
// Phase B (l.374-408): L, M and G warps, one TMA-fed 10 KB strip slot each (16 slots per SM as in v4).
#include "tcg.cuh"
#include "stiefel_dev.cuh"
#include "tc_common.cuh"

namespace ob200 {
using namespace tc;

constexpr int V6_THREADS = 640;
constexpr int V6_WORK = 512;                                   // L + M + G threads (relative id = tid - 128)
// shared-memory map (bytes from the 1024-aligned base); phase B aliases the phase-A operand space
constexpr uint32_t V6_A = 0;                                   // 48 KB int8 digit planes of A
constexpr uint32_t V6_Q = V6_A + TC_ABLOCK;                    // 28 KB int8 digit image of p
constexpr uint32_t V6_TILE = ST_NB * ST_P * 8;                 // 32 KB dense fp64 tile
constexpr uint32_t V6_HALF = V6_TILE / 2;                      // 16 KB: one 64-row half of it
constexpr uint32_t V6_R = V6_Q + TC_QBYTES;                    // r tile (TMA)
constexpr uint32_t V6_PO = V6_R + V6_TILE;                     // p_old tile (TMA), overwritten in place with p
constexpr uint32_t V6_WB = ST_NB * WS * 8;                     // 36 KB padded tile
constexpr uint32_t V6_W = V6_PO + V6_TILE;                     // W tile (stride WS)
constexpr uint32_t V6_Y = V6_W + V6_WB;                        // Y tile (stride WS)
constexpr uint32_t V6_S = V6_Y + V6_WB;                        // -lambda: the 32 eigenvalues of S, negated (region kept at p x WS doubles)
constexpr uint32_t V6_ACC = V6_S + ST_P * WS * 8;              // CTA Kulisch accumulators (5 scalars)
constexpr uint32_t V6_NACC = 5;
constexpr uint32_t V6_BAR = V6_ACC + V6_NACC * KUL_STRIDE * 8; // mbarriers
constexpr uint32_t V6_NBAR = 40;
constexpr uint32_t V6_MISC = V6_BAR + V6_NBAR * 8;
constexpr uint32_t V6_MISC_BYTES = 768;
constexpr uint32_t V6_TOTAL = V6_MISC + V6_MISC_BYTES;
// phase B
constexpr uint32_t V6_STRIP_TILE = 8 * ST_P * 8;               // 2 KB
constexpr uint32_t V6_SLOT = 5 * V6_STRIP_TILE;                // 10 KB: W, s, p, r, Y tiles of one 8-row strip
constexpr uint32_t V6_NSLOT = 16;
constexpr uint32_t V6_GRAW = V6_NSLOT * V6_SLOT;               // 8 KB scratch (inside the W tile region)
constexpr uint32_t V6_G = V6_Y;                                // -sym(G), stride GS (inside the Y tile region)
static_assert(V6_GRAW >= V6_W && V6_GRAW + 8192 <= V6_Y, "G scratch must sit in the W tile region");
static_assert(ST_P * GS * 8 <= V6_WB, "G must fit the Y tile region");
static_assert(V6_TOTAL + 1024 <= 232448, "shared memory budget (227 KB per CTA)");

// (the stage works in 64-row halves: r / p_old of half h stream in while the other half is being sliced)
enum { B6_A_FULL = 2, B6_Q_FULL = 3, B6_MMA_DONE = 4 /*,5*/, B6_TMEM_EMPTY = 6 /*,7*/,
       B6_SLOT = 8 /* .. 23 */, B6_W_FULL = 24 /*,25*/, B6_W_EMPTY = 26 /*,27*/,
       B6_RP_FULL = 28 /*,29*/, B6_R_EMPTY = 30 /*,31*/, B6_PO_EMPTY = 32 /*,33*/, B6_P_FULL = 34 /*,35*/ };

struct V6Misc {
  CgShared sh;
  double s_part[16];
  double s_invq, s_q;
  int s_fe[5];
  int s_E[4];
  int s_next_strip;
  int done;
  int s_ok, s_last;
  uint32_t s_tmem;
  unsigned long long s_stamp[4];
};
static_assert(sizeof(V6Misc) <= V6_MISC_BYTES, "misc block");

__device__ __forceinline__ void bulk_prefetch_l2_v6(const void *g, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global_v6() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void bar_cta() { asm volatile("bar.sync 0, 640;" ::: "memory"); }    // all five warp groups
__device__ __forceinline__ void bar_work() { asm volatile("bar.sync 1, 512;" ::: "memory"); }   // L + M + G

__device__ __forceinline__ void strip_fetch6(unsigned char *slot, uint64_t *bar, int sidx, unsigned n_rows,
                                             const double *W, const double *S, const double *Pn, const double *R,
                                             const double *Y) {
  const unsigned row0 = (unsigned)sidx * 8u;
  const unsigned rows = n_rows - row0 < 8u ? n_rows - row0 : 8u;
  const uint32_t bytes = rows * ST_P * (uint32_t)sizeof(double);
  const size_t off = (size_t)row0 * ST_P;
  mbar_expect_tx(bar, 5 * bytes);
  bulk_g2s(slot, W + off, bytes, bar);
  bulk_g2s(slot + V6_STRIP_TILE, S + off, bytes, bar);
  bulk_g2s(slot + 2 * V6_STRIP_TILE, Pn + off, bytes, bar);
  bulk_g2s(slot + 3 * V6_STRIP_TILE, R + off, bytes, bar);
  bulk_g2s(slot + 4 * V6_STRIP_TILE, Y + off, bytes, bar);
}

// Recombination in the fragment arrangement of tcgen05.ld 16x256b (8 columns): this thread's 2 rows x 2 columns of
// sum_u D_u 2^(-8u)   (out[0], out[1] = row m, columns 2j, 2j+1;  out[2], out[3] = row m + 8).
__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void recombine_frag8(uint32_t taddr, double (&out)[4]) {
  long long part[2][4];
#pragma unroll
  for (int gq = 0; gq < 2; ++gq) {                            // accumulators 0..3 -> part[0], 4..7 -> part[1]
    uint32_t v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) tmem_ld_16x256b_x1(taddr + (4 * gq + u) * TC_N, v[u]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 4; ++c)
      part[gq][c] = (long long)(int)v[3][c] + ((long long)(int)v[2][c] << 8) + ((long long)(int)v[1][c] << 16) +
                    ((long long)(int)v[0][c] << 24);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {   // |part| < 2^48: exact int64 -> double through the 2^52 + 2^51 bias
    const double d1 = __longlong_as_double(0x4338000000000000ll + part[1][c]) - 6755399441055744.0;
    const double d0 = __longlong_as_double(0x4338000000000000ll + part[0][c]) - 6755399441055744.0;
    out[c] = fma(d1, 0x1p-32, d0) * 0x1p-24;
  }
}

// Grid-wide reduction barrier of the 512 work threads (rt = relative thread id): the protocol of grid_reduce_barrier
// (common.cuh) for ONE GPU, with the CTA-level synchronisation on named barrier 1.  Row-sharded runs instantiate the
// kernel with MULTI = true and use grid_reduce_barrier_multi below; the one-GPU instantiation carries no exchange code:
// less to hold in the instruction cache of the persistent loop (3 800 instructions, 3 us per iteration).
__device__ __forceinline__ bool grid_reduce_barrier_w(V6Misc &ms, int rt, unsigned *counter, unsigned &gen,
                                                      int *abort_flag, const CommDev &cm, unsigned long long gphase,
                                                      u64 *set, int off, int count, RedView &view,
                                                      unsigned long long *stamps) {
  (void)gphase; (void)off; (void)count;
  bar_work();
  if (rt == 0) {
    gen += 1;
    const unsigned target = gen * gridDim.x;
    __threadfence();
    if (stamps) stamps[0] = globaltimer_ns();
    atom_add_acqrel_u32(counter, 1u);
    int ok = cm.world == 1;
    unsigned spins = 0;
    while (ok && ld_acquire_u32(counter) < target) {
      if (++spins > (1u << 24)) {
        if (*((volatile int *)abort_flag) || spins > (1u << 25)) ok = 0;
      }
    }
    if (stamps) stamps[1] = globaltimer_ns();
    if (!ok) atomicExch(abort_flag, 1);
    __threadfence();
    ms.s_ok = ok;
  }
  bar_work();
  view.world = 1;
  view.stride = (size_t)cm.words_per_set;
  view.base0 = set;
  return ms.s_ok != 0;
}

// The same with the machine-wide exchange of grid_reduce_barrier (common.cuh): row-sharded runs (instantiated only in the
// MULTI variant of the kernel, so the one-GPU loop stays lean).
__device__ __forceinline__ bool grid_reduce_barrier_multi(V6Misc &ms, int rt, unsigned *counter, unsigned &gen,
                                                      int *abort_flag, const CommDev &cm, unsigned long long gphase,
                                                      u64 *set, int off, int count, RedView &view,
                                                      unsigned long long *stamps) {
  bar_work();
  if (rt == 0) {
    gen += 1;
    const unsigned target = gen * gridDim.x;
    // (gpu scope: what this phase wrote is consumed on this GPU only; the words that travel are re-stored to the peers by
    // the publishing CTA below and released at system scope there)
    __threadfence();
    if (stamps) stamps[0] = globaltimer_ns();
    const unsigned old = atom_add_acqrel_u32(counter, 1u);
    int ok = 1;
    ms.s_last = (old + 1u == target);
    if (cm.world == 1) {
      unsigned spins = 0;
      while (ld_acquire_u32(counter) < target) {
        if (++spins > (1u << 24)) {
          if (*((volatile int *)abort_flag) || spins > (1u << 25)) { ok = 0; break; }
        }
      }
      if (stamps) stamps[1] = globaltimer_ns();
      if (!ok) atomicExch(abort_flag, 1);
      __threadfence();
    }
    ms.s_ok = ok;
  }
  bar_work();
  view.world = cm.world;
  view.stride = (size_t)cm.words_per_set;
  if (cm.world == 1) {
    view.base0 = set;
    return ms.s_ok != 0;
  }
  const int slot = (int)(gphase % ACC_SLOTS);
  const size_t slot_off = (size_t)(slot * MAX_RANKS) * cm.words_per_set;
  if (ms.s_last) {   // CTA-uniform: this CTA completed the local reduction -> publish it to every rank
    __threadfence();
    constexpr int PUB_MAX = 6;                       // count <= PUB_MAX * 512
    u64 wv[PUB_MAX];
#pragma unroll
    for (int k = 0; k < PUB_MAX; ++k) {
      const int i = rt + k * V6_WORK;
      wv[k] = (i < count) ? __ldcg(set + off + i) : 0ull;
    }
    for (int r = 0; r < cm.world; ++r) {
      u64 *dst = cm.inbox[r] + slot_off + (size_t)cm.rank * cm.words_per_set + off;
#pragma unroll
      for (int k = 0; k < PUB_MAX; ++k) {
        const int i = rt + k * V6_WORK;
        if (i < count) dst[i] = wv[k];
      }
    }
    bar_work();
    if (rt < cm.world) st_release_sys_u64(cm.flags[rt] + slot * MAX_RANKS + cm.rank, gphase + 1ull);
  }
  if (rt < cm.world) {
    const unsigned long long *f = cm.flags[cm.rank] + slot * MAX_RANKS + rt;
    unsigned spins = 0;
    while (ld_acquire_sys_u64(f) < gphase + 1ull) {
      if (++spins > (1u << 24)) {
        if (*((volatile int *)abort_flag) || spins > (1u << 25)) { atomicExch(abort_flag, 1); break; }
      }
    }
  }
  bar_work();
  if (rt == 0) {
    if (stamps) stamps[1] = globaltimer_ns();
    ms.s_ok = (*((volatile int *)abort_flag) == 0);
  }
  bar_work();
  view.base0 = cm.inbox[cm.rank] + slot_off;
  return ms.s_ok != 0;
}

__device__ __forceinline__ void flush_scalars_w(u64 *sacc, u64 *gacc, int nscal, int rt) {
  for (int i = rt; i < nscal * KUL_STRIDE; i += V6_WORK) {
    const u64 v = sacc[i];
    if (v) {
      atomicAdd(gacc + i, v);
      sacc[i] = 0;
    }
  }
}

extern __shared__ __align__(1024) unsigned char v6_smem_raw[];

// debug timeline (CTA 0, third block of an iteration, one lane per role): slot <- globaltimer
#ifdef OB200_TIMELINE_BUILD
#define TL6(slot) do { if (a.dbg && blockIdx.x == 0 && i == 2) a.dbg[4096 + (slot)] = globaltimer_ns(); } while (0)
#define TL6B(slot) do { if (a.dbg && blockIdx.x == 0) a.dbg[4096 + (slot)] = globaltimer_ns(); } while (0)
#define TLI(e) do { if (a.dbg && blockIdx.x == 0 && i >= 0 && i < 8) a.dbg[4096 + 64 + 8 * i + (e)] = globaltimer_ns(); } while (0)
#else
#define TLI(e) do { } while (0)
#define TL6(slot) do { } while (0)
#define TL6B(slot) do { } while (0)
#endif

struct V6Part {   // ownership of this CTA
  unsigned n_rows, h0, h1, bfirst;
  int nb_local;
};
__device__ __forceinline__ V6Part v6_partition(unsigned long long n_rows_ull) {
  V6Part q;
  q.n_rows = (unsigned)n_rows_ull;
  const unsigned nhalf = (q.n_rows + 63u) >> 6;
  q.h0 = (unsigned)((unsigned long long)nhalf * blockIdx.x / gridDim.x);
  q.h1 = (unsigned)((unsigned long long)nhalf * (blockIdx.x + 1ull) / gridDim.x);
  q.bfirst = q.h0 >> 1;
  q.nb_local = (q.h1 > q.h0) ? (int)(((q.h1 - 1u) >> 1) - q.bfirst + 1u) : 0;
  return q;
}

// ===== S: TMA producer (warp 0) and MMA issuer (warp 1), one elected lane each; nothing else =====
__device__ __forceinline__ void v6_run_service(const TcgCommon &a, const StiefelArgs &st, const unsigned char *planes,
                                               unsigned char *base) {
  unsigned char *Asm = base + V6_A;
  unsigned char *Qsm = base + V6_Q;
  uint64_t *mb = reinterpret_cast<uint64_t *>(base + V6_BAR);
  volatile V6Misc &ms = *reinterpret_cast<V6Misc *>(base + V6_MISC);
  const int tid = threadIdx.x;
  const V6Part pt = v6_partition(st.n_rows);
  const unsigned n_rows32 = pt.n_rows;
  const uint32_t tmem_base = ms.s_tmem;
  unsigned use = 0;
  for (;;) {
    bar_cta();
    if (ms.done) break;
    const unsigned long long k = ms.sh.k;
    if (tid == 0) {
      const double *p_old = (k & 1ull) ? a.p1 : a.p0;
      fence_proxy_async_global_v6();   // r / p written with generic stores by other CTAs (ordered by the grid barrier)
      for (int i = 0; i < pt.nb_local; ++i) {
        const unsigned u = use + i, b = pt.bfirst + i, r0 = b * ST_NB;
        TL6(0);
        // demand loads first (the TMA unit serves its queue in order), half by half: r into the half L has read, p_old
        // into the half M has taken p from
#pragma unroll 1
        for (unsigned h = 0; h < 2; ++h) {
          const unsigned rh = r0 + 64u * h;
          const unsigned rows_h = rh < n_rows32 ? (n_rows32 - rh < 64u ? n_rows32 - rh : 64u) : 0u;
          const uint32_t bytes_h = rows_h * ST_P * (uint32_t)sizeof(double);
          const size_t off_h = (size_t)rh * ST_P;
          if (u > 0) mbar_wait_guarded(&mb[B6_R_EMPTY + h], (u - 1) & 1);       // L has read r of this half (previous block)
          if (h == 0) TL6(1);
          mbar_expect_tx(&mb[B6_RP_FULL + h], k ? 2 * bytes_h : bytes_h);
          if (bytes_h) bulk_g2s(base + V6_R + h * V6_HALF, a.r + off_h, bytes_h, &mb[B6_RP_FULL + h]);
          if (k) {
            if (u > 0) mbar_wait_guarded(&mb[B6_PO_EMPTY + h], (u - 1) & 1);    // M has taken p of this half (previous block)
            fence_proxy_async_smem();   // L's generic-proxy stores of p into this half are ordered before the async-proxy write
            if (bytes_h) bulk_g2s(base + V6_PO + h * V6_HALF, p_old + off_h, bytes_h, &mb[B6_RP_FULL + h]);
          }
          if (h == 0) TL6(2);
        }
        if (u > 0) mbar_wait_guarded(&mb[B6_MMA_DONE + ((u - 1) & 1)], ((u - 1) >> 1) & 1);   // A image free
        mbar_expect_tx(&mb[B6_A_FULL], TC_ABLOCK);
        bulk_g2s(Asm, planes + (size_t)b * TC_ABLOCK, TC_ABLOCK, &mb[B6_A_FULL]);
        TL6(3);
#ifndef OB200_PF
#define OB200_PF 0   // bulk L2 prefetches off: measured, the demand loads of the next half queue behind them in the TMA unit
#endif
        if (OB200_PF && i + 1 < pt.nb_local) {   // L2 prefetch of what the next blocks need: A planes, Y; then r / p_old one further
          const unsigned rn = r0 + ST_NB;
          const unsigned rows1 = n_rows32 - rn < ST_NB ? n_rows32 - rn : ST_NB;
          const uint32_t bytes1 = rows1 * ST_P * (uint32_t)sizeof(double);
          bulk_prefetch_l2_v6(planes + (size_t)(b + 1) * TC_ABLOCK, TC_ABLOCK);
          bulk_prefetch_l2_v6(st.Y + (size_t)rn * ST_P, bytes1);
          if ((OB200_PF & 2) && i + 2 < pt.nb_local) {
            const unsigned r2 = rn + ST_NB;
            const unsigned rows2 = n_rows32 - r2 < ST_NB ? n_rows32 - r2 : ST_NB;
            const uint32_t bytes2 = rows2 * ST_P * (uint32_t)sizeof(double);
            bulk_prefetch_l2_v6(a.r + (size_t)r2 * ST_P, bytes2);
            if (k) bulk_prefetch_l2_v6(p_old + (size_t)r2 * ST_P, bytes2);
          }
        }
        TL6(9);
      }
    } else if (tid == 32) {
      for (int i = 0; i < pt.nb_local; ++i) {
        const unsigned u = use + i;
        TL6(4);
        mbar_wait_guarded(&mb[B6_Q_FULL], u & 1);
        TL6(5);
        mbar_wait_guarded(&mb[B6_A_FULL], u & 1);
        TL6(6);
        if (u >= 2) mbar_wait_guarded(&mb[B6_TMEM_EMPTY + (u & 1)], ((u >> 1) - 1) & 1);   // M has drained this accumulator set
        TL6(7);
        tc_fence_after();
        issue_block_mmas(smem_u32(Asm), smem_u32(Qsm), tmem_base + (u & 1) * TC_TMEM_COLS);
        umma_commit(&mb[B6_MMA_DONE + (u & 1)]);
        TL6(8);
      }
    }
    __syncwarp();
    use += (unsigned)pt.nb_local;
  }
}

// role: 1 = L, 2 = M, 3 = G (warp-uniform).  ONE instance of the loop for the three work roles, all on the same register
// budget: the phase-A role bodies are separate branches, everything else (reductions, scalar stage, phase B) is shared
// code -- three per-role copies of it evicted each other from the instruction cache (stall_no_instruction).
// (historical note) The whole CG loop used to be instantiated per role so that each warp group's code was compiled
// against its own register budget.
// MROLE: the instance for the M role (its own register budget); the other instance serves L and G (runtime `role`).
template <bool MROLE, bool MULTI>
__device__ __forceinline__ void v6_run(const int role_rt, const TcgCommon &a, const StiefelArgs &st, const int *plane_exp,
                                       unsigned char *base) {
  const int role0 = MROLE ? 2 : role_rt;
  unsigned char *Qsm = base + V6_Q;
  const unsigned char *Rsm = base + V6_R;
  unsigned char *POsm = base + V6_PO;
  double *Wsm = reinterpret_cast<double *>(base + V6_W);
  double *Ysm = reinterpret_cast<double *>(base + V6_Y);
  const double *Ssm = reinterpret_cast<const double *>(base + V6_S);
  double *Gsm = reinterpret_cast<double *>(base + V6_G);
  u64 *sacc = reinterpret_cast<u64 *>(base + V6_ACC);
  uint64_t *mb = reinterpret_cast<uint64_t *>(base + V6_BAR);
  V6Misc &ms = *reinterpret_cast<V6Misc *>(base + V6_MISC);
  CgShared &sh = ms.sh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rt = tid - 128, rw = warp - 4;                     // relative ids among the work threads / warps
  const int m = lane >> 2, j = lane & 3;
  const uint32_t tmem_base = ms.s_tmem;

  const V6Part pt = v6_partition(st.n_rows);
  const unsigned n_rows32 = pt.n_rows, h0 = pt.h0, h1 = pt.h1, bfirst = pt.bfirst;
  const int nb_local = pt.nb_local;
  const unsigned row_lo = h0 * 64u;
  const unsigned row_hi = (h1 * 64u < n_rows32) ? h1 * 64u : n_rows32;
  // phase B: 8-row strips split evenly over the CTAs, handed to the work warps dynamically, top-down
  const unsigned nstrips = (n_rows32 + 7u) >> 3;
  const int s_lo = (int)((unsigned long long)nstrips * blockIdx.x / gridDim.x);
  const int s_hi = (int)((unsigned long long)nstrips * (blockIdx.x + 1ull) / gridDim.x);
  unsigned char *slot = base + rw * V6_SLOT;
  uint64_t *sb = &mb[B6_SLOT + rw];
  unsigned bpar = 0;               // parity of this warp's strip-slot mbarrier
  unsigned gen = 0, phase = 0;
  unsigned use = 0;                // blocks processed so far by this CTA (mbarrier phase bookkeeping)
  unsigned wcnt = 0;               // M: fills of this warp's half-block W staging area; G: drains, 16 bits per half
  int exit_reason = -1;
  unsigned long long dbg_prev = 0;

  for (;;) {
    // (broadcast from lane 0: the compiler then knows the role is warp-uniform and emits no reconvergence scaffolding
    // around the shuffles / votes inside role-dependent branches)
    const int role = MROLE ? 2 : __shfl_sync(0xffffffffu, role0, 0);
    if (exit_reason == -1) {
      if (sh.k >= a.max_iterations) exit_reason = 1;
      else if (sqrt(sh.rv) <= a.target) exit_reason = 0;
    }
    if (rt == 0) ms.done = (exit_reason != -1);
    bar_cta();                     // S learns here whether another phase A follows (and reads sh.k)
    if (exit_reason != -1) break;
    const unsigned long long k = sh.k;
    const double beta = sh.beta;
    double *p_new = (k & 1ull) ? a.p0 : a.p1;
    const double inv_q = ms.s_invq, q = ms.s_q;

    // ------------------------------ phase A ------------------------------
    u64 *set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    {   // recycle the set used two phases from now (every CTA clears its slice; a grid barrier intervenes)
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      const int per = (ACC_WORDS + gridDim.x - 1) / gridDim.x;
      const int z0 = per * blockIdx.x;
      for (int i = rt; i < per && z0 + i < ACC_WORDS; i += V6_WORK) nxt[z0 + i] = 0;
    }
    if (!MROLE && role == 1) {
      // ===== L: p = -r + beta p_old, digit slices -- ONE pass over the staged block =====
      // The digit slices need a scale 2^E with |p| < 2^E over the block BEFORE the first element is cut.  Instead of a
      // maximum pass and a second pass, E comes from a bound that is known when the block arrives:
      //   max |p_new| <= max |r| + |beta| max |p_old|      (both maxima per 128-row block, exact, kept in blk_stats:
      //   max |r| from phase B of the previous iteration / the init kernel, max |p_old| from this role one iteration ago)
      // It is a deterministic function of exactly reduced data (identical on every CTA / GPU that handles the block) and
      // at most a few bits above the true maximum: p is quantised to 2^(E-54), i.e. no coarser than ~2^-52 of the
      // block maximum.
      const int t = rt, cp = t & 15, g = t >> 4;             // columns cp, cp + 16 ; rows 8g .. 8g+7 of a 64-row half
      // (a half warp reads / writes 16 consecutive doubles of a row: conflict-free shared memory, full HBM sectors)
      FixAcc fa0 = {0, 0}, fa1 = {0, 0};                     // <p,p>, <p,r>
      const int fe0 = ms.s_fe[SC_PP], fe1 = ms.s_fe[SC_PR];
      const double fq0 = scalbn(1.0, 90 - fe0), fq1 = scalbn(1.0, 90 - fe1);
      unsigned ovf = 0;
      const unsigned long long nbs = a.nblk_stats;
      const unsigned long long *Rcur = a.blk_stats + (k & 1ull) * nbs;
      unsigned long long *Rnext = a.blk_stats + ((k + 1ull) & 1ull) * nbs;
      const unsigned long long *Pcur = a.blk_stats + 2 * nbs + (k & 1ull) * 4 * nbs;
      unsigned long long *Pnext = a.blk_stats + 2 * nbs + ((k + 1ull) & 1ull) * 4 * nbs;
      const double abeta = fabs(beta);
      const bool kk = k != 0;
      // bound statistics of a block (uniform over the CTA: every thread reads the same five words); fetched one block
      // ahead so that the L2 round trip is off the critical path
      auto block_bound = [&](unsigned b) {
        const double rmax = __longlong_as_double((long long)__ldcg(Rcur + b));
        const ulonglong2 pm01 = __ldcg(reinterpret_cast<const ulonglong2 *>(Pcur + 4 * (size_t)b));
        const ulonglong2 pm23 = __ldcg(reinterpret_cast<const ulonglong2 *>(Pcur + 4 * (size_t)b + 2));
        const double pmax = fmax(fmax(__longlong_as_double((long long)pm01.x), __longlong_as_double((long long)pm01.y)),
                                 fmax(__longlong_as_double((long long)pm23.x), __longlong_as_double((long long)pm23.y)));
        return fma(abeta, pmax, rmax) * (1.0 + 0x1p-40);
      };
      double bound_next = nb_local > 0 ? block_bound(bfirst) : 0.0;
      for (int i = 0; i < nb_local; ++i) {
        const unsigned u = use + i, b = bfirst + i, r0 = b * ST_NB;
        const double bound = bound_next;
        if (i + 1 < nb_local) bound_next = block_bound(b + 1);
        const int E = (bound > 0.0) ? (int)((__double_as_longlong(bound) >> 52) & 0x7ff) - 1023 + 1 : 0;
        const double scale = scalbn(1.0, 54 - E);
        if (t == 0 && 2u * b >= h0) __stcg(Rnext + b, 0ull);           // max |r| of the NEXT iteration starts from zero
        if (t == 0) TL6(10);
        if (u > 0) mbar_wait_guarded(&mb[B6_MMA_DONE + ((u - 1) & 1)], ((u - 1) >> 1) & 1);   // digit image free
        int mxb = 0;
#pragma unroll 1
        for (unsigned h = 0; h < 2; ++h) {                   // the two 64-row halves of the block, as they arrive
          const bool mine = 2u * b + h >= h0 && 2u * b + h < h1;
          const unsigned char *rrow = Rsm + h * V6_HALF + (8u * g) * 256u + 8u * cp;
          unsigned char *prow = POsm + h * V6_HALF + (8u * g) * 256u + 8u * cp;
          mbar_wait_guarded(&mb[B6_RP_FULL + h], u & 1);
          if (t == 0 && h == 0) { TL6(11); TLI(0); }
          if (!kk && u > 0) mbar_wait_guarded(&mb[B6_PO_EMPTY + h], (u - 1) & 1);   // first iteration: no TMA into the p tile orders this
          if (t == 0 && h == 0) TL6(12);
#pragma unroll 1
          for (int z = 0; z < 2; ++z) {                      // one column of the pair at a time (not unrolled: code size)
            uint32_t lo[8], hi[8];
            double pp0 = 0.0, pp1 = 0.0, pr0 = 0.0, pr1 = 0.0;
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) {
              const unsigned grow = r0 + 64u * h + 8u * g + ii;
              const bool valid = grow < n_rows32, ok = valid && mine;
              // branch-free (selects): rows beyond n and, in the first iteration, the p_old tile hold stale shared memory
              double rv = *reinterpret_cast<const double *>(rrow + ii * 256u + 128u * z);
              double po = *reinterpret_cast<const double *>(prow + ii * 256u + 128u * z);
              rv = valid ? rv : 0.0;
              po = (valid && kk) ? po : 0.0;
              const double pv = fma(beta, po, -rv);                    // l.420; first iteration: beta = 0, p = -r (l.256)
              *reinterpret_cast<double *>(prow + ii * 256u + 128u * z) = pv;   // the M role takes its fragments from here
              if (ok) __stcg(p_new + (size_t)grow * ST_P + cp + 16 * z, pv);
              const double pm = ok ? pv : 0.0;
              if (ii & 1) { pp1 = fma(pm, pm, pp1); pr1 = fma(pm, rv, pr1); }
              else        { pp0 = fma(pm, pm, pp0); pr0 = fma(pm, rv, pr0); }
              mxb = max(mxb, __double2hiint(pv) & 0x7fffffff);   // high word of |pv|: an ordered integer
              const unsigned long long uu = ((unsigned long long)__double2ll_rn(pv * scale) + TC_DIGIT_BIAS) ^ TC_DIGIT_BIAS;
              lo[ii] = (uint32_t)uu;
              hi[ii] = (uint32_t)(uu >> 32);
            }
            if (t == 0 && h == 0) TL6(13 + z);
            // exact-reduction unit: this thread's 8 elements of one column
            fixacc_add(fa0, pp0 + pp1, fq0, ovf);
            fixacc_add(fa1, pr0 + pr1, fq1, ovf);
            // rows k = 64 h + 8 g .. + 7 of column n: the 8-byte half (g & 1) of k-chunk 4 h + (g >> 1)
            const uint32_t off = sw128_chunk_off((uint32_t)(cp + 16 * z), 4u * h + (uint32_t)(g >> 1)) + 8u * (uint32_t)(g & 1);
#pragma unroll
            for (int sl = 0; sl < TC_SLICES; ++sl) {
              const int d = 6 - sl;                                   // digit index held by slice sl
              const uint32_t sel = (d & 3) | (((d & 3) + 4) << 4);    // byte d of a -> pos 0, byte d of b -> pos 1
              uint32_t wq[2];
#pragma unroll
              for (int q4 = 0; q4 < 2; ++q4) {
                const uint32_t x0 = d < 4 ? lo[4 * q4] : hi[4 * q4], x1 = d < 4 ? lo[4 * q4 + 1] : hi[4 * q4 + 1];
                const uint32_t x2 = d < 4 ? lo[4 * q4 + 2] : hi[4 * q4 + 2], x3 = d < 4 ? lo[4 * q4 + 3] : hi[4 * q4 + 3];
                const uint32_t t01 = __byte_perm(x0, x1, sel), t23 = __byte_perm(x2, x3, sel);
                wq[q4] = __byte_perm(t01, t23, 0x5410);
              }
              *reinterpret_cast<uint2 *>(Qsm + sl * TC_QTILE + off) = make_uint2(wq[0], wq[1]);
            }
          }
          mbar_arrive(&mb[B6_R_EMPTY + h]);                   // this half of the r tile may be refilled with the next block
          mbar_arrive(&mb[B6_P_FULL + h]);                    // p of this half is in the stage: M may take its fragments
        }
        fence_proxy_async_smem();
        if (t == 0) ms.s_E[u & 3] = E;
        mbar_arrive(&mb[B6_Q_FULL]);
        if (t == 0) { TL6(15); TLI(1); }
        // max |p| of this block (all 128 rows) for the next iteration's bound: one word per L warp, no barrier
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mxb = max(mxb, __shfl_xor_sync(0xffffffffu, mxb, o));
        // stored as the largest double with this high word (an upper bound of max |p|, at most 2^-20 above it)
        if (lane == 0) __stcg(Pnext + 4 * (size_t)b + (warp - 4), ((unsigned long long)(unsigned)mxb << 32) | 0xffffffffull);
        // a |p| above the bound cannot happen with finite data; non-finite data (Inf / NaN high words order above
        // every finite number) is flagged here and by the exact accumulators
        if (mxb > __double2hiint(bound)) ovf = 1u;
      }
      fixacc_flush(fa0, sacc + SC_PP * KUL_STRIDE, fe0);
      fixacc_flush(fa1, sacc + SC_PR * KUL_STRIDE, fe1);
      if (ovf) atomicOr((unsigned long long *)(set + ACC_FLAG_OFF), 1ull);
    } else if (MROLE) {
      // ===== M: p fragments, TMEM read-back, W = Z - p S, stores, partial sums =====
      // warp (qd, g16): TMEM lanes 32 qd + 16 g16 + [0, 16) = block rows 64 g16 + 16 qd + [0, 16), all 32 columns
      const int w = warp - 8, qd = w & 3, g16 = w >> 2;
      FixAcc fa0 = {0, 0}, fa1 = {0, 0};                     // <p,W>, <W,W>
      const int fe0 = ms.s_fe[SC_PHP], fe1 = ms.s_fe[SC_HPHP];
      const double fq0 = scalbn(1.0, 90 - fe0), fq1 = scalbn(1.0, 90 - fe1);
      unsigned ovf = 0;
      // Pass i = -1 is a DRY RUN of the block body (no waits, no arrivals, no stores to HBM, partial sums forced to zero)
      // while this role would otherwise idle waiting for the first block: it pulls the read-back / store code into the
      // instruction cache, which the other phases of the iteration have evicted (measured: the first read-back of a
      // phase took 4.3 us instead of 1.3 us).
      for (int i = -1; i < nb_local; ++i) {
        const bool dry = i < 0;
        const unsigned u = use + (dry ? 0 : i), b = bfirst + (dry ? 0 : i), r0 = b * ST_NB;
        const unsigned hh = 2u * b + (unsigned)g16;
        const bool own = dry || (hh >= h0 && hh < h1);
        if (tid == 256) TL6(21);
        // L has finished this warp's half of block u: p is in the stage (release / acquire through the P_FULL mbarrier;
        // L cannot be more than one block ahead of this role, so the parity is unambiguous)
        if (!dry) mbar_wait_guarded(&mb[B6_P_FULL + g16], u & 1);
        if (tid == 256) TL6(22);
        // this thread's elements of p in the accumulator arrangement: pc[s][nt] = p[row 8 s + m][8 nt + 2 j + {0, 1}]
        double2 pc[2][4];
        if (own) {
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const unsigned char *prow = POsm + g16 * V6_HALF + (16u * qd + 8u * s + m) * 256u + 16u * j;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) pc[s][nt] = *reinterpret_cast<const double2 *>(prow + 64u * nt);
          }
        }
        if (!dry) mbar_arrive(&mb[B6_PO_EMPTY + g16]);        // this half of the p tile may be refilled (p_old of the next block)
        if (tid == 256) { TL6(26); TLI(2); }
        if (own) {
          // The solve runs in the eigenbasis of S (ob200_stpcg rotates g, Y and s): W = A p - p Lambda, an elementwise
          // shift by the column's eigenvalue instead of a 128 x 32 x 32 fp64 product (the fp64 pipe is the scarce unit of
          // phase A: DESIGN.md 4.2b).  T = -lambda o p now, W = Z + T after the read-back.
          double acc[2][4][2];                                // [row tile s][n-tile][c]: row 8 s + m, columns 8 nt + 2 j + c
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const double2 nl = *reinterpret_cast<const double2 *>(Ssm + 8 * nt + 2 * j);   // -lambda of the two columns
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              acc[s][nt][0] = nl.x * pc[s][nt].x;
              acc[s][nt][1] = nl.y * pc[s][nt].y;
            }
          }
          if (tid == 256) TL6(27);
          // MMAs of block u complete: Z = A p from TMEM in the accumulator arrangement, W = Z + T
          if (!dry) mbar_wait_guarded(&mb[B6_MMA_DONE + (u & 1)], (u >> 1) & 1);
          if (tid == 256) { TL6(28); TLI(3); }
          tc_fence_after();
          const int E = ms.s_E[u & 3];
          const double sc = scalbn(1.0, __ldg(plane_exp + b) + E + 10);
          const uint32_t tacc = tmem_base + (u & 1) * TC_TMEM_COLS + ((uint32_t)(32 * qd + 16 * g16) << 16);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            double out[4];
            recombine_frag8(tacc + 8 * nt, out);
            acc[0][nt][0] = fma(out[0], sc, acc[0][nt][0]); acc[0][nt][1] = fma(out[1], sc, acc[0][nt][1]);
            acc[1][nt][0] = fma(out[2], sc, acc[1][nt][0]); acc[1][nt][1] = fma(out[3], sc, acc[1][nt][1]);
          }
          tc_fence_before();
          if (!dry) mbar_arrive(&mb[B6_TMEM_EMPTY + (u & 1)]);  // the accumulator set may be overwritten
          if (tid == 256) { TL6(23); TLI(4); }
          if (tid == 256) TL6(24);
          if (!dry && wcnt > 0) mbar_wait_guarded(&mb[B6_W_EMPTY + g16], (wcnt - 1) & 1);   // G is done with the previous occupant
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const unsigned row = 64u * g16 + 16u * qd + 8u * s + m, grow = r0 + row;
            const bool valid = !dry && grow < n_rows32;
            double pw = 0.0, ww = 0.0;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const int col = 8 * nt + 2 * j;
              const double px = pc[s][nt].x, py = pc[s][nt].y;
              const double wx = acc[s][nt][0], wy = acc[s][nt][1];
              pw = fma(px, wx, pw); pw = fma(py, wy, pw);
              ww = fma(wx, wx, ww); ww = fma(wy, wy, ww);
              const double2 wv = make_double2(wx, wy);
              // (dry run: stale operands -- the tile is rewritten by the first real block before G is signalled)
              *reinterpret_cast<double2 *>(Wsm + row * WS + col) = wv;
              if (valid) stcg2(a.Hp + (size_t)grow * ST_P + col, wv);
            }
            fixacc_add(fa0, dry ? 0.0 : pw, fq0, ovf);   // exact-reduction unit: this lane's 8 elements of the row
            fixacc_add(fa1, dry ? 0.0 : ww, fq1, ovf);
          }
          if (!dry) {
            mbar_arrive(&mb[B6_W_FULL + g16]);                // this thread's part of the half's W is staged
            wcnt += 1u;
          }
          if (tid == 256) { TL6(25); TLI(5); }
        } else {
          mbar_wait_guarded(&mb[B6_MMA_DONE + (u & 1)], (u >> 1) & 1);
          tc_fence_before();
          mbar_arrive(&mb[B6_TMEM_EMPTY + (u & 1)]);
        }
      }
      fixacc_flush(fa0, sacc + SC_PHP * KUL_STRIDE, fe0);
      fixacc_flush(fa1, sacc + SC_HPHP * KUL_STRIDE, fe1);
      if (ovf) atomicOr((unsigned long long *)(set + ACC_FLAG_OFF), 1ull);
    } else {
      // ===== G: projection Gram Y^T W, one exact unit per owned 64-row half =====
      const int gw = warp - 16, gt = tid - 512, mp = gw >> 1, np = gw & 1;   // warp gw: 16 x 16 quadrant (mp, np) of the Gram
      unsigned ovf = 0;
      i64 gfix[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) gfix[nt][0] = gfix[nt][1] = gfix[nt][2] = gfix[nt][3] = 0;
      // Y of a half -> its shared-memory area with cp.async (16-byte chunks, padded rows): issued one half ahead
      auto fetch_y = [&](unsigned hh) {
        const unsigned rbase = hh * 64u, hsel = hh & 1u;
        const int cpy = gt & 15, gy = gt >> 4;                // 8 rows x 16 column pairs per pass
#pragma unroll
        for (int ps = 0; ps < 8; ++ps) {
          const unsigned rloc = 8u * ps + gy, grow = rbase + rloc;
          double *dst = Ysm + (64u * hsel + rloc) * WS + 2 * cpy;
          if (grow < n_rows32) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)),
                         "l"(st.Y + (size_t)grow * ST_P + 2 * cpy) : "memory");
          } else {
            *reinterpret_cast<double2 *>(dst) = make_double2(0.0, 0.0);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      if (h1 > h0) fetch_y(h0);
      for (unsigned hh = h0; hh < h1; ++hh) {
        const unsigned hsel = hh & 1u;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        nbar_sync(NB_MSYNC2, 128);       // Y(hh) visible to all G warps; everybody is done with the other half's area
        if (hh + 1 < h1) fetch_y(hh + 1);
        if (tid == 512 && hh == h0 + 4) TL6B(30);
        mbar_wait_guarded(&mb[B6_W_FULL + hsel], (wcnt >> (16 * hsel)) & 1);
        if (tid == 512 && hh == h0 + 4) TL6B(31);
        { const int i = (int)((hh >> 1) - bfirst); if (tid == 512 && (hh & 1u) == 0) TLI(6); }
        const double *Yh = Ysm + 64u * hsel * WS, *Wh = Wsm + 64u * hsel * WS;
        // G[16 mp .. +15][16 np .. +15] over the 64 rows with small fp64 MMAs (m8n8k4): 2 x 2 tiles x two k-halves =
        // 8 independent chains of 8 k-steps, the k-halves added in a fixed order
        double ga[2][2][2][2];                                // [k-half][row tile][n-tile][c]
#pragma unroll
        for (int kh = 0; kh < 2; ++kh)
#pragma unroll
          for (int it = 0; it < 2; ++it)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) ga[kh][it][nt][0] = ga[kh][it][nt][1] = 0.0;
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            const int krow = 32 * kh + 4 * qq + j;
            const double y0 = Yh[krow * WS + 16 * mp + m], y1 = Yh[krow * WS + 16 * mp + 8 + m];
            const double w0 = Wh[krow * WS + 16 * np + m], w1 = Wh[krow * WS + 16 * np + 8 + m];
            dmma884(ga[kh][0][0][0], ga[kh][0][0][1], y0, w0);
            dmma884(ga[kh][0][1][0], ga[kh][0][1][1], y0, w1);
            dmma884(ga[kh][1][0][0], ga[kh][1][0][1], y1, w0);
            dmma884(ga[kh][1][1][0], ga[kh][1][1][1], y1, w1);
          }
        }
        mbar_arrive(&mb[B6_W_EMPTY + hsel]);                  // M may stage the next occupant of this half
        if (tid == 512 && hh == h0 + 4) TL6B(32);
        { const int i = (int)((hh >> 1) - bfirst); if (tid == 512 && (hh & 1u) == 1) TLI(7); }
        wcnt += 1u << (16 * hsel);
        // tile (it, nt): Gram row 16 mp + 8 it + m, columns 16 np + 8 nt + 2 j + {0, 1}
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int it = 0; it < 2; ++it)
            gram_accumulate(ga[0][it][nt][0] + ga[1][it][nt][0], ga[0][it][nt][1] + ga[1][it][nt][1], inv_q,
                            gfix[2 * nt + it], &ovf);
      }
      // gfix[2 nt + hrow] <-> 8 x 8 Gram tile (row tile 2 mp + hrow, column tile 2 np + nt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow)
          gram_flush(set, 4 * (2 * mp + hrow) + 2 * np + nt, lane, gfix[2 * nt + hrow], (nt | hrow) == 0 ? ovf : 0);
    }
    use += (unsigned)nb_local;
    if (tid == 128) TL6B(41);   // L done
    if (tid == 256) TL6B(42);   // M done
    if (tid == 512) TL6B(40);   // G done
    bar_work();
    flush_scalars_w(sacc, set, 4, rt);
    RedView rvw;
    if (!(MULTI ? grid_reduce_barrier_multi(ms, rt, a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, 0, ACC_WORDS,
                                            rvw, a.dbg ? ms.s_stamp : nullptr)
                : grid_reduce_barrier_w(ms, rt, a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, 0, ACC_WORDS, rvw,
                                        a.dbg ? ms.s_stamp : nullptr))) { exit_reason = -2; continue; }
    if (tid == 256) TL6B(43);     // barrier A released
    // first strip of phase B for this warp: start streaming it in before the scalar stage
    int cur = 0;
    if (lane == 0) cur = atomicSub(&ms.s_next_strip, 1);   // top-down: lines touched last in phase A first
    cur = __shfl_sync(0xffffffffu, cur, 0);
    if (cur >= s_lo && lane == 0) {
      fence_proxy_async_smem();
      fence_proxy_async_global_v6();
      strip_fetch6(slot, sb, cur, n_rows32, a.Hp, a.s, p_new, a.r, st.Y);
    }
    {
      const u64 flag = rvw.load(ACC_FLAG_OFF);
      double c = 0.0;
      {
        // G (fixed point) -> shared memory (M and G warps) while the four L warps finalize the four exact scalars
        double *Graw = reinterpret_cast<double *>(base + V6_GRAW);
        if (role >= 2) {
          for (int e = rt - 128; e < ST_P * ST_P; e += 384) {
            u64 hi, lo;
            if (rvw.world == 1) {
              const ulonglong2 wv = __ldcg(reinterpret_cast<const ulonglong2 *>(rvw.base0 + ACC_GRAM_OFF + 2 * e));
              hi = wv.x; lo = wv.y;
            } else {
              hi = 0; lo = 0;
              for (int r = 0; r < rvw.world; ++r) {           // one 16-byte load per source rank
                const ulonglong2 wv = __ldcg(reinterpret_cast<const ulonglong2 *>(rvw.base0 + (size_t)r * rvw.stride + ACC_GRAM_OFF + 2 * e));
                hi += wv.x; lo += wv.y;
              }
            }
            Graw[e] = fix2_to_double((i64)hi, (i64)lo, q);
          }
        } else {
          const int o = rw * KUL_STRIDE;                      // L warp w finalizes scalar w
          const double x = kul_finalize_warp([&rvw, o](int jj) { return rvw.load(o + jj); });
          if (lane == 0) sh.red[rw] = x;
        }
        bar_work();
        for (int e = rt; e < ST_P * ST_P; e += V6_WORK) {
          const int i = e >> 5, jj = e & 31;
          const double sg = 0.5 * (Graw[e] + Graw[jj * ST_P + i]);
          Gsm[i * GS + jj] = -sg;
          c = fma(sg, sg, c);
        }
      }
      c = warp_sum(c);
      if (lane == 0) ms.s_part[rw] = c;
      bar_work();
      if (flag != 0) {
        if (cur >= s_lo) mbar_wait_guarded(sb, bpar);        // drain the outstanding fetch before leaving
        exit_reason = -3;
        continue;
      }
      if (role == 1) {
        if (rw == 0) {
          // lanes 0..2 evaluate the long-latency operations concurrently, lane 0 takes the decisions
          double nG2 = 0.0;
#pragma unroll
          for (int ww = 0; ww < 16; ++ww) nG2 += ms.s_part[ww];
          const double nHp2 = fmax(sh.red[SC_HPHP] - nG2, 0.0);
          double slow = 0.0;
          if (lane == 0) slow = sqrt(nHp2);
          else if (lane == 1) slow = sqrt(sh.red[SC_PP]);
          else if (lane == 2) slow = __ddiv_rn(sh.rv, sh.red[SC_PHP]);                    // alpha, l.341
          const double sq_nHp2 = __shfl_sync(0xffffffffu, slow, 0), sq_np2 = __shfl_sync(0xffffffffu, slow, 1);
          const double alpha = __shfl_sync(0xffffffffu, slow, 2);
          if (lane == 0) {
            decide_after_A_pre(sh, sh.red[SC_PHP], sq_nHp2, sq_np2, alpha, sh.red[SC_PR], a.Delta, a.epsilon);
            // ||r + alpha Hp||^2 <= 2 (||r||^2 + alpha^2 ||Hp||^2)
            ms.s_fe[SC_RV] = fixacc_exponent(2.0 * (sh.rv + sh.step * sh.step * sh.red[SC_HPHP]));
          }
        }
      }
      bar_work();
    }
    ++phase;
    if (tid == 256) TL6B(44);     // scalar stage done
    const double step = sh.step;
    if (sh.action != ACT_CONTINUE) {
      if (cur >= s_lo) mbar_wait_guarded(sb, bpar);
      const size_t e0 = (size_t)row_lo * ST_P, e1 = (size_t)row_hi * ST_P;
      for (size_t e = e0 + 2 * (size_t)rt; e < e1; e += 2 * (size_t)V6_WORK) {
        double2 sv = ldcg2(a.s + e);
        const double2 pv = ldcg2(p_new + e);
        sv.x = fma(step, pv.x, sv.x);
        sv.y = fma(step, pv.y, sv.y);
        stcg2(a.s + e, sv);
      }
      exit_reason = sh.action - 1;
      continue;
    }

    // ------------------------------ phase B ------------------------------
    set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    {
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      const int per = (ACC_WORDS + gridDim.x - 1) / gridDim.x;
      const int z0 = per * blockIdx.x;
      for (int i = rt; i < per && z0 + i < ACC_WORDS; i += V6_WORK) nxt[z0 + i] = 0;
    }
    {
      unsigned ovfb = 0;
      FixAcc fb = {0, 0};
      const int feb = ms.s_fe[SC_RV];
      const double fqb = scalbn(1.0, 90 - feb);
      unsigned long long *Rmax = a.blk_stats + ((k + 1ull) & 1ull) * a.nblk_stats;
      while (cur >= s_lo) {
        const int sidx = cur;
        int nxt = 0;
        if (lane == 0) nxt = atomicSub(&ms.s_next_strip, 1);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        const unsigned grow = (unsigned)sidx * 8u + m;
        const bool valid = grow < n_rows32;
        const size_t rowoff = (size_t)grow * ST_P;
        mbar_wait_guarded(sb, bpar);
        bpar ^= 1;
        const double *tW = reinterpret_cast<const double *>(slot) + m * ST_P;
        const double *tS = tW + 8 * ST_P, *tP = tS + 8 * ST_P, *tR = tP + 8 * ST_P, *tY = tR + 8 * ST_P;
        unsigned tok = 0;
        // s += alpha p first (l.374): its operands leave the registers before the fp64 MMAs need theirs
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int col = 8 * t + 2 * j;
          if (valid) {
            double2 sv = *reinterpret_cast<const double2 *>(tS + col);
            const double2 pv = *reinterpret_cast<const double2 *>(tP + col);
            sv.x = fma(step, pv.x, sv.x);  sv.y = fma(step, pv.y, sv.y);
            stcg2(a.s + rowoff + col, sv);
            tok |= __double2hiint(sv.x);
          }
        }
        double acc[4][2];
        double2 rv[4], yx[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int col = 8 * t + 2 * j;
          if (valid) {
            const double2 wv = *reinterpret_cast<const double2 *>(tW + col);
            acc[t][0] = wv.x; acc[t][1] = wv.y;
            rv[t] = *reinterpret_cast<const double2 *>(tR + col);
            yx[t] = *reinterpret_cast<const double2 *>(tY + 8 * j + 2 * t);
          } else {
            acc[t][0] = acc[t][1] = 0.0;
            rv[t] = yx[t] = make_double2(0.0, 0.0);
          }
          tok |= __double2hiint(acc[t][0]) | __double2hiint(rv[t].x) | __double2hiint(yx[t].x);
        }
        // every lane's shared-memory reads have returned (tok depends on all of them): the slot may be refilled
        tok = __reduce_or_sync(0xffffffffu, tok);
        if (nxt >= s_lo && lane == 0 && (tok | 1u)) {
          fence_proxy_async_smem();
          strip_fetch6(slot, sb, nxt, n_rows32, a.Hp, a.s, p_new, a.r, st.Y);
        }
        strip_rightmul_v(yx, Gsm, lane, acc);   // Hp = W - Y symG
        double rr = 0.0;
        int rmh = 0;                                      // max |r| of the strip through the (ordered) high words
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int col = 8 * t + 2 * j;
          rv[t].x = fma(step, acc[t][0], rv[t].x); rv[t].y = fma(step, acc[t][1], rv[t].y);   // l.377
          rr = fma(rv[t].x, rv[t].x, rr); rr = fma(rv[t].y, rv[t].y, rr);                      // l.383,408
          rmh = max(rmh, max(__double2hiint(rv[t].x) & 0x7fffffff, __double2hiint(rv[t].y) & 0x7fffffff));
          if (valid) stcg2(a.r + rowoff + col, rv[t]);
        }
        fixacc_add(fb, rr, fqb, ovfb);      // exact-reduction unit: this lane's 8 elements of the strip
        // max |r| per 128-row block for the scale bound of the next phase A (order independent)
        rmh = __reduce_max_sync(0xffffffffu, rmh);
        // (upper bound of the maximum: the largest double with this high word)
        if (lane == 0) atomicMax(Rmax + ((unsigned)sidx >> 4), ((unsigned long long)(unsigned)rmh << 32) | 0xffffffffull);
        cur = nxt;
      }
      if (tid == 256) TL6B(45);   // phase B strips done (warp 8)
      fixacc_flush(fb, sacc + SC_RV * KUL_STRIDE, feb);
      if (ovfb) atomicAdd(sacc + SC_RV * KUL_STRIDE + KUL_LIMBS, 1ull);   // non-finite / bound violated: poison <r,r>
    }
    bar_work();
    if (tid == 256) TL6B(46);     // phase B done, CTA-wide
    if (rt == 0) ms.s_next_strip = s_hi - 1;
    flush_scalars_w(sacc + SC_RV * KUL_STRIDE, set + SC_RV * KUL_STRIDE, 1, rt);
    if (!(MULTI ? grid_reduce_barrier_multi(ms, rt, a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set,
                                            SC_RV * KUL_STRIDE, KUL_STRIDE, rvw, a.dbg ? ms.s_stamp + 2 : nullptr)
                : grid_reduce_barrier_w(ms, rt, a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, SC_RV * KUL_STRIDE,
                                        KUL_STRIDE, rvw, a.dbg ? ms.s_stamp + 2 : nullptr))) { exit_reason = -2; continue; }
    if (role == 1) {
      if (rw == 0) {
        const int o = SC_RV * KUL_STRIDE;
        const double x = kul_finalize_warp([&rvw, o](int jj) { return rvw.load(o + jj); });
        if (lane == 0) {
          update_after_B(sh, x);
          // bounds for the next iteration's exact accumulators (integer exponent arithmetic only)
          const int e = half_exponent(st.op_norm_bound * st.op_norm_bound * sh.pk_M_2 * 16.0) + 2;   // |G_ij| <= ||H|| ||p||
          ms.s_invq = scalbn(1.0, 90 - e);
          ms.s_q = scalbn(1.0, e - 90);
          ms.s_fe[SC_PHP] = fixacc_exponent(st.op_norm_bound * sh.pk_M_2);
          ms.s_fe[SC_HPHP] = fixacc_exponent(st.op_norm_bound * st.op_norm_bound * sh.pk_M_2);
          ms.s_fe[SC_PP] = fixacc_exponent(sh.pk_M_2);
          ms.s_fe[SC_PR] = half_exponent(sh.pk_M_2 * sh.rv) + 2;                                      // |<p,r>| <= ||p|| ||r||
        }
      }
    }
    bar_work();
    ++phase;
    if (tid == 256) TL6B(47);     // iteration done
    if (a.dbg && rt == 0) {   // [work A, wait A, work B, wait B]; work = previous release -> arrival
      if (dbg_prev) atomicAdd(a.dbg + 4 * blockIdx.x + 0, ms.s_stamp[0] - dbg_prev);
      atomicAdd(a.dbg + 4 * blockIdx.x + 1, ms.s_stamp[1] - ms.s_stamp[0]);
      atomicAdd(a.dbg + 4 * blockIdx.x + 2, ms.s_stamp[2] - ms.s_stamp[1]);
      atomicAdd(a.dbg + 4 * blockIdx.x + 3, ms.s_stamp[3] - ms.s_stamp[2]);
      dbg_prev = ms.s_stamp[3];
    }
  }

  if (blockIdx.x == 0 && rt == 0) {
    TcgDeviceResult *res = a.result;
    res->num_iterations = sh.k;
    res->final_rv = sh.rv;
    res->phases = phase;
    if (exit_reason < 0) {
      res->status = (exit_reason == -3) ? 4 /*OB200_NUMERIC_RANGE*/ : 5 /*OB200_ABORTED*/;
      res->exit_reason = -1;
      res->update_step_M_norm = 0.0;
    } else {
      res->status = 0;
      res->exit_reason = exit_reason;
      res->update_step_M_norm = (exit_reason >= 2) ? a.Delta : sqrt(sh.sk_M_2);
    }
  }
}

template <bool MULTI>
__global__ void __launch_bounds__(V6_THREADS, 1)
tcg_stiefel_v6_kernel(TcgCommon a, StiefelArgs st, const unsigned char *planes, const int *plane_exp) {
  // the dynamic shared-memory window is declared 1024-byte aligned (SWIZZLE_128B operand images); all pointers are
  // derived from the array itself so that the compiler keeps them in the shared address space (LDS / STS, not generic)
  unsigned char *base = v6_smem_raw;
  if (smem_u32(base) & 1023u) __trap();
  V6Misc &ms = *reinterpret_cast<V6Misc *>(base + V6_MISC);
  u64 *sacc = reinterpret_cast<u64 *>(base + V6_ACC);
  double *Ssm = reinterpret_cast<double *>(base + V6_S);
  uint64_t *mb = reinterpret_cast<uint64_t *>(base + V6_BAR);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (int)(V6_NACC * KUL_STRIDE); i += blockDim.x) sacc[i] = 0;
  if (tid < ST_P) Ssm[tid] = -st.S[tid];   // st.S: the 32 eigenvalues of S (the solve runs in its eigenbasis)
  if (tid == 0) {
    CgShared &sh = ms.sh;
    sh.rv = a.rv0;
    sh.sk_M_pk = 0.0;
    sh.sk_M_2 = 0.0;
    sh.pk_M_2 = a.rv0;
    sh.alpha = sh.beta = sh.kappa = sh.step = 0.0;
    sh.k = 0;
    sh.action = ACT_CONTINUE;
    sh.status = 0;
    ms.done = 0;
    const unsigned nstrips = ((unsigned)st.n_rows + 7u) >> 3;
    ms.s_next_strip = (int)((unsigned long long)nstrips * (blockIdx.x + 1ull) / gridDim.x) - 1;
    const int e = gram_exponent(st.op_norm_bound * sqrt(a.rv0) * 4.0);
    ms.s_invq = scalbn(1.0, 90 - e);
    ms.s_q = scalbn(1.0, e - 90);
    ms.s_fe[SC_PHP] = fixacc_exponent(st.op_norm_bound * a.rv0);                         // |<p,W>| <= ||H|| ||p||^2
    ms.s_fe[SC_HPHP] = fixacc_exponent(st.op_norm_bound * st.op_norm_bound * a.rv0);
    ms.s_fe[SC_PP] = fixacc_exponent(a.rv0);                                             // ||p||^2 = pk_M_2 (l.266)
    ms.s_fe[SC_PR] = fixacc_exponent(a.rv0);                                             // |<p,r>| <= ||p|| ||r||
    ms.s_fe[SC_RV] = 0;
    for (int h = 0; h < 2; ++h) {
      mbar_init(&mb[B6_RP_FULL + h], 1);
      mbar_init(&mb[B6_R_EMPTY + h], 128);
      mbar_init(&mb[B6_PO_EMPTY + h], 128);
      mbar_init(&mb[B6_P_FULL + h], 128);
    }
    mbar_init(&mb[B6_A_FULL], 1);
    mbar_init(&mb[B6_Q_FULL], 128);
    mbar_init(&mb[B6_MMA_DONE], 1);
    mbar_init(&mb[B6_MMA_DONE + 1], 1);
    mbar_init(&mb[B6_TMEM_EMPTY], 256);
    mbar_init(&mb[B6_TMEM_EMPTY + 1], 256);
    for (int s = 0; s < (int)V6_NSLOT; ++s) mbar_init(&mb[B6_SLOT + s], 1);
    mbar_init(&mb[B6_W_FULL], 128);
    mbar_init(&mb[B6_W_FULL + 1], 128);
    mbar_init(&mb[B6_W_EMPTY], 128);
    mbar_init(&mb[B6_W_EMPTY + 1], 128);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&ms.s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // register budgets per warp group: 640 threads are launched with 96 registers each (61440 in the CTA's pool)
  //   S 32 x 128 + L 104 x 128 + M 120 x 256 + G 104 x 128 = 61440
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    v6_run_service(a, st, planes, base);
  } else if (warp >= 8 && warp < 16) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    v6_run<true, MULTI>(2, a, st, plane_exp, base);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    v6_run<false, MULTI>(warp < 8 ? 1 : 3, a, st, plane_exp, base);
  }
  tc_fence_before();
  bar_cta();
  if (warp == 0) tmem_dealloc(ms.s_tmem, 512);
}

cudaError_t launch_tcg_stiefel_v6(const TcgCommon &a, unsigned long long n_rows, const unsigned short *A,
                                  const double *Y, const double *S_dev, double op_norm_bound,
                                  const unsigned char *planes, const int *plane_exp, int sm_count, cudaStream_t stm) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tcg_stiefel_v6_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6_TOTAL);
    if (e) return e;
    e = cudaFuncSetAttribute(tcg_stiefel_v6_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6_TOTAL);
    if (e) return e;
    attr = true;
  }
  const unsigned long long nhalf = (n_rows + 63ull) / 64ull;
  int grid = sm_count;
  if ((unsigned long long)grid > nhalf) grid = (int)nhalf;
  TcgCommon ac = a;
  StiefelArgs sa{n_rows, A, Y, S_dev, op_norm_bound};
  const unsigned char *pl = planes;
  const int *pe = plane_exp;
  void *args[] = {(void *)&ac, (void *)&sa, (void *)&pl, (void *)&pe};
  const void *fn = a.cm.world > 1 ? (const void *)tcg_stiefel_v6_kernel<true> : (const void *)tcg_stiefel_v6_kernel<false>;
  return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(V6_THREADS), args, V6_TOTAL, stm);
}

}  // namespace ob200
