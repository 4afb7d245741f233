// Persistent fused truncated-CG for the Stiefel trace-minimisation Hessian, v5: warp-specialised
// roles with their own register budgets (setmaxnreg), every bulk operand of phase A staged by the
// TMA unit, the TMEM read-back done by the warps that consume it.  Same mathematics, reductions and
// scalar logic as tcg_stiefel_tc_kernel (v4, tcg_stiefel_tc.cu; reference line map in tcg.cuh):
// each CG iteration of IterativeSolvers.h:285-422 is two fused phases separated by an exact grid-wide
// (machine-wide) reduction.
//
// 20 warps = 5 warp groups:
//   S (warps 0-3, 24 registers)  : warp 0 lane 0 = TMA producer: r / p_old tiles of the next block into a 64 KB stage,
//                                  the block's three int8 digit planes of A (48 KB), L2 prefetch of what follows;
//                                  warp 1 lane 0 = MMA issuer: 13 tcgen05.mma kind::i8 per block into one of two
//                                  256-column TMEM accumulator sets, tcgen05.commit -> mbarrier.
//   L (warps 4-7, 80 registers)  : r, p_old from the stage -> p = -r + beta p (l.420; written back for the rows this CTA
//                                  owns), <p,p>, <p,r>, block maximum; second pass over the stage: seven balanced int8
//                                  digit slices of p straight into the UMMA K-major SWIZZLE_128B operand image.
//   M (warps 8-15, 144 registers): TMEM -> registers in the mma.sync accumulator arrangement (tcgen05.ld 16x256b: no
//                                  shared-memory round trip), integer recombination = Z = A p; W = Z - p S on the fp64
//                                  tensor cores, W written back and staged per 64-row half, <p,W>, <W,W>.
//   G (warps 16-19, 88 registers): projection Gram Y^T W of the staged half (fp64 tensor cores, exact fixed-point
//                                  accumulation), Y of the next half fetched with cp.async meanwhile.
// The fp64 MMA has a dependent-issue latency of ~210 cycles on this part (tools/probe/dmma16.cu): M and G are latency
// chains, so they run as separate, pipelined roles.  Hand-offs through mbarriers only.  Ownership is by 64-row HALF blocks (balanced to 1/11 instead of 1/6 of a CTA's
// work): a block shared by two CTAs is sliced and multiplied by both (the MMA needs all 128 rows of p as K), everything
// else -- p / W stores, the fp64 MMAs, the Gram, all partial sums -- is done for the owned half only.  The A images are
// row-permuted (tc_row_of_lane) so that either half occupies 16 lanes of every TMEM lane quarter: all eight M warps
// work on a half block.
// Phase B (l.374-408): the M warps, two 10 KB strip slots each (16 TMA-fed slots per SM as in v4).
#include "tcg.cuh"
#include "stiefel_dev.cuh"
#include "tc_common.cuh"

namespace ob200 {
using namespace tc;

constexpr int V5_THREADS = 640;
// shared-memory map (bytes from the 1024-aligned base); phase B aliases the phase-A operand space
constexpr uint32_t V5_A = 0;                                   // 48 KB int8 digit planes of A
constexpr uint32_t V5_Q = V5_A + TC_ABLOCK;                    // 28 KB int8 digit image of p
constexpr uint32_t V5_TILE = ST_NB * ST_P * 8;                 // 32 KB dense fp64 tile
constexpr uint32_t V5_R = V5_Q + TC_QBYTES;                    // r tile (TMA)
constexpr uint32_t V5_PO = V5_R + V5_TILE;                     // p_old tile (TMA)
constexpr uint32_t V5_WB = ST_NB * WS * 8;                     // 36 KB padded tile
constexpr uint32_t V5_W = V5_PO + V5_TILE;                     // W tile (stride WS)
constexpr uint32_t V5_Y = V5_W + V5_WB;                        // Y tile (stride WS)
constexpr uint32_t V5_S = V5_Y + V5_WB;                        // -S (stride WS)
constexpr uint32_t V5_ACC = V5_S + ST_P * WS * 8;              // CTA Kulisch accumulators (5 scalars)
constexpr uint32_t V5_NACC = 5;
constexpr uint32_t V5_BAR = V5_ACC + V5_NACC * KUL_STRIDE * 8; // mbarriers
constexpr uint32_t V5_NBAR = 32;
constexpr uint32_t V5_MISC = V5_BAR + V5_NBAR * 8;
constexpr uint32_t V5_MISC_BYTES = 768;
constexpr uint32_t V5_TOTAL = V5_MISC + V5_MISC_BYTES;         // (+ up to 1 KB of alignment padding after the static part)
// phase B
constexpr uint32_t V5_STRIP_TILE = 8 * ST_P * 8;               // 2 KB
constexpr uint32_t V5_SLOT = 5 * V5_STRIP_TILE;                // 10 KB: W, s, p, r, Y tiles of one 8-row strip
constexpr uint32_t V5_NSLOT = 16;
constexpr uint32_t V5_GRAW = V5_NSLOT * V5_SLOT;               // 8 KB scratch (inside the W tile region)
constexpr uint32_t V5_G = V5_Y;                                // -sym(G), stride GS (inside the Y tile region)
static_assert(V5_GRAW >= V5_W && V5_GRAW + 8192 <= V5_Y, "G scratch must sit in the W tile region");
static_assert(ST_P * GS * 8 <= V5_WB, "G must fit the Y tile region");
static_assert(V5_TOTAL + 1024 <= 232448, "shared memory budget (227 KB per CTA)");

enum { B5_RP_FULL = 0, B5_RP_EMPTY = 1, B5_A_FULL = 2, B5_Q_FULL = 3, B5_MMA_DONE = 4 /*,5*/, B5_TMEM_EMPTY = 6 /*,7*/,
       B5_SLOT = 8 /* .. 23 */, B5_W_FULL = 24 /*,25*/, B5_W_EMPTY = 26 /*,27*/ };

struct V5Misc {
  CgShared sh;
  double s_part[20];
  double s_invq, s_q;
  double s_lmax[8];
  int s_fe[5];
  int s_E[4];
  int s_next_strip;
  uint32_t s_tmem;
  unsigned long long s_stamp[4];
};
static_assert(sizeof(V5Misc) <= V5_MISC_BYTES, "misc block");

__device__ __forceinline__ void bulk_prefetch_l2_v5(const void *g, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 0;" ::: "memory"); }

// Digit slicing for the L role of v5 (128 threads): thread (cp, g) holds P[16 g + i][2 cp + z], i < 16, i.e. one whole
// 16-byte k-chunk (k = 16 g .. 16 g + 15) per (slice, n): seven 16-byte stores per column.
__device__ __forceinline__ void slice_tile16_to_smem(const double (&p)[16][2], double scale, unsigned char *Qsm, int cp,
                                                     int g) {
#pragma unroll
  for (int z = 0; z < 2; ++z) {
    const uint32_t n = 2 * cp + z;
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const unsigned long long u =
          ((unsigned long long)__double2ll_rn(p[i][z] * scale) + TC_DIGIT_BIAS) ^ TC_DIGIT_BIAS;
      lo[i] = (uint32_t)u;
      hi[i] = (uint32_t)(u >> 32);
    }
    const uint32_t off = sw128_chunk_off(n, (uint32_t)g);
#pragma unroll
    for (int s = 0; s < TC_SLICES; ++s) {
      const int d = 6 - s;                                    // digit index held by slice s
      const uint32_t sel = (d & 3) | (((d & 3) + 4) << 4);    // byte d of a -> pos 0, byte d of b -> pos 1
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t x0 = d < 4 ? lo[4 * q] : hi[4 * q], x1 = d < 4 ? lo[4 * q + 1] : hi[4 * q + 1];
        const uint32_t x2 = d < 4 ? lo[4 * q + 2] : hi[4 * q + 2], x3 = d < 4 ? lo[4 * q + 3] : hi[4 * q + 3];
        const uint32_t t01 = __byte_perm(x0, x1, sel), t23 = __byte_perm(x2, x3, sel);
        w[q] = __byte_perm(t01, t23, 0x5410);
      }
      *reinterpret_cast<uint4 *>(Qsm + s * TC_QTILE + off) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__device__ __forceinline__ void strip_fetch5(unsigned char *slot, uint64_t *bar, int sidx, unsigned n_rows,
                                             const double *W, const double *S, const double *Pn, const double *R,
                                             const double *Y) {
  const unsigned row0 = (unsigned)sidx * 8u;
  const unsigned rows = n_rows - row0 < 8u ? n_rows - row0 : 8u;
  const uint32_t bytes = rows * ST_P * (uint32_t)sizeof(double);
  const size_t off = (size_t)row0 * ST_P;
  mbar_expect_tx(bar, 5 * bytes);
  bulk_g2s(slot, W + off, bytes, bar);
  bulk_g2s(slot + V5_STRIP_TILE, S + off, bytes, bar);
  bulk_g2s(slot + 2 * V5_STRIP_TILE, Pn + off, bytes, bar);
  bulk_g2s(slot + 3 * V5_STRIP_TILE, R + off, bytes, bar);
  bulk_g2s(slot + 4 * V5_STRIP_TILE, Y + off, bytes, bar);
}

extern __shared__ __align__(1024) unsigned char v5_smem_raw[];

// debug timeline (CTA 0, third block of an iteration, one lane per role): slot <- globaltimer
#ifdef OB200_TIMELINE_BUILD
#define TL5(slot) do { if (a.dbg && blockIdx.x == 0 && i == 2) a.dbg[4096 + (slot)] = globaltimer_ns(); } while (0)
#define TL5B(slot) do { if (a.dbg && blockIdx.x == 0) a.dbg[4096 + (slot)] = globaltimer_ns(); } while (0)
#else
#define TL5(slot) do { } while (0)
#define TL5B(slot) do { } while (0)
#endif

// ROLE: 0 = S (service: TMA producer + MMA issuer), 1 = L, 2 = M, 3 = G.  The whole CG loop is instantiated per role so
// that each warp group's code is compiled against its own register budget.
template <int ROLE>
__device__ __forceinline__ void v5_run(const TcgCommon &a, const StiefelArgs &st, const unsigned char *planes,
                                       const int *plane_exp, unsigned char *base) {
  unsigned char *Asm = base + V5_A;
  unsigned char *Qsm = base + V5_Q;
  const unsigned char *Rsm = base + V5_R;
  const unsigned char *POsm = base + V5_PO;
  double *Wsm = reinterpret_cast<double *>(base + V5_W);
  double *Ysm = reinterpret_cast<double *>(base + V5_Y);
  const double *Ssm = reinterpret_cast<const double *>(base + V5_S);
  double *Gsm = reinterpret_cast<double *>(base + V5_G);
  u64 *sacc = reinterpret_cast<u64 *>(base + V5_ACC);
  uint64_t *mb = reinterpret_cast<uint64_t *>(base + V5_BAR);
  V5Misc &ms = *reinterpret_cast<V5Misc *>(base + V5_MISC);
  CgShared &sh = ms.sh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = lane >> 2, j = lane & 3;
  const uint32_t tmem_base = ms.s_tmem;

  // ownership in phase A: 64-row half blocks [h0, h1)
  const unsigned n_rows32 = (unsigned)st.n_rows;
  const unsigned nhalf = (n_rows32 + 63u) >> 6;
  const unsigned h0 = (unsigned)((unsigned long long)nhalf * blockIdx.x / gridDim.x);
  const unsigned h1 = (unsigned)((unsigned long long)nhalf * (blockIdx.x + 1ull) / gridDim.x);
  const unsigned bfirst = h0 >> 1;
  const int nb_local = (h1 > h0) ? (int)(((h1 - 1u) >> 1) - bfirst + 1u) : 0;
  const unsigned row_lo = h0 * 64u;
  const unsigned row_hi = (h1 * 64u < n_rows32) ? h1 * 64u : n_rows32;
  // phase B: 8-row strips split evenly over the CTAs, handed to the M warps dynamically, top-down
  const unsigned nstrips = (n_rows32 + 7u) >> 3;
  const int s_lo = (int)((unsigned long long)nstrips * blockIdx.x / gridDim.x);
  const int s_hi = (int)((unsigned long long)nstrips * (blockIdx.x + 1ull) / gridDim.x);
  unsigned bpar0 = 0, bpar1 = 0;   // M: parities of this warp's two strip-slot mbarriers
  unsigned gen = 0, phase = 0;
  unsigned use = 0;                // blocks processed so far by this CTA (mbarrier phase bookkeeping)
  unsigned wuse = 0;               // M / G: fill / drain count of the two half-block W staging areas (16 bits each)
  int exit_reason = -1;
  unsigned long long dbg_prev = 0;

  for (;;) {
    const unsigned long long k = sh.k;
    if (k >= a.max_iterations) { exit_reason = 1; break; }
    if (sqrt(sh.rv) <= a.target) { exit_reason = 0; break; }
    const double beta = sh.beta;
    const double *p_old = (k & 1ull) ? a.p1 : a.p0;
    double *p_new = (k & 1ull) ? a.p0 : a.p1;
    const double inv_q = ms.s_invq, q = ms.s_q;

    // ------------------------------ phase A ------------------------------
    u64 *set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    {   // recycle the set used two phases from now (every CTA clears its slice; a grid barrier intervenes)
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      const int per = (ACC_WORDS + gridDim.x - 1) / gridDim.x;
      const int z0 = per * blockIdx.x;
      for (int i = tid; i < per && z0 + i < ACC_WORDS; i += blockDim.x) nxt[z0 + i] = 0;
    }
    if constexpr (ROLE == 0) {
      // ===== S: TMA producer (warp 0) and MMA issuer (warp 1), one elected lane each =====
      if (warp == 0 && lane == 0) {
        fence_proxy_async_global();   // r / p written with generic stores by other CTAs (ordered by the grid barrier)
        for (int i = 0; i < nb_local; ++i) {
          const unsigned u = use + i, b = bfirst + i, r0 = b * ST_NB;
          const unsigned rows = n_rows32 - r0 < ST_NB ? n_rows32 - r0 : ST_NB;
          const uint32_t bytes = rows * ST_P * (uint32_t)sizeof(double);
          const size_t off = (size_t)r0 * ST_P;
          TL5(0);
          if (u > 0) mbar_wait_guarded(&mb[B5_RP_EMPTY], (u - 1) & 1);          // L has taken the previous block out of the stage
          TL5(1);
          mbar_expect_tx(&mb[B5_RP_FULL], k ? 2 * bytes : bytes);
          bulk_g2s(base + V5_R, a.r + off, bytes, &mb[B5_RP_FULL]);
          if (k) bulk_g2s(base + V5_PO, p_old + off, bytes, &mb[B5_RP_FULL]);
          if (i + 1 < nb_local) {   // what the next block needs that is not fetched early: A planes, Y; then r / p_old one further
            const unsigned rn = r0 + ST_NB;
            const unsigned rows1 = n_rows32 - rn < ST_NB ? n_rows32 - rn : ST_NB;
            const uint32_t bytes1 = rows1 * ST_P * (uint32_t)sizeof(double);
            bulk_prefetch_l2_v5(planes + (size_t)(b + 1) * TC_ABLOCK, TC_ABLOCK);
            bulk_prefetch_l2_v5(st.Y + (size_t)rn * ST_P, bytes1);
            if (i + 2 < nb_local) {
              const unsigned r2 = rn + ST_NB;
              const unsigned rows2 = n_rows32 - r2 < ST_NB ? n_rows32 - r2 : ST_NB;
              const uint32_t bytes2 = rows2 * ST_P * (uint32_t)sizeof(double);
              bulk_prefetch_l2_v5(a.r + (size_t)r2 * ST_P, bytes2);
              if (k) bulk_prefetch_l2_v5(p_old + (size_t)r2 * ST_P, bytes2);
            }
          }
          TL5(2);
          if (u > 0) mbar_wait_guarded(&mb[B5_MMA_DONE + ((u - 1) & 1)], ((u - 1) >> 1) & 1);   // A image free
          mbar_expect_tx(&mb[B5_A_FULL], TC_ABLOCK);
          bulk_g2s(Asm, planes + (size_t)b * TC_ABLOCK, TC_ABLOCK, &mb[B5_A_FULL]);
          TL5(3);
        }
      } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < nb_local; ++i) {
          const unsigned u = use + i;
          TL5(4);
          mbar_wait_guarded(&mb[B5_Q_FULL], u & 1);
          TL5(5);
          mbar_wait_guarded(&mb[B5_A_FULL], u & 1);
          TL5(6);
          if (u >= 2) mbar_wait_guarded(&mb[B5_TMEM_EMPTY + (u & 1)], ((u >> 1) - 1) & 1);   // M has drained this accumulator set
          TL5(7);
          tc_fence_after();
          issue_block_mmas(smem_u32(Asm), smem_u32(Qsm), tmem_base + (u & 1) * TC_TMEM_COLS);
          umma_commit(&mb[B5_MMA_DONE + (u & 1)]);
          TL5(8);
        }
      }
      __syncwarp();
    } else if constexpr (ROLE == 1) {
      // ===== L: p = -r + beta p_old, digit slices -- ONE pass over the staged block =====
      // The digit slices need a scale 2^E with |p| < 2^E over the block BEFORE the first element is cut.  Instead of a
      // maximum pass and a second pass, E comes from a bound that is known when the block arrives:
      //   max |p_new| <= max |r| + |beta| max |p_old|      (both maxima per 128-row block, exact, kept in blk_stats:
      //   max |r| from phase B of the previous iteration / the init kernel, max |p_old| from this role one iteration ago)
      // It is a deterministic function of exactly reduced data (identical on every CTA / GPU that handles the block) and
      // at most a few bits above the true maximum: p is quantised to 2^(E-54), i.e. no coarser than ~2^-52 of the
      // block maximum.
      const int t = tid - 128, cp = t & 15, g = t >> 4;      // columns 2cp, 2cp+1 ; rows 16g .. 16g+15 of the block
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
      for (int i = 0; i < nb_local; ++i) {
        const unsigned u = use + i, b = bfirst + i, r0 = b * ST_NB;
        const unsigned hh = 2u * b + (unsigned)(g >> 2);     // this thread's half block
        const bool mine = hh >= h0 && hh < h1;
        const unsigned char *rrow = Rsm + (16u * g) * 256u + 16u * cp, *prow = POsm + (16u * g) * 256u + 16u * cp;
        // scale from the bound (uniform over the CTA: every thread reads the same five words)
        const double rmax = __longlong_as_double((long long)__ldcg(Rcur + b));
        const ulonglong2 pm01 = __ldcg(reinterpret_cast<const ulonglong2 *>(Pcur + 4 * (size_t)b));
        const ulonglong2 pm23 = __ldcg(reinterpret_cast<const ulonglong2 *>(Pcur + 4 * (size_t)b + 2));
        const double pmax = fmax(fmax(__longlong_as_double((long long)pm01.x), __longlong_as_double((long long)pm01.y)),
                                 fmax(__longlong_as_double((long long)pm23.x), __longlong_as_double((long long)pm23.y)));
        const double bound = fma(abeta, pmax, rmax) * (1.0 + 0x1p-40);
        const int E = (bound > 0.0) ? (int)((__double_as_longlong(bound) >> 52) & 0x7ff) - 1023 + 1 : 0;
        const double scale = scalbn(1.0, 54 - E);
        if (t == 0 && 2u * b >= h0) __stcg(Rnext + b, 0ull);           // max |r| of the NEXT iteration starts from zero
        if (t == 0) TL5(10);
        mbar_wait_guarded(&mb[B5_RP_FULL], u & 1);
        if (t == 0) TL5(11);
        if (u > 0) mbar_wait_guarded(&mb[B5_MMA_DONE + ((u - 1) & 1)], ((u - 1) >> 1) & 1);   // digit image free
        if (t == 0) TL5(12);
        double mx = 0.0;
#pragma unroll
        for (int z = 0; z < 2; ++z) {                        // one column of the pair at a time
          uint32_t lo[16], hi[16];
          double pp0 = 0.0, pp1 = 0.0, pr0 = 0.0, pr1 = 0.0;
#pragma unroll
          for (int ii = 0; ii < 16; ++ii) {
            const unsigned grow = r0 + 16u * g + ii;
            const bool valid = grow < n_rows32, ok = valid && mine;
            // branch-free (selects): rows beyond n and, in the first iteration, the p_old tile hold stale shared memory
            double rv = *reinterpret_cast<const double *>(rrow + ii * 256u + 8u * z);
            double po = *reinterpret_cast<const double *>(prow + ii * 256u + 8u * z);
            rv = valid ? rv : 0.0;
            po = (valid && kk) ? po : 0.0;
            const double pv = fma(beta, po, -rv);                    // l.420; first iteration: beta = 0, p = -r (l.256)
            if (ok) __stcg(p_new + (size_t)grow * ST_P + 2 * cp + z, pv);
            const double pm = ok ? pv : 0.0;
            if (ii & 1) { pp1 = fma(pm, pm, pp1); pr1 = fma(pm, rv, pr1); }
            else        { pp0 = fma(pm, pm, pp0); pr0 = fma(pm, rv, pr0); }
            mx = fmax(mx, fabs(pv));
            const unsigned long long uu = ((unsigned long long)__double2ll_rn(pv * scale) + TC_DIGIT_BIAS) ^ TC_DIGIT_BIAS;
            lo[ii] = (uint32_t)uu;
            hi[ii] = (uint32_t)(uu >> 32);
          }
          // exact-reduction unit: this thread's 16 elements of one column (inside one half block)
          fixacc_add(fa0, pp0 + pp1, fq0, ovf);
          fixacc_add(fa1, pr0 + pr1, fq1, ovf);
          const uint32_t off = sw128_chunk_off((uint32_t)(2 * cp + z), (uint32_t)g);
#pragma unroll
          for (int sl = 0; sl < TC_SLICES; ++sl) {
            const int d = 6 - sl;                                   // digit index held by slice sl
            const uint32_t sel = (d & 3) | (((d & 3) + 4) << 4);    // byte d of a -> pos 0, byte d of b -> pos 1
            uint32_t wq[4];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint32_t x0 = d < 4 ? lo[4 * q4] : hi[4 * q4], x1 = d < 4 ? lo[4 * q4 + 1] : hi[4 * q4 + 1];
              const uint32_t x2 = d < 4 ? lo[4 * q4 + 2] : hi[4 * q4 + 2], x3 = d < 4 ? lo[4 * q4 + 3] : hi[4 * q4 + 3];
              const uint32_t t01 = __byte_perm(x0, x1, sel), t23 = __byte_perm(x2, x3, sel);
              wq[q4] = __byte_perm(t01, t23, 0x5410);
            }
            *reinterpret_cast<uint4 *>(Qsm + sl * TC_QTILE + off) = make_uint4(wq[0], wq[1], wq[2], wq[3]);
          }
        }
        mbar_arrive(&mb[B5_RP_EMPTY]);                        // the stage may be refilled with the next block
        fence_proxy_async_smem();
        if (t == 0) ms.s_E[u & 3] = E;
        mbar_arrive(&mb[B5_Q_FULL]);
        if (t == 0) TL5(15);
        // max |p| of this block (all 128 rows) for the next iteration's bound: one word per L warp, no barrier
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) __stcg(Pnext + 4 * (size_t)b + (warp - 4), (unsigned long long)__double_as_longlong(mx));
        // a |p| above the bound cannot happen with finite data; non-finite data is flagged by the exact accumulators
        if (!(mx <= bound)) ovf = 1u;
      }
      fixacc_flush(fa0, sacc + SC_PP * KUL_STRIDE, fe0);
      fixacc_flush(fa1, sacc + SC_PR * KUL_STRIDE, fe1);
      if (ovf) atomicOr((unsigned long long *)(set + ACC_FLAG_OFF), 1ull);
    } else if constexpr (ROLE == 2) {
      // ===== M: TMEM read-back, W = Z - p S, stores, partial sums =====
      const int w = warp - 8, qd = w & 3, hc = w >> 2;       // TMEM lane quarter (= warp % 4), column half
      FixAcc fa0 = {0, 0}, fa1 = {0, 0};                     // <p,W>, <W,W>
      const int fe0 = ms.s_fe[SC_PHP], fe1 = ms.s_fe[SC_HPHP];
      const double fq0 = scalbn(1.0, 90 - fe0), fq1 = scalbn(1.0, 90 - fe1);
      unsigned ovf = 0;
      for (int i = 0; i < nb_local; ++i) {
        const unsigned u = use + i, b = bfirst + i, r0 = b * ST_NB;
        const bool own0 = 2u * b >= h0 && 2u * b < h1, own1 = 2u * b + 1u >= h0 && 2u * b + 1u < h1;
        if (tid == 256) TL5(21);
        // MMAs of block u complete: the accumulators are ready, and -- through the release / acquire chain L -> MMA
        // issuer -> tcgen05.commit -- every L thread's stores of p (and E) for this block are visible.  (Waiting on
        // Q_FULL here instead would alias: L may run two blocks ahead of M, and an mbarrier only tells odd from even.)
        mbar_wait_guarded(&mb[B5_MMA_DONE + (u & 1)], (u >> 1) & 1);
        if (tid == 256) TL5(22);
        tc_fence_after();
        const int E = ms.s_E[u & 3];
        const double sc = scalbn(1.0, __ldg(plane_exp + b) + E + 10);
        const uint32_t tacc = tmem_base + (u & 1) * TC_TMEM_COLS + ((uint32_t)(32 * qd) << 16) + 16 * hc;
        // The two 16-lane groups (half blocks) of this warp's TMEM lane quarter, one after the other (one copy of the
        // code, small register footprint: with 226 KB of shared memory the L1 is 2 KB, a spill is an L2 round trip).
        // Per group: the rows of p as fp64 MMA A fragments (strip s = rows 8 s + m of the group; L2 loads issued first,
        // in flight during the TMEM read-back -- measured: an exposed L2 load costs 2-3 us under this traffic),
        // Z = A p from TMEM in the accumulator arrangement, W = Z - p S, stores, partial sums.
        // (Measured alternative: both groups interleaved with the p loads after the read-backs: 131 us per iteration
        // instead of 120 -- the loads' latency is exposed; holding both groups' fragments across the read-backs spills.)
#pragma unroll 1
        for (int g16 = 0; g16 < 2; ++g16) {
          if (!(g16 ? own1 : own0)) continue;
          double pa[2][8];                                    // pa[s][qq] = p[row 8 s + m of the group][4 qq + j]
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const unsigned grow = r0 + 64u * g16 + 16u * qd + 8u * s + m;
            const bool ld = grow < n_rows32;
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) pa[s][qq] = ld ? __ldcg(p_new + (size_t)grow * ST_P + 4 * qq + j) : 0.0;
          }
          double acc[2][4];                                   // [tile tt][rows m: c 0,1 ; rows m + 8: c 2,3] (m16n8k16 C layout)
          {
            double out[8];
            recombine_frag16(tacc + ((uint32_t)(16 * g16) << 16), out);
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[0][c] = out[c] * sc; acc[1][c] = out[4 + c] * sc; }
          }
          if (g16 == 1 || !own1) {   // both read-backs of the block done: the accumulator set may be overwritten
            tc_fence_before();
            mbar_arrive(&mb[B5_TMEM_EMPTY + (u & 1)]);
          }
          if (tid == 256) TL5(23 + 3 * g16);
          // W = A p - p S (Ssm holds -S): two 16 x 8 tiles, two k-steps of 16 each (2 chains of 2 large fp64 MMAs)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            double af[8];
#pragma unroll
            for (int ii = 0; ii < 8; ++ii) af[ii] = pa[ii & 1][4 * ks + (ii >> 1)];
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
              double bf[4];
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) bf[ii] = Ssm[(16 * ks + j + 4 * ii) * WS + 16 * hc + 8 * tt + m];
              dmma16816(acc[tt], af, bf);
            }
          }
          if (tid == 256) TL5(24 + 3 * g16);
          {   // the half's W staging area is free once G has finished the Gram of its previous occupant
            const unsigned cnt = (wuse >> (16 * g16)) & 0xffffu;
            if (cnt > 0) mbar_wait_guarded(&mb[B5_W_EMPTY + g16], (cnt - 1) & 1);
          }
          const int src0 = 4 * m + 2 * (j & 1);
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            const unsigned row = 64u * g16 + 16u * qd + 8u * s + m, grow = r0 + row;
            const bool valid = grow < n_rows32;
            double pw = 0.0, ww = 0.0;
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
              const int col = 16 * hc + 8 * tt + 2 * j;
              // p[row][col + c] lives in lane (m, 2 (j & 1) + c) as A-fragment element qq = 2 (2 hc + tt) + (j >> 1)
              const double e0 = hc ? pa[s][4 + 2 * tt] : pa[s][2 * tt];
              const double e1 = hc ? pa[s][5 + 2 * tt] : pa[s][2 * tt + 1];
              const double a0 = __shfl_sync(0xffffffffu, e0, src0), b0 = __shfl_sync(0xffffffffu, e1, src0);
              const double a1 = __shfl_sync(0xffffffffu, e0, src0 + 1), b1 = __shfl_sync(0xffffffffu, e1, src0 + 1);
              const double px = (j >> 1) ? b0 : a0, py = (j >> 1) ? b1 : a1;
              const double wx = acc[tt][2 * s], wy = acc[tt][2 * s + 1];
              pw = fma(px, wx, pw); pw = fma(py, wy, pw);
              ww = fma(wx, wx, ww); ww = fma(wy, wy, ww);
              const double2 wv = make_double2(wx, wy);
              *reinterpret_cast<double2 *>(Wsm + row * WS + col) = wv;
              if (valid) stcg2(a.Hp + (size_t)grow * ST_P + col, wv);
            }
            fixacc_add(fa0, pw, fq0, ovf);   // exact-reduction unit: this lane's 4 elements of the strip row
            fixacc_add(fa1, ww, fq1, ovf);
          }
          mbar_arrive(&mb[B5_W_FULL + g16]);                  // this thread's part of the half's W is staged
          wuse += 1u << (16 * g16);
          if (tid == 256) TL5(25 + 3 * g16);
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
        mbar_wait_guarded(&mb[B5_W_FULL + hsel], (wuse >> (16 * hsel)) & 1);
        const double *Yh = Ysm + 64u * hsel * WS, *Wh = Wsm + 64u * hsel * WS;
        // G[16 mp .. +15][16 np .. +15] over the 64 rows: A = Y^T (16 x 16 per k-step), B = W (16 x 8 per tile);
        // two n-tiles x two k-halves = 4 independent chains of 2 large fp64 MMAs
        double ga[2][2][4];                                   // [n-tile][k-half][c]
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) ga[nt][kh][0] = ga[nt][kh][1] = ga[nt][kh][2] = ga[nt][kh][3] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          double af[8];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) af[ii] = Yh[(16 * ks + j + 4 * (ii >> 1)) * WS + 16 * mp + m + 8 * (ii & 1)];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            double bf[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) bf[ii] = Wh[(16 * ks + j + 4 * ii) * WS + 16 * np + 8 * nt + m];
            dmma16816(ga[nt][ks >> 1], af, bf);
          }
        }
        mbar_arrive(&mb[B5_W_EMPTY + hsel]);                  // M may stage the next occupant of this half
        wuse += 1u << (16 * hsel);
        // c[0], c[1]: Gram row 16 mp + m, columns 16 np + 8 nt + 2 j + {0, 1};  c[2], c[3]: row + 8
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          gram_accumulate(ga[nt][0][0] + ga[nt][1][0], ga[nt][0][1] + ga[nt][1][1], inv_q, gfix[2 * nt], &ovf);
          gram_accumulate(ga[nt][0][2] + ga[nt][1][2], ga[nt][0][3] + ga[nt][1][3], inv_q, gfix[2 * nt + 1], &ovf);
        }
      }
      // gfix[2 nt + hrow] <-> 8 x 8 Gram tile (row tile 2 mp + hrow, column tile 2 np + nt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow)
          gram_flush(set, 4 * (2 * mp + hrow) + 2 * np + nt, lane, gfix[2 * nt + hrow], (nt | hrow) == 0 ? ovf : 0);
    }
    use += (unsigned)nb_local;
    if (tid == 0) TL5B(40);     // S done
    if (tid == 128) TL5B(41);   // L done
    if (tid == 256) TL5B(42);   // M done
    bar_all();
    flush_scalars(sacc, set, 4);
    RedView rvw;
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, 0, ACC_WORDS, rvw,
                             a.dbg ? ms.s_stamp : nullptr, 128u)) { exit_reason = -2; break; }
    if (tid == 256) TL5B(43);     // barrier A released
    // first two strips of phase B for each M warp: start streaming them in before the scalar stage
    int cur0 = -1, cur1 = -1;
    unsigned char *slot0 = base + (ROLE == 2 ? (2 * (warp - 8)) * V5_SLOT : 0), *slot1 = slot0 + V5_SLOT;
    uint64_t *sb0 = &mb[B5_SLOT + (ROLE == 2 ? 2 * (warp - 8) : 0)], *sb1 = sb0 + 1;
    if constexpr (ROLE == 2) {
      if (lane == 0) { cur0 = atomicSub(&ms.s_next_strip, 1); cur1 = atomicSub(&ms.s_next_strip, 1); }
      cur0 = __shfl_sync(0xffffffffu, cur0, 0);
      cur1 = __shfl_sync(0xffffffffu, cur1, 0);
      if (lane == 0) {
        fence_proxy_async_smem();
        fence_proxy_async_global();
        if (cur0 >= s_lo) strip_fetch5(slot0, sb0, cur0, n_rows32, a.Hp, a.s, p_new, a.r, st.Y);
        if (cur1 >= s_lo) strip_fetch5(slot1, sb1, cur1, n_rows32, a.Hp, a.s, p_new, a.r, st.Y);
      }
    }
    {
      const u64 flag = rvw.load(ACC_FLAG_OFF);
      double c = 0.0;
      {
        // G (fixed point) -> shared memory (M and G warps) while the four L warps finalize the four exact scalars
        double *Graw = reinterpret_cast<double *>(base + V5_GRAW);
        if constexpr (ROLE >= 2) {
          for (int e = tid - 256; e < ST_P * ST_P; e += 384) {
            u64 hi, lo;
            if (rvw.world == 1) {
              const ulonglong2 wv = __ldcg(reinterpret_cast<const ulonglong2 *>(rvw.base0 + ACC_GRAM_OFF + 2 * e));
              hi = wv.x; lo = wv.y;
            } else {
              hi = rvw.load(ACC_GRAM_OFF + 2 * e); lo = rvw.load(ACC_GRAM_OFF + 2 * e + 1);
            }
            Graw[e] = fix2_to_double((i64)hi, (i64)lo, q);
          }
        } else if constexpr (ROLE == 1) {
          const int o = (warp - 4) * KUL_STRIDE;              // L warp w finalizes scalar w
          const double x = kul_finalize_warp([&rvw, o](int jj) { return rvw.load(o + jj); });
          if (lane == 0) sh.red[warp - 4] = x;
        }
        bar_all();
        for (int e = tid; e < ST_P * ST_P; e += V5_THREADS) {
          const int i = e >> 5, jj = e & 31;
          const double sg = 0.5 * (Graw[e] + Graw[jj * ST_P + i]);
          Gsm[i * GS + jj] = -sg;
          c = fma(sg, sg, c);
        }
      }
      c = warp_sum(c);
      if (lane == 0) ms.s_part[warp] = c;
      bar_all();
      if (flag != 0) {
        if constexpr (ROLE == 2) {   // drain the outstanding fetches before leaving
          if (cur0 >= s_lo) mbar_wait_guarded(sb0, bpar0);
          if (cur1 >= s_lo) mbar_wait_guarded(sb1, bpar1);
        }
        exit_reason = -3;
        break;
      }
      if constexpr (ROLE == 1) {
        if (warp == 4) {
          // lanes 0..2 evaluate the long-latency operations concurrently, lane 0 takes the decisions
          double nG2 = 0.0;
#pragma unroll
          for (int ww = 0; ww < V5_THREADS / 32; ++ww) nG2 += ms.s_part[ww];
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
      bar_all();
    }
    ++phase;
    if (tid == 256) TL5B(44);     // scalar stage done
    const double step = sh.step;
    if (sh.action != ACT_CONTINUE) {
      if constexpr (ROLE == 2) {
        if (cur0 >= s_lo) mbar_wait_guarded(sb0, bpar0);
        if (cur1 >= s_lo) mbar_wait_guarded(sb1, bpar1);
      }
      const size_t e0 = (size_t)row_lo * ST_P, e1 = (size_t)row_hi * ST_P;
      for (size_t e = e0 + 2 * (size_t)tid; e < e1; e += 2 * (size_t)blockDim.x) {
        double2 sv = ldcg2(a.s + e);
        const double2 pv = ldcg2(p_new + e);
        sv.x = fma(step, pv.x, sv.x);
        sv.y = fma(step, pv.y, sv.y);
        stcg2(a.s + e, sv);
      }
      exit_reason = sh.action - 1;
      break;
    }

    // ------------------------------ phase B ------------------------------
    set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    {
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      const int per = (ACC_WORDS + gridDim.x - 1) / gridDim.x;
      const int z0 = per * blockIdx.x;
      for (int i = tid; i < per && z0 + i < ACC_WORDS; i += blockDim.x) nxt[z0 + i] = 0;
    }
    if constexpr (ROLE == 2) {
      unsigned ovfb = 0;
      FixAcc fb = {0, 0};
      const int feb = ms.s_fe[SC_RV];
      const double fqb = scalbn(1.0, 90 - feb);
      int sl = 0;
      for (;;) {
        const int sidx = sl ? cur1 : cur0;
        if (sidx < s_lo) break;     // strips are handed out in decreasing order: the other slot holds nothing either
        unsigned char *slot = sl ? slot1 : slot0;
        uint64_t *sb = sl ? sb1 : sb0;
        int nxt = 0;
        if (lane == 0) nxt = atomicSub(&ms.s_next_strip, 1);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        const unsigned grow = (unsigned)sidx * 8u + m;
        const bool valid = grow < n_rows32;
        const size_t rowoff = (size_t)grow * ST_P;
        mbar_wait_guarded(sb, sl ? bpar1 : bpar0);
        if (sl) bpar1 ^= 1; else bpar0 ^= 1;
        double acc[4][2];
        double2 sv[4], pv[4], rv[4], yx[4];
        {
          const double *tW = reinterpret_cast<const double *>(slot) + m * ST_P;
          const double *tS = tW + 8 * ST_P, *tP = tS + 8 * ST_P, *tR = tP + 8 * ST_P, *tY = tR + 8 * ST_P;
          unsigned tok = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int col = 8 * t + 2 * j;
            if (valid) {
              const double2 wv = *reinterpret_cast<const double2 *>(tW + col);
              acc[t][0] = wv.x; acc[t][1] = wv.y;
              sv[t] = *reinterpret_cast<const double2 *>(tS + col);
              pv[t] = *reinterpret_cast<const double2 *>(tP + col);
              rv[t] = *reinterpret_cast<const double2 *>(tR + col);
              yx[t] = *reinterpret_cast<const double2 *>(tY + 8 * j + 2 * t);
            } else {
              acc[t][0] = acc[t][1] = 0.0;
              sv[t] = pv[t] = rv[t] = yx[t] = make_double2(0.0, 0.0);
            }
            tok |= __double2hiint(acc[t][0]) | __double2hiint(sv[t].x) | __double2hiint(pv[t].x) |
                   __double2hiint(rv[t].x) | __double2hiint(yx[t].x);
          }
          // every lane's shared-memory reads have returned (tok depends on all of them): the slot may be refilled
          tok = __reduce_or_sync(0xffffffffu, tok);
          if (nxt >= s_lo && lane == 0 && (tok | 1u)) {
            fence_proxy_async_smem();
            strip_fetch5(slot, sb, nxt, n_rows32, a.Hp, a.s, p_new, a.r, st.Y);
          }
        }
        strip_rightmul_v(yx, Gsm, lane, acc);   // Hp = W - Y symG
        double rr = 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int col = 8 * t + 2 * j;
          sv[t].x = fma(step, pv[t].x, sv[t].x);  sv[t].y = fma(step, pv[t].y, sv[t].y);      // l.374
          rv[t].x = fma(step, acc[t][0], rv[t].x); rv[t].y = fma(step, acc[t][1], rv[t].y);   // l.377
          rr = fma(rv[t].x, rv[t].x, rr); rr = fma(rv[t].y, rv[t].y, rr);                      // l.383,408
          if (valid) {
            stcg2(a.s + rowoff + col, sv[t]);
            stcg2(a.r + rowoff + col, rv[t]);
          }
        }
        fixacc_add(fb, rr, fqb, ovfb);      // exact-reduction unit: this lane's 8 elements of the strip
        {   // max |r| per 128-row block for the scale bound of the next phase A (order independent)
          double rm = 0.0;
#pragma unroll
          for (int t = 0; t < 4; ++t) rm = fmax(rm, fmax(fabs(rv[t].x), fabs(rv[t].y)));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, o));
          if (lane == 0)
            atomicMax(a.blk_stats + ((k + 1ull) & 1ull) * a.nblk_stats + ((unsigned)sidx >> 4),
                      (unsigned long long)__double_as_longlong(rm));
        }
        if (sl) cur1 = nxt; else cur0 = nxt;
        sl ^= 1;
      }
      if (tid == 256) TL5B(45);   // phase B strips done (warp 8)
      fixacc_flush(fb, sacc + SC_RV * KUL_STRIDE, feb);
      if (ovfb) atomicAdd(sacc + SC_RV * KUL_STRIDE + KUL_LIMBS, 1ull);   // non-finite / bound violated: poison <r,r>
    }
    bar_all();
    if (tid == 256) TL5B(46);     // phase B done, CTA-wide
    if (tid == 0) ms.s_next_strip = s_hi - 1;
    flush_scalars(sacc + SC_RV * KUL_STRIDE, set + SC_RV * KUL_STRIDE, 1);
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, SC_RV * KUL_STRIDE,
                             KUL_STRIDE, rvw, a.dbg ? ms.s_stamp + 2 : nullptr, 128u)) { exit_reason = -2; break; }
    if constexpr (ROLE == 1) {
      if (warp == 4) {
        const int o = SC_RV * KUL_STRIDE;
        const double x = kul_finalize_warp([&rvw, o](int jj) { return rvw.load(o + jj); });
        if (lane == 0) sh.red[SC_RV] = x;
      }
    }
    bar_all();
    if (tid == 128) {
      update_after_B(sh, sh.red[SC_RV]);
      // bounds for the next iteration's exact accumulators (integer exponent arithmetic only)
      const int e = half_exponent(st.op_norm_bound * st.op_norm_bound * sh.pk_M_2 * 16.0) + 2;   // |G_ij| <= ||H|| ||p||
      ms.s_invq = scalbn(1.0, 90 - e);
      ms.s_q = scalbn(1.0, e - 90);
      ms.s_fe[SC_PHP] = fixacc_exponent(st.op_norm_bound * sh.pk_M_2);
      ms.s_fe[SC_HPHP] = fixacc_exponent(st.op_norm_bound * st.op_norm_bound * sh.pk_M_2);
      ms.s_fe[SC_PP] = fixacc_exponent(sh.pk_M_2);
      ms.s_fe[SC_PR] = half_exponent(sh.pk_M_2 * sh.rv) + 2;                                      // |<p,r>| <= ||p|| ||r||
    }
    bar_all();
    ++phase;
    if (tid == 256) TL5B(47);     // iteration done
    if (a.dbg && tid == 0) {   // [work A, wait A, work B, wait B]; work = previous release -> arrival
      if (dbg_prev) atomicAdd(a.dbg + 4 * blockIdx.x + 0, ms.s_stamp[0] - dbg_prev);
      atomicAdd(a.dbg + 4 * blockIdx.x + 1, ms.s_stamp[1] - ms.s_stamp[0]);
      atomicAdd(a.dbg + 4 * blockIdx.x + 2, ms.s_stamp[2] - ms.s_stamp[1]);
      atomicAdd(a.dbg + 4 * blockIdx.x + 3, ms.s_stamp[3] - ms.s_stamp[2]);
      dbg_prev = ms.s_stamp[3];
    }
  }

  tc_fence_before();
  bar_all();
  if constexpr (ROLE == 0) {
    if (warp == 0) tmem_dealloc(tmem_base, 512);
    if (blockIdx.x == 0 && tid == 0) {
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
}

__global__ void __launch_bounds__(V5_THREADS, 1)
tcg_stiefel_v5_kernel(TcgCommon a, StiefelArgs st, const unsigned char *planes, const int *plane_exp) {
  // the dynamic shared-memory window is declared 1024-byte aligned (SWIZZLE_128B operand images); all pointers are
  // derived from the array itself so that the compiler keeps them in the shared address space (LDS / STS, not generic)
  unsigned char *base = v5_smem_raw;
  if (smem_u32(base) & 1023u) __trap();
  V5Misc &ms = *reinterpret_cast<V5Misc *>(base + V5_MISC);
  u64 *sacc = reinterpret_cast<u64 *>(base + V5_ACC);
  double *Ssm = reinterpret_cast<double *>(base + V5_S);
  uint64_t *mb = reinterpret_cast<uint64_t *>(base + V5_BAR);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (int)(V5_NACC * KUL_STRIDE); i += blockDim.x) sacc[i] = 0;
  for (int e = tid; e < ST_P * ST_P; e += blockDim.x) Ssm[(e >> 5) * WS + (e & 31)] = -st.S[e];
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
    mbar_init(&mb[B5_RP_FULL], 1);
    mbar_init(&mb[B5_RP_EMPTY], 128);
    mbar_init(&mb[B5_A_FULL], 1);
    mbar_init(&mb[B5_Q_FULL], 128);
    mbar_init(&mb[B5_MMA_DONE], 1);
    mbar_init(&mb[B5_MMA_DONE + 1], 1);
    mbar_init(&mb[B5_TMEM_EMPTY], 256);
    mbar_init(&mb[B5_TMEM_EMPTY + 1], 256);
    for (int s = 0; s < (int)V5_NSLOT; ++s) mbar_init(&mb[B5_SLOT + s], 1);
    mbar_init(&mb[B5_W_FULL], 256);
    mbar_init(&mb[B5_W_FULL + 1], 256);
    mbar_init(&mb[B5_W_EMPTY], 128);
    mbar_init(&mb[B5_W_EMPTY + 1], 128);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&ms.s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // register budgets per warp group: 640 threads are launched with 96 registers each (61440 in the CTA's pool)
  //   S 24 x 128 + L 88 x 128 + M 128 x 256 + G 112 x 128 = 61440
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    v5_run<0>(a, st, planes, plane_exp, base);
  } else if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    v5_run<1>(a, st, planes, plane_exp, base);
  } else if (warp < 16) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    v5_run<2>(a, st, planes, plane_exp, base);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    v5_run<3>(a, st, planes, plane_exp, base);
  }
}

cudaError_t launch_tcg_stiefel_v5(const TcgCommon &a, unsigned long long n_rows, const unsigned short *A,
                                  const double *Y, const double *S_dev, double op_norm_bound,
                                  const unsigned char *planes, const int *plane_exp, int sm_count, cudaStream_t stm) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tcg_stiefel_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V5_TOTAL);
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
  return cudaLaunchCooperativeKernel((const void *)tcg_stiefel_v5_kernel, dim3(grid), dim3(V5_THREADS), args,
                                     V5_TOTAL, stm);
}

}  // namespace ob200
