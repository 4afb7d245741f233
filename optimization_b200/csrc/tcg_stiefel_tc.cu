// Persistent fused truncated-CG for the Stiefel trace-minimisation Hessian, v4:
// the block contraction A * p runs on the 5th-generation tensor cores (tcgen05,
// accumulators in TMEM) through the exact int8 digit-plane scheme of tc_common.cuh;
// the small fp64 x fp64 products (p S, the projection Gram Y^T W, Y symG) stay on
// the fp64 tensor cores.  Same iteration structure, reductions and scalar logic as
// tcg_stiefel_kernel (see tcg_stiefel.cu / tcg.cuh for the reference line map).
//
// 16 warps.  Phase A is a software pipeline over the CTA's 128-row blocks with two
// balanced roles (about 5 us per block each, measured):
//   L (warps 0-7)  : block u : stream r, p_old -> p (written back), block maximum, digit
//                    slicing of p into the UMMA operand image; thread 0 bulk-copies (TMA 1-D)
//                    the precomputed A digit planes and issues the 13 tcgen05.mma of the block;
//                    block u-1 : TMEM -> registers, integer recombination, Z = A p -> Wsm[u-1 & 1]
//   M (warps 8-15) : block u : W = Z - p S (fp64 MMA), W written back, <p,W>, <W,W>,
//                    projection Gram Y^T W, exact accumulation of all partial sums
// hand-offs through mbarriers only (TMA / UMMA completion, Z full, W buffer empty).
#include "tcg.cuh"
#include "stiefel_dev.cuh"
#include "tc_common.cuh"

namespace ob200 {
using namespace tc;

constexpr int V3_THREADS = 512;
constexpr size_t V3_A = 0;                                   // 48 KB int8 digit planes of A (1024-aligned)
constexpr size_t V3_Q = V3_A + TC_ABLOCK;                    // 28 KB int8 digit image of p
constexpr size_t V3_WBUF = sizeof(double) * ST_NB * WS;      // 128 x 36 doubles
constexpr size_t V3_W = V3_Q + TC_QBYTES;                    // two W tiles (L fills one while M works on the other)
constexpr size_t V3_Y = V3_W + 2 * V3_WBUF;
constexpr size_t V3_S = V3_Y + sizeof(double) * ST_NB * WS;
constexpr size_t V3_G = V3_S + sizeof(double) * ST_P * WS;
constexpr size_t V3_ACC = V3_G + sizeof(double) * ST_P * WS;
constexpr size_t V3_BAR = V3_ACC + sizeof(u64) * ACC_NSCAL * KUL_STRIDE;
constexpr size_t V3_TOTAL = V3_BAR + 256 + 1024;             // + alignment slack

// mbarrier slots
enum { MB_A_FULL = 0, MB_Q_FULL = 1, MB_MMA_DONE = 2 /*,3*/, MB_Z_FULL = 4 /*,5*/, MB_W_EMPTY = 6 /*,7*/ };

__device__ __forceinline__ void bulk_prefetch_l2(const void *g, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Two 8x8 tiles (mt, nt0) and (mt, nt0 + 1) of Y^T W over one 64-row half block; the Y
// fragments (shared by both tiles) come straight from global memory (L1, prefetched).
__device__ __forceinline__ void gram_pair_half(const double *Yg /* row 0 of the half, or null */, int rows_valid,
                                               const double *Zsm, int mt, int nt0, int lane, double (&g)[2][2]) {
  const int m = lane >> 2, j = lane & 3;
  g[0][0] = g[0][1] = g[1][0] = g[1][1] = 0.0;
#pragma unroll
  for (int qq = 0; qq < 2; ++qq) {
    double av[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int krow = 4 * (8 * qq + q) + j;
      av[q] = (krow < rows_valid) ? __ldg(Yg + (size_t)krow * ST_P + 8 * mt + m) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int krow = 4 * (8 * qq + q) + j;
      dmma884(g[0][0], g[0][1], av[q], Zsm[krow * WS + 8 * nt0 + m]);
      dmma884(g[1][0], g[1][1], av[q], Zsm[krow * WS + 8 * (nt0 + 1) + m]);
    }
  }
}

// Both tiles (mt, nt0), (mt, nt0 + 1) of Y^T W over one whole 128-row block: four independent
// accumulator chains per tile over k (8 steps each; the fp64 MMA has a long dependent-issue latency),
// added in a fixed order.  The block is the exact-reduction unit of the Gram.
__device__ __forceinline__ void gram_pair_block(const double *Xsm, const double *Zsm, int mt, int nt0, int lane,
                                                double &g00, double &g01, double &g10, double &g11) {
  const int m = lane >> 2, j = lane & 3;
  double a0[4][2], a1[4][2];
#pragma unroll
  for (int s = 0; s < 4; ++s) { a0[s][0] = a0[s][1] = a1[s][0] = a1[s][1] = 0.0; }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int krow = 4 * (8 * s + q) + j;
      const double x = Xsm[krow * WS + 8 * mt + m];
      dmma884(a0[s][0], a0[s][1], x, Zsm[krow * WS + 8 * nt0 + m]);
      dmma884(a1[s][0], a1[s][1], x, Zsm[krow * WS + 8 * (nt0 + 1) + m]);
    }
  }
  g00 = (a0[0][0] + a0[1][0]) + (a0[2][0] + a0[3][0]);
  g01 = (a0[0][1] + a0[1][1]) + (a0[2][1] + a0[3][1]);
  g10 = (a1[0][0] + a1[1][0]) + (a1[2][0] + a1[3][0]);
  g11 = (a1[0][1] + a1[1][1]) + (a1[2][1] + a1[3][1]);
}

// Two 8-row strips (rows 8*mw.. and 64+8*mw.. of the block) of W = (A p) - p S on the fp64 tensor
// cores, interleaved so that eight independent accumulator chains are in flight: accumulators start
// from the tcgen05 result staged in Wsm, A fragments pa0 / pa1 hold the strips' rows of p; W goes back
// to Wsm (for the projection Gram) and to the Hp buffer; partials of <p,W> and <W,W>.
__device__ __forceinline__ void ps_strips(const double (&pa0)[8], const double (&pa1)[8], bool own0, bool own1, int mw,
                                          unsigned r0, double *Wsm, const double *Ssm, double *Hp,
                                          unsigned n_rows, int lane, FixAcc &fa0, FixAcc &fa1, double fq0,
                                          double fq1, unsigned &ovf, unsigned long long *tl = nullptr) {
  const int m = lane >> 2, j = lane & 3;
  const int row0 = 8 * mw + m, row1 = 64 + 8 * mw + m;
  double acc0[4][2], acc1[4][2];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const double2 w0 = *reinterpret_cast<const double2 *>(Wsm + row0 * WS + 8 * t + 2 * j);
    const double2 w1 = *reinterpret_cast<const double2 *>(Wsm + row1 * WS + 8 * t + 2 * j);
    acc0[t][0] = w0.x; acc0[t][1] = w0.y;
    acc1[t][0] = w1.x; acc1[t][1] = w1.y;
  }
  if (tl) tl[14] = globaltimer_ns();
#pragma unroll
  for (int qq = 0; qq < 8; ++qq) {                        // W = A p - p S   (Ssm holds -S)
    const double *Mrow = Ssm + (4 * qq + j) * WS + m;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const double sv = Mrow[8 * t];
      dmma884(acc0[t][0], acc0[t][1], pa0[qq], sv);
      dmma884(acc1[t][0], acc1[t][1], pa1[qq], sv);
    }
  }
  if (tl) tl[15] = globaltimer_ns();
  const int src0 = 4 * m + 2 * (j & 1);
#pragma unroll
  for (int hs = 0; hs < 2; ++hs) {
    if (!(hs ? own1 : own0)) continue;
    const int row = hs ? row1 : row0;
    const unsigned grow = r0 + row;
    const bool valid = grow < n_rows;
    double pw = 0.0, ww = 0.0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int col = 8 * t + 2 * j;
      const double e0 = hs ? pa1[2 * t] : pa0[2 * t], e1 = hs ? pa1[2 * t + 1] : pa0[2 * t + 1];
      // p[row][8t + 2j + c] lives in lane (m, 2(j&1) + c) as A-fragment element q = 2t + (j >> 1)
      const double a0 = __shfl_sync(0xffffffffu, e0, src0), b0 = __shfl_sync(0xffffffffu, e1, src0);
      const double a1 = __shfl_sync(0xffffffffu, e0, src0 + 1), b1 = __shfl_sync(0xffffffffu, e1, src0 + 1);
      const double px = (j >> 1) ? b0 : a0, py = (j >> 1) ? b1 : a1;
      const double wx = hs ? acc1[t][0] : acc0[t][0], wy = hs ? acc1[t][1] : acc0[t][1];
      pw = fma(px, wx, pw); pw = fma(py, wy, pw);
      ww = fma(wx, wx, ww); ww = fma(wy, wy, ww);
      const double2 wv = make_double2(wx, wy);
      *reinterpret_cast<double2 *>(Wsm + row * WS + col) = wv;
      if (valid) stcg2(Hp + (size_t)grow * ST_P + col, wv);
    }
    fixacc_add(fa0, pw, fq0, ovf);   // exact-reduction unit: this lane's 8 elements of the strip
    fixacc_add(fa1, ww, fq1, ovf);
  }
  if (tl) tl[16] = globaltimer_ns();
}

// Phase B staging: the five 2 KB tiles (W, s, p, r, Y) of one 8-row strip are fetched with bulk
// asynchronous copies (TMA 1-D) into this warp's 10 KB shared-memory slot, so the next strip streams in
// while the current one is being processed.
constexpr uint32_t STRIP_TILE = 8 * ST_P * sizeof(double);   // 2 KB
constexpr uint32_t STRIP_SLOT = 5 * STRIP_TILE;              // 10 KB per warp
static_assert(16 * STRIP_SLOT + 8192 <= V3_S, "phase-B staging must not overlap S / G / the CTA accumulators");
__device__ __forceinline__ void strip_fetch(unsigned char *slot, uint64_t *bar, int sidx, unsigned n_rows,
                                            const double *W, const double *S, const double *Pn, const double *R,
                                            const double *Y) {
  const unsigned row0 = (unsigned)sidx * 8u;
  const unsigned rows = n_rows - row0 < 8u ? n_rows - row0 : 8u;
  const uint32_t bytes = rows * ST_P * (uint32_t)sizeof(double);
  const size_t off = (size_t)row0 * ST_P;
  mbar_expect_tx(bar, 5 * bytes);
  bulk_g2s(slot, W + off, bytes, bar);
  bulk_g2s(slot + STRIP_TILE, S + off, bytes, bar);
  bulk_g2s(slot + 2 * STRIP_TILE, Pn + off, bytes, bar);
  bulk_g2s(slot + 3 * STRIP_TILE, R + off, bytes, bar);
  bulk_g2s(slot + 4 * STRIP_TILE, Y + off, bytes, bar);
}

// stand-alone HVP mode: only the W and Y tiles of the strip
__device__ __forceinline__ void strip_fetch_wy(unsigned char *slot, uint64_t *bar, int sidx, unsigned n_rows,
                                               const double *W, const double *Y) {
  const unsigned row0 = (unsigned)sidx * 8u;
  const unsigned rows = n_rows - row0 < 8u ? n_rows - row0 : 8u;
  const uint32_t bytes = rows * ST_P * (uint32_t)sizeof(double);
  const size_t off = (size_t)row0 * ST_P;
  mbar_expect_tx(bar, 2 * bytes);
  bulk_g2s(slot, W + off, bytes, bar);
  bulk_g2s(slot + 4 * STRIP_TILE, Y + off, bytes, bar);
}

extern __shared__ __align__(16) unsigned char v3_smem_raw[];

// debug timeline (CTA 0, third block of an iteration): slot <- globaltimer
#ifdef OB200_TIMELINE_BUILD
#define TL(slot) do { if (a.dbg && blockIdx.x == 0 && i == 2 && lane == 0 && (warp == 0 || warp == 8)) \
    a.dbg[4096 + (slot)] = globaltimer_ns(); } while (0)
#define TLB(slot) do { if (a.dbg && blockIdx.x == 0 && tid == 0) a.dbg[4096 + 20 + (slot)] = globaltimer_ns(); } while (0)
#else
#define TL(slot) do { } while (0)
#define TLB(slot) do { } while (0)
#endif

// MODE 0: the whole Steihaug-Toint solve.  MODE 1: ONE stand-alone Hessian-vector product out = Hess f(Y)[V]
// (reference call sites IterativeSolvers.h:294, TNT.h:512) with the same two fused phases and no host round trip:
// a.r = V (input), a.Hp = W workspace, a.s = out, a.g -> device scalar <V,V> (bound for the exact fixed-point Gram),
// `planes_sum_dev` / `planes_sum_expected`: the content checksum of A the digit planes were built from (a mismatch
// ends the launch with status 6: the host rebuilds the planes and relaunches).
template <int MODE>
__global__ void __launch_bounds__(V3_THREADS, 1)
tcg_stiefel_tc_kernel(TcgCommon a, StiefelArgs st, const unsigned char *planes, const int *plane_exp,
                      const unsigned long long *planes_sum_dev, unsigned long long planes_sum_expected) {
  __shared__ CgShared sh;
  __shared__ double s_part[16];
  __shared__ double s_invq, s_q;
  __shared__ int s_lmax[16];
  __shared__ int s_next_strip;
  __shared__ __align__(8) uint64_t s_bmb[16];   // phase B: one mbarrier per warp (strip staging)
  __shared__ int s_fe[5];      // fixacc exponents: <p,W>, <W,W>, <p,p>, <p,r>, <r,r>
  __shared__ uint32_t s_tmem;
  __shared__ unsigned long long s_stamp[4];   // barrier arrival / release times (profiling aid)
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(v3_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char *Asm = base + V3_A;
  unsigned char *Qsm = base + V3_Q;
  double *Wsm = reinterpret_cast<double *>(base + V3_W);
  double *Ysm = reinterpret_cast<double *>(base + V3_Y);
  double *Ssm = reinterpret_cast<double *>(base + V3_S);
  double *Gsm = reinterpret_cast<double *>(base + V3_G);
  u64 *sacc = reinterpret_cast<u64 *>(base + V3_ACC);
  uint64_t *mb = reinterpret_cast<uint64_t *>(base + V3_BAR);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = lane >> 2, j = lane & 3;
  const bool is_L = warp < 8;
  const int mw = warp - 8;
  if (MODE == 1) {
    a.rv0 = __ldcg(a.g);
    a.target = -1.0;                                         // never "converged": the single pass always runs
    a.max_iterations = 1;
    if (__ldcg(planes_sum_dev) != planes_sum_expected) {     // uniform over the grid: stale digit planes
      if (blockIdx.x == 0 && tid == 0) { a.result->status = 6; a.result->exit_reason = -1; a.result->phases = 0; }
      return;
    }
  }
  for (int i = tid; i < ACC_NSCAL * KUL_STRIDE; i += blockDim.x) sacc[i] = 0;
  for (int e = tid; e < ST_P * ST_P; e += blockDim.x) Ssm[(e >> 5) * WS + (e & 31)] = -st.S[e];
  if (tid == 0) {
    sh.rv = a.rv0;
    sh.sk_M_pk = 0.0;
    sh.sk_M_2 = 0.0;
    sh.pk_M_2 = a.rv0;
    sh.alpha = sh.beta = sh.kappa = sh.step = 0.0;
    sh.k = 0;
    sh.action = ACT_CONTINUE;
    sh.status = 0;
    s_next_strip = -1;     // set below once the partition is known
    const int e = gram_exponent(st.op_norm_bound * sqrt(a.rv0) * 4.0);
    s_invq = scalbn(1.0, 90 - e);
    s_q = scalbn(1.0, e - 90);
    s_fe[SC_PHP] = fixacc_exponent(st.op_norm_bound * a.rv0);                         // |<p,W>| <= ||H|| ||p||^2
    s_fe[SC_HPHP] = fixacc_exponent(st.op_norm_bound * st.op_norm_bound * a.rv0);
    s_fe[SC_PP] = fixacc_exponent(a.rv0);                                             // ||p||^2 = pk_M_2 (l.266)
    s_fe[SC_PR] = fixacc_exponent(a.rv0);                                             // |<p,r>| <= ||p|| ||r||
    mbar_init(&mb[MB_A_FULL], 1);
    mbar_init(&mb[MB_Q_FULL], 256);
    mbar_init(&mb[MB_MMA_DONE], 1);
    mbar_init(&mb[MB_MMA_DONE + 1], 1);
    mbar_init(&mb[MB_Z_FULL], 256);
    mbar_init(&mb[MB_Z_FULL + 1], 256);
    mbar_init(&mb[MB_W_EMPTY], 256);
    mbar_init(&mb[MB_W_EMPTY + 1], 256);
    for (int w = 0; w < 16; ++w) mbar_init(&s_bmb[w], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  // ownership: whole 128-row operator blocks [bfirst, bfirst + nb_local)
  // (32-bit row bookkeeping: local row counts stay below 2^31)
  const unsigned n_rows32 = (unsigned)st.n_rows;
  const unsigned nblk = (n_rows32 + ST_NB - 1u) / ST_NB;
  const unsigned bfirst = (unsigned)((unsigned long long)nblk * blockIdx.x / gridDim.x);
  const unsigned bend = (unsigned)((unsigned long long)nblk * (blockIdx.x + 1ull) / gridDim.x);
  const int nb_local = (int)(bend - bfirst);
  const unsigned row_lo = bfirst * ST_NB;
  const unsigned row_hi = (bend * ST_NB < n_rows32) ? bend * ST_NB : n_rows32;
  // phase B has its own, finer partition: 8-row strips split evenly over the CTAs and handed to the warps
  // dynamically (results do not depend on who processes which strip: exact integer accumulation)
  const unsigned nstrips = (n_rows32 + 7u) >> 3;
  const int s_lo = (int)((unsigned long long)nstrips * blockIdx.x / gridDim.x);
  const int s_hi = (int)((unsigned long long)nstrips * (blockIdx.x + 1ull) / gridDim.x);
  if (tid == 0) s_next_strip = s_hi - 1;
  unsigned bpar = 0;           // parity of this warp's phase-B mbarrier
  unsigned gen = 0, phase = 0;
  unsigned use = 0;            // blocks processed so far by this CTA (mbarrier phase bookkeeping)
  int exit_reason = -1;
  unsigned long long dbg_prev = 0;

  for (;;) {
    const unsigned long long k = sh.k;
    if (k >= a.max_iterations) { exit_reason = 1; break; }
    if (sqrt(sh.rv) <= a.target) { exit_reason = 0; break; }
    const double beta = sh.beta;
    const double *p_old = (k & 1ull) ? a.p1 : a.p0;
    double *p_new = (k & 1ull) ? a.p0 : a.p1;
    const double inv_q = s_invq, q = s_q;
    FixAcc fa0 = {0, 0}, fa1 = {0, 0};            // L: <p,p>, <p,r> ; M: <p,W>, <W,W>
    const int fe0 = is_L ? s_fe[SC_PP] : s_fe[SC_PHP], fe1 = is_L ? s_fe[SC_PR] : s_fe[SC_HPHP];
    const double fq0 = scalbn(1.0, 90 - fe0), fq1 = scalbn(1.0, 90 - fe1);

    // ------------------------------ phase A ------------------------------
    u64 *set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    {   // recycle the set used two phases from now (every CTA clears its slice; a grid barrier intervenes)
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      const int per = (ACC_WORDS + gridDim.x - 1) / gridDim.x;
      const int z0 = per * blockIdx.x;
      for (int i = tid; i < per && z0 + i < ACC_WORDS; i += blockDim.x) nxt[z0 + i] = 0;
    }
    unsigned ovf = 0;
    if (is_L) {
      // ===== L warps: block u -> p, digit image, MMAs ; block u-1 -> TMEM read-back =====
      const int cp = tid & 15, g = tid >> 4;           // columns 2cp, 2cp+1 ; rows 8g .. 8g+7 of the block
      const int q4 = warp & 3, chalf = warp >> 2;      // read-back: TMEM lane quarter, column half
      int E_prev = 0;
      for (int i = 0; i <= nb_local; ++i) {
        const unsigned u = use + i;
        int E_cur = 0, pe_cur = 0;
        if (i < nb_local) {
          const unsigned b = bfirst + i, r0 = b * ST_NB;
          TL(0);
#ifndef OB200_PFDIST
#define OB200_PFDIST 1
#endif
          // the block OB200_PFDIST ahead: r, p_old, Y and the A digit planes -> L2 through the TMA unit (bulk
          // prefetches are queued by the copy engine, not dropped under load like per-thread prefetch hints)
          if (OB200_PFDIST > 0 && tid == 32 && i + OB200_PFDIST < nb_local) {
            const unsigned rn = r0 + OB200_PFDIST * ST_NB;
            const unsigned rows = n_rows32 - rn < ST_NB ? n_rows32 - rn : ST_NB;
            const uint32_t bytes = rows * ST_P * (uint32_t)sizeof(double);
            const size_t noff = (size_t)rn * ST_P;
            bulk_prefetch_l2(a.r + noff, bytes);
            if (k) bulk_prefetch_l2(p_old + noff, bytes);
            bulk_prefetch_l2(planes + (size_t)(b + OB200_PFDIST) * TC_ABLOCK, TC_ABLOCK);
            bulk_prefetch_l2(st.Y + noff, bytes);
          }
          if (i == 0 && tid == 0) {   // every MMA of the previous phase has completed: the A image is free
            mbar_expect_tx(&mb[MB_A_FULL], TC_ABLOCK);
            bulk_g2s(Asm, planes + b * (size_t)TC_ABLOCK, TC_ABLOCK, &mb[MB_A_FULL]);
          }
          pe_cur = __ldg(plane_exp + b);        // (with the loads below: its L2 round trip is hidden, not in the read-back)
          double2 rv[8], po[8];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {      // all 16 loads of this thread in flight at once
            const unsigned grow = r0 + 8 * g + ii;
            rv[ii] = po[ii] = make_double2(0.0, 0.0);
            if (grow < n_rows32) {
              const size_t off = (size_t)grow * ST_P + 2 * cp;
              rv[ii] = ldcg2(a.r + off);
              if (k) po[ii] = ldcg2(p_old + off);
            }
          }
          double p[8][2];
          double pp = 0.0, pr = 0.0;
          int mxh = 0;                          // block maximum of |p| through the (ordered) high words
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const unsigned grow = r0 + 8 * g + ii;
            double2 pv;
            if (MODE == 1) {
              pv = rv[ii];                                  // stand-alone HVP: the operand itself
            } else if (k) {
              pv.x = fma(beta, po[ii].x, -rv[ii].x);        // l.420
              pv.y = fma(beta, po[ii].y, -rv[ii].y);
            } else {
              pv.x = -rv[ii].x;                             // l.256
              pv.y = -rv[ii].y;
            }
            if (grow < n_rows32) {
              if (MODE == 0) stcg2(p_new + (size_t)grow * ST_P + 2 * cp, pv);
              pp = fma(pv.x, pv.x, pp); pp = fma(pv.y, pv.y, pp);
              pr = fma(pv.x, rv[ii].x, pr); pr = fma(pv.y, rv[ii].y, pr);
            }
            p[ii][0] = pv.x;
            p[ii][1] = pv.y;
            mxh = max(mxh, max(__double2hiint(pv.x) & 0x7fffffff, __double2hiint(pv.y) & 0x7fffffff));
          }
          TL(1);
          // next block's r, p_old, Y and A digit planes -> L2, issued AFTER this block's demand loads have
          // returned: the fetch then runs during the slicing / MMA / read-back instead of competing with them
          // exact-reduction unit: this lane's 8 x 2 elements
          fixacc_add(fa0, pp, fq0, ovf);
          fixacc_add(fa1, pr, fq1, ovf);
          mxh = __reduce_max_sync(0xffffffffu, mxh);
          if (lane == 0) s_lmax[(u & 1) * 8 + warp] = mxh;
          nbar_sync(NB_LSYNC, 256);
          mxh = s_lmax[(u & 1) * 8];
#pragma unroll
          for (int w = 1; w < 8; ++w) mxh = max(mxh, s_lmax[(u & 1) * 8 + w]);
          // |p| < 2^E over the block (non-finite data: the Kulisch accumulators flag the partial sums); a block of
          // zeros or subnormals (biased exponent 0) takes E = 0
          const int E = (mxh >> 20) ? (mxh >> 20) - 1023 + 1 : 0;
          E_cur = E + pe_cur;   // exponent of the block's product scale: block maximum of p + plane exponent of A
          TL(2);
          if (i > 0) {   // MMAs of block u-1 complete: digit image and A image are free again
            mbar_wait(&mb[MB_MMA_DONE + ((u - 1) & 1)], ((u - 1) >> 1) & 1);
            if (tid == 0) {
              mbar_expect_tx(&mb[MB_A_FULL], TC_ABLOCK);
              bulk_g2s(Asm, planes + b * (size_t)TC_ABLOCK, TC_ABLOCK, &mb[MB_A_FULL]);
            }
          }
          TL(3);
          slice_tile_to_smem(p, scalbn(1.0, 54 - E), Qsm, tid);
          fence_proxy_async_smem();
          mbar_arrive(&mb[MB_Q_FULL]);
          TL(4);
          if (tid == 0) {
            // every L thread has sliced block u -- and, earlier in program order, drained the TMEM
            // accumulators (u & 1) of block u-2
            mbar_wait(&mb[MB_Q_FULL], u & 1);
            TL(6);
            mbar_wait(&mb[MB_A_FULL], u & 1);
            TL(7);
            tc_fence_after();
            issue_block_mmas(smem_u32(Asm), smem_u32(Qsm), tmem_base + (u & 1) * TC_TMEM_COLS);
            umma_commit(&mb[MB_MMA_DONE + (u & 1)]);
            TL(8);
          }
          __syncwarp();
        }
        if (i > 0) {
          // read back block v = u-1: Z = A p -> Wsm[v & 1]
          const unsigned v = u - 1;
          if (i == nb_local) mbar_wait(&mb[MB_MMA_DONE + (v & 1)], (v >> 1) & 1);
          tc_fence_after();
          if (v >= 2) mbar_wait(&mb[MB_W_EMPTY + (v & 1)], ((v >> 1) - 1) & 1);    // M is done with this tile
          double out[16];
          recombine_row16(tmem_base + (v & 1) * TC_TMEM_COLS + ((uint32_t)(32 * q4) << 16) + 16 * chalf, out);
          tc_fence_before();
          const double sc = scalbn(1.0, E_prev + 10);
          double *wrow = Wsm + (v & 1) * (ST_NB * WS) + tc_row_of_lane((uint32_t)(32 * q4 + lane)) * WS + 16 * chalf;
#pragma unroll
          for (int c = 0; c < 16; c += 2) *reinterpret_cast<double2 *>(wrow + c) = make_double2(out[c] * sc, out[c + 1] * sc);
          mbar_arrive(&mb[MB_Z_FULL + (v & 1)]);
          TL(5);
        }
        E_prev = E_cur;
      }
    } else {
      // ===== M warps =====
      i64 gfix[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
      for (int i = 0; i < nb_local; ++i) {
        const unsigned u = use + i;
        const unsigned b = bfirst + i, r0 = b * ST_NB;
        double *Wb = Wsm + (u & 1) * (ST_NB * WS);
        TL(9);
        {   // Y of this block -> shared memory (free since the Gram of the previous block: NB_MSYNC3)
          const int mt_ = tid - 256, cpy = mt_ & 15, gy = mt_ >> 4;
          double2 yv[8];
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const unsigned grow = r0 + 8 * gy + ii;
            yv[ii] = (grow < n_rows32) ? ldcg2(st.Y + (size_t)grow * ST_P + 2 * cpy) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) *reinterpret_cast<double2 *>(Ysm + (8 * gy + ii) * WS + 2 * cpy) = yv[ii];
        }
        mbar_wait(&mb[MB_Z_FULL + (u & 1)], (u >> 1) & 1);
        TL(10);
        // p rows of this warp's strips as DMMA A fragments (from L2: every L thread stored p of block u
        // before its Z_FULL arrival)
        double pa0[8], pa1[8];
        {
          const unsigned g0 = r0 + 8 * mw + m, g1 = g0 + 64;
          const bool l0 = g0 < n_rows32, l1 = g1 < n_rows32;
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const double *psrc = MODE == 1 ? a.r : p_new;
            pa0[qq] = l0 ? __ldcg(psrc + (size_t)g0 * ST_P + 4 * qq + j) : 0.0;
            pa1[qq] = l1 ? __ldcg(psrc + (size_t)g1 * ST_P + 4 * qq + j) : 0.0;
          }
        }
#ifdef OB200_TIMELINE_BUILD
        ps_strips(pa0, pa1, true, true, mw, r0, Wb, Ssm, a.Hp, n_rows32, lane, fa0, fa1, fq0, fq1, ovf,
                  (a.dbg && blockIdx.x == 0 && i == 2 && tid == 256) ? a.dbg + 4096 : nullptr);
#else
        ps_strips(pa0, pa1, true, true, mw, r0, Wb, Ssm, a.Hp, n_rows32, lane, fa0, fa1, fq0, fq1, ovf);
#endif
        nbar_sync(NB_MSYNC2, 256);                               // W complete in Wb, Y complete in Ysm
        TL(12);
        {
          double g00, g01, g10, g11;
          gram_pair_block(Ysm, Wb, mw >> 1, 2 * (mw & 1), lane, g00, g01, g10, g11);
          gram_accumulate(g00, g01, inv_q, gfix[0], &ovf);
          gram_accumulate(g10, g11, inv_q, gfix[1], &ovf);
        }
        mbar_arrive(&mb[MB_W_EMPTY + (u & 1)]);                  // L may refill this W tile
        nbar_sync(NB_MSYNC3, 256);                               // Ysm free for the next block
        TL(13);
      }
      gram_flush(set, 2 * mw, lane, gfix[0], ovf);
      gram_flush(set, 2 * mw + 1, lane, gfix[1], 0);
    }
    use += (unsigned)nb_local;
    fixacc_flush(fa0, sacc + (is_L ? SC_PP : SC_PHP) * KUL_STRIDE, fe0);
    fixacc_flush(fa1, sacc + (is_L ? SC_PR : SC_HPHP) * KUL_STRIDE, fe1);
    if (is_L && ovf) atomicOr((unsigned long long *)(set + ACC_FLAG_OFF), 1ull);
    __syncthreads();
    flush_scalars(sacc, set, 4);
    RedView rvw;
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, 0, ACC_WORDS, rvw,
                             a.dbg ? s_stamp : nullptr)) { exit_reason = -2; break; }
    TLB(0);
    // first strip of phase B for this warp: start streaming it in before the scalar stage
    unsigned char *slot = base + warp * STRIP_SLOT;
    int cur = 0;
    if (lane == 0) cur = atomicSub(&s_next_strip, 1);      // top-down: lines touched last in phase A first
    cur = __shfl_sync(0xffffffffu, cur, 0);
    if (cur >= s_lo && lane == 0) {
      fence_proxy_async_smem();
      if (MODE == 1) strip_fetch_wy(slot, &s_bmb[warp], cur, n_rows32, a.Hp, st.Y);
      else strip_fetch(slot, &s_bmb[warp], cur, n_rows32, a.Hp, a.s, p_new, a.r, st.Y);
    }
    {
      const u64 flag = rvw.load(ACC_FLAG_OFF);
      double c = 0.0;
      {
        // G (fixed point) -> shared memory with one 16-byte load per entry (warps 4-15) while warps 0-3 finalize the
        // four exact scalars (both are latency chains on L2); then sym(G) from shared memory
        double *Graw = reinterpret_cast<double *>(base + 16 * STRIP_SLOT);   // 8 KB of scratch behind the strip slots
        if (warp >= 4) {
          for (int e = tid - 128; e < ST_P * ST_P; e += 384) {
            u64 hi, lo;
            if (rvw.world == 1) {
              const ulonglong2 w = __ldcg(reinterpret_cast<const ulonglong2 *>(rvw.base0 + ACC_GRAM_OFF + 2 * e));
              hi = w.x; lo = w.y;
            } else {
              hi = rvw.load(ACC_GRAM_OFF + 2 * e); lo = rvw.load(ACC_GRAM_OFF + 2 * e + 1);
            }
            Graw[e] = fix2_to_double((i64)hi, (i64)lo, q);
          }
        } else {
          finalize_scalars(rvw, sh, 0, 4);
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = tid + 512 * h;
          const int i = e >> 5, jj = e & 31;
          const double sg = 0.5 * (Graw[e] + Graw[jj * ST_P + i]);
          Gsm[i * GS + jj] = -sg;
          c = fma(sg, sg, c);
        }
      }
      TLB(10);
      TLB(11);
      c = warp_sum(c);
      if (lane == 0) s_part[warp] = c;
      __syncthreads();
      TLB(12);
      if (flag != 0) {
        if (cur >= s_lo) mbar_wait(&s_bmb[warp], bpar);    // drain the outstanding fetch before leaving
        exit_reason = -3;
        break;
      }
      if (MODE == 0 && warp == 0) {
        // lanes 0..2 evaluate the long-latency operations concurrently, lane 0 takes the decisions
        double nG2 = 0.0;
#pragma unroll
        for (int w = 0; w < 16; ++w) nG2 += s_part[w];
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
          s_fe[SC_RV] = fixacc_exponent(2.0 * (sh.rv + sh.step * sh.step * sh.red[SC_HPHP]));
        }
        TLB(13);
      }
      __syncthreads();
    }
    ++phase;
    TLB(1);
    const double step = sh.step;
    if (MODE == 1) {
      // ---- stand-alone HVP, second phase: out = W - Y sym(G), strips as in phase B ----
      while (cur >= s_lo) {
        const int sidx = cur;
        int nxt = 0;
        if (lane == 0) nxt = atomicSub(&s_next_strip, 1);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        const unsigned grow = (unsigned)sidx * 8u + m;
        const bool valid = grow < n_rows32;
        mbar_wait(&s_bmb[warp], bpar);
        bpar ^= 1;
        double acc[4][2];
        double2 yx[4];
        const double *tW = reinterpret_cast<const double *>(slot) + m * ST_P, *tY = tW + 4 * 8 * ST_P;
        unsigned tok = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if (valid) {
            const double2 w = *reinterpret_cast<const double2 *>(tW + 8 * t + 2 * j);
            acc[t][0] = w.x; acc[t][1] = w.y;
            yx[t] = *reinterpret_cast<const double2 *>(tY + 8 * j + 2 * t);
          } else {
            acc[t][0] = acc[t][1] = 0.0;
            yx[t] = make_double2(0.0, 0.0);
          }
          tok |= __double2hiint(acc[t][0]) | __double2hiint(yx[t].x);
        }
        tok = __reduce_or_sync(0xffffffffu, tok);
        if (nxt >= s_lo && lane == 0 && (tok | 1u)) {
          fence_proxy_async_smem();
          strip_fetch_wy(slot, &s_bmb[warp], nxt, n_rows32, a.Hp, st.Y);
        }
        strip_rightmul_v(yx, Gsm, lane, acc);
        if (valid) {
#pragma unroll
          for (int t = 0; t < 4; ++t)
            stcg2(a.s + (size_t)grow * ST_P + 8 * t + 2 * j, make_double2(acc[t][0], acc[t][1]));
        }
        cur = nxt;
      }
      exit_reason = 0;
      break;
    }
    if (sh.action != ACT_CONTINUE) {
      if (cur >= s_lo) mbar_wait(&s_bmb[warp], bpar);      // drain the outstanding fetch before leaving
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
    {   // recycle the set used two phases from now (every CTA clears its slice; a grid barrier intervenes)
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      const int per = (ACC_WORDS + gridDim.x - 1) / gridDim.x;
      const int z0 = per * blockIdx.x;
      for (int i = tid; i < per && z0 + i < ACC_WORDS; i += blockDim.x) nxt[z0 + i] = 0;
    }
    unsigned ovfb = 0;
    {
      FixAcc fb = {0, 0};
      const int feb = s_fe[SC_RV];
      const double fqb = scalbn(1.0, 90 - feb);
      while (cur >= s_lo) {
        const int sidx = cur;
        int nxt = 0;
        if (lane == 0) nxt = atomicSub(&s_next_strip, 1);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        const unsigned grow = (unsigned)sidx * 8u + m;
        const bool valid = grow < n_rows32;
        const size_t rowoff = (size_t)grow * ST_P;
        mbar_wait(&s_bmb[warp], bpar);
        bpar ^= 1;
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
              const double2 w = *reinterpret_cast<const double2 *>(tW + col);
              acc[t][0] = w.x; acc[t][1] = w.y;
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
            strip_fetch(slot, &s_bmb[warp], nxt, n_rows32, a.Hp, a.s, p_new, a.r, st.Y);
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
        cur = nxt;
      }
      TLB(6);
      fixacc_flush(fb, sacc + SC_RV * KUL_STRIDE, feb);
      if (ovfb) atomicAdd(sacc + SC_RV * KUL_STRIDE + KUL_LIMBS, 1ull);   // non-finite / bound violated: poison <r,r>
    }
    __syncthreads();
    TLB(7);
    if (tid == 0) s_next_strip = s_hi - 1;
    flush_scalars(sacc + SC_RV * KUL_STRIDE, set + SC_RV * KUL_STRIDE, 1);
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, SC_RV * KUL_STRIDE,
                             KUL_STRIDE, rvw, a.dbg ? s_stamp + 2 : nullptr)) { exit_reason = -2; break; }
    TLB(8);
    finalize_scalars(rvw, sh, SC_RV, 1);
    __syncthreads();
    if (tid == 0) {
      update_after_B(sh, sh.red[SC_RV]);
      // bounds for the next iteration's exact accumulators (integer exponent arithmetic only)
      const int e = half_exponent(st.op_norm_bound * st.op_norm_bound * sh.pk_M_2 * 16.0) + 2;   // |G_ij| <= ||H|| ||p||
      s_invq = scalbn(1.0, 90 - e);
      s_q = scalbn(1.0, e - 90);
      s_fe[SC_PHP] = fixacc_exponent(st.op_norm_bound * sh.pk_M_2);
      s_fe[SC_HPHP] = fixacc_exponent(st.op_norm_bound * st.op_norm_bound * sh.pk_M_2);
      s_fe[SC_PP] = fixacc_exponent(sh.pk_M_2);
      s_fe[SC_PR] = half_exponent(sh.pk_M_2 * sh.rv) + 2;                                        // |<p,r>| <= ||p|| ||r||
    }
    __syncthreads();
    TLB(9);
    ++phase;
    if (a.dbg && tid == 0) {   // [work A, wait A, work B, wait B]; work = previous release -> arrival
      if (dbg_prev) atomicAdd(a.dbg + 4 * blockIdx.x + 0, s_stamp[0] - dbg_prev);
      atomicAdd(a.dbg + 4 * blockIdx.x + 1, s_stamp[1] - s_stamp[0]);
      atomicAdd(a.dbg + 4 * blockIdx.x + 2, s_stamp[2] - s_stamp[1]);
      atomicAdd(a.dbg + 4 * blockIdx.x + 3, s_stamp[3] - s_stamp[2]);
      dbg_prev = s_stamp[3];
    }
  }

  tc_fence_before();
  __syncthreads();
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

cudaError_t launch_tcg_stiefel_tc(const TcgCommon &a, unsigned long long n_rows, const unsigned short *A,
                                  const double *Y, const double *S_dev, double op_norm_bound,
                                  const unsigned char *planes, const int *plane_exp, int grid, cudaStream_t stm,
                                  int hvp_mode, const unsigned long long *planes_sum_dev,
                                  unsigned long long planes_sum_expected) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tcg_stiefel_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3_TOTAL);
    if (e) return e;
    e = cudaFuncSetAttribute(tcg_stiefel_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3_TOTAL);
    if (e) return e;
    attr = true;
  }
  TcgCommon ac = a;
  StiefelArgs sa{n_rows, A, Y, S_dev, op_norm_bound};
  const unsigned char *pl = planes;
  const int *pe = plane_exp;
  const unsigned long long *sd = planes_sum_dev;
  unsigned long long se = planes_sum_expected;
  void *args[] = {(void *)&ac, (void *)&sa, (void *)&pl, (void *)&pe, (void *)&sd, (void *)&se};
  const void *fn = hvp_mode ? (const void *)tcg_stiefel_tc_kernel<1> : (const void *)tcg_stiefel_tc_kernel<0>;
  return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(V3_THREADS), args, V3_TOTAL, stm);
}

}  // namespace ob200
