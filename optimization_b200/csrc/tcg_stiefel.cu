// Persistent fused truncated-CG for the Stiefel trace-minimisation Hessian
//   Hess f(Y)[V] = P_Y(A V - V S),  P_Y(Z) = Z - Y sym(Y^T Z),  S = sym(Y^T A Y)
// with A block-diagonal (128x128 dense blocks stored as bf16, exact by
// construction) and p = 32 columns, plus the stand-alone model kernels
// (apply / Gram / row-block GEMM) that the TNT outer loop uses.
//
// HBM layout: n x 32 row-major doubles (256 B rows); A: ceil(n/128) blocks of
// 128x128 bf16 row-major (32 KB per block), rows/cols beyond n zero.
//
// One CTA (512 threads = 16 warps) owns a contiguous range of 128-row blocks.
// Per block, phase A:
//   1. all threads: p = -r + beta p_old (written back), p and Y staged in shared memory
//   2. warp w: 8-row strip W = A[8w..8w+7, :] p - p S on the fp64 tensor cores
//      (mma.sync m8n8k4 f64; A fragments go straight from global to registers,
//      converted bf16->f64 on the fly); W written to the Hp buffer and staged
//   3. warp (mt,nt): 8x8 tile of the projection Gram Y^T W over the block,
//      quantised to the bounded two-limb fixed point and accumulated exactly
// phase B (no shared-memory hazards, no CTA barriers):
//   warp w: Hp = W - Y symG on the tensor cores, s += alpha p, r += alpha Hp,
//   partial of <r, r>.
// kappa = <p, Hp> and ||Hp||^2 are obtained from the SAME reduction as the Gram
// (p tangent => <p, Y symG> = 0; Y orthonormal => ||W - Y symG||^2 = ||W||^2 -
// ||symG||_F^2), so a CG step needs two grid barriers, not three.
//
// Algorithmic bytes per CG step: read r, p_old, A(bf16), Y | write p, W |
// read W, Y, s, p, r | write s, r  = 12 N e + 2 * 128 * n  (SURVEY.md 8(d)).
#include "tcg.cuh"
#include "stiefel_dev.cuh"

namespace ob200 {

extern __shared__ __align__(16) unsigned char st_smem[];

// ---- persistent kernel, v2: warp-specialised phase A -----------------------------
// 16 warps.  Phase A: warps 0-7 ("L") stream r, p_old, Y of the next 128-row block,
// form p, write it back and stage p / Y in a double-buffered shared-memory ring;
// warps 8-15 ("M") run the fp64 tensor-core contractions of the current block
// (W strips, then the projection Gram).  L and M hand buffers over with named
// barriers, so HBM streaming overlaps the DMMA pipe.  Work is split between CTAs
// in 64-row half blocks (balanced to 4 %); a block shared by two CTAs is staged
// by both, computed/written only for the owned half.
// Phase B: all 16 warps, 8-row strips, traversed in REVERSE so the p / W / Y / r
// lines written or read last in phase A are still L2 resident.
constexpr size_t V2_PSZ = sizeof(double) * ST_NB * PS;   // 33792
constexpr size_t V2_YSZ = sizeof(double) * ST_NB * WS;   // 36864
constexpr size_t V2_P = 0;
constexpr size_t V2_Y = V2_P + 2 * V2_PSZ;
constexpr size_t V2_W = V2_Y + 2 * V2_YSZ;
constexpr size_t V2_S = V2_W + V2_YSZ;
constexpr size_t V2_G = V2_S + sizeof(double) * ST_P * WS;
constexpr size_t V2_ACC = V2_G + sizeof(double) * ST_P * WS;
constexpr size_t V2_TOTAL = V2_ACC + sizeof(u64) * ACC_NSCAL * KUL_STRIDE;

__global__ void __launch_bounds__(TCG_THREADS, 1) tcg_stiefel_kernel(TcgCommon a, StiefelArgs st) {
  __shared__ CgShared sh;
  __shared__ double s_part[TCG_WARPS];
  __shared__ double s_invq, s_q;
  double *Psm0 = reinterpret_cast<double *>(st_smem + V2_P);
  double *Ysm0 = reinterpret_cast<double *>(st_smem + V2_Y);
  double *Wsm = reinterpret_cast<double *>(st_smem + V2_W);
  double *Ssm = reinterpret_cast<double *>(st_smem + V2_S);
  double *Gsm = reinterpret_cast<double *>(st_smem + V2_G);
  u64 *sacc = reinterpret_cast<u64 *>(st_smem + V2_ACC);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = lane >> 2, j = lane & 3;
  const bool is_L = warp < 8;
  const int mw = warp - 8;
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
    // |G_ij| <= ||Y e_i|| ||W e_j|| <= (||A|| + ||S||) ||p||_F ; ||p||_F^2 = pk_M_2 (l.266,417)
    const int e = gram_exponent(st.op_norm_bound * sqrt(a.rv0) * 4.0);
    s_invq = scalbn(1.0, 90 - e);
    s_q = scalbn(1.0, e - 90);
  }
  __syncthreads();

  // ownership: 64-row half blocks [h0, h1)
  const unsigned long long nhalf = (st.n_rows + 63ull) / 64ull;
  const unsigned long long h0 = nhalf * blockIdx.x / gridDim.x, h1 = nhalf * (blockIdx.x + 1ull) / gridDim.x;
  const unsigned long long row_lo = h0 * 64ull;
  const unsigned long long row_hi = (h1 * 64ull < st.n_rows) ? h1 * 64ull : st.n_rows;
  const unsigned long long bfirst = h0 >> 1;
  const int nb_local = (h1 > h0) ? (int)(((h1 - 1) >> 1) - bfirst + 1) : 0;
  const long long s_lo = (long long)(row_lo >> 3), s_hi = (long long)((row_hi + 7ull) >> 3);   // 8-row strips
  unsigned gen = 0, phase = 0;
  int exit_reason = -1;
  i64 gfix[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  unsigned long long dbg_prev = 0;

  for (;;) {
    const unsigned long long k = sh.k;
    if (k >= a.max_iterations) { exit_reason = 1; break; }
    if (sqrt(sh.rv) <= a.target) { exit_reason = 0; break; }
    const double beta = sh.beta;
    const double *p_old = (k & 1ull) ? a.p1 : a.p0;
    double *p_new = (k & 1ull) ? a.p0 : a.p1;
    const double inv_q = s_invq, q = s_q;

    // ------------------------------ phase A ------------------------------
    u64 *set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    if (blockIdx.x == 0) {
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      for (int i = tid; i < ACC_WORDS; i += blockDim.x) nxt[i] = 0;
    }
    unsigned long long stA[2] = {0, 0}, stB[2] = {0, 0};
    unsigned ovf = 0;
    if (is_L) {
      // ===== loader warps: stream, form p, stage =====
      for (int i = 0; i < nb_local; ++i) {
        const int buf = i & 1;
        const unsigned long long b = bfirst + i, r0 = b * ST_NB;
        double *Psm = Psm0 + buf * (ST_NB * PS);
        double *Ysm = Ysm0 + buf * (ST_NB * WS);
        if (i + 1 < nb_local)   // next block's A (32 KB = 256 lines) into L2 for the M warps
          prefetch_l2(reinterpret_cast<const char *>(st.A + (size_t)(b + 1) * ST_NB * ST_NB) + 128 * tid);
        if (i >= 2) nbar_sync(NB_EMPTY + buf, TCG_THREADS);
#pragma unroll
        for (int chunk = 0; chunk < 2; ++chunk) {   // chunk == 64-row half: the unit of the exact reduction
          double pp = 0.0, pr = 0.0;
          double2 rv[4], yv[4], po[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int idx2 = tid + 256 * (4 * chunk + u);
            const int row = idx2 >> 4, c2 = (idx2 & 15) * 2;
            const unsigned long long grow = r0 + row;
            rv[u] = yv[u] = po[u] = make_double2(0.0, 0.0);
            if (grow < st.n_rows) {
              const size_t off = (size_t)grow * ST_P + c2;
              rv[u] = ldcg2(a.r + off);
              yv[u] = ldcg2(st.Y + off);
              if (k) po[u] = ldcg2(p_old + off);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int idx2 = tid + 256 * (4 * chunk + u);
            const int row = idx2 >> 4, c2 = (idx2 & 15) * 2;
            const unsigned long long grow = r0 + row;
            double2 pv;
            if (k) {
              pv.x = fma(beta, po[u].x, -rv[u].x);        // l.420
              pv.y = fma(beta, po[u].y, -rv[u].y);
            } else {
              pv.x = -rv[u].x;                            // l.256
              pv.y = -rv[u].y;
            }
            if (grow >= row_lo && grow < row_hi) {
              stcg2(p_new + (size_t)grow * ST_P + c2, pv);
              pp = fma(pv.x, pv.x, pp); pp = fma(pv.y, pv.y, pp);
              pr = fma(pv.x, rv[u].x, pr); pr = fma(pv.y, rv[u].y, pr);
            }
            Psm[row * PS + c2] = pv.x;
            Psm[row * PS + c2 + 1] = pv.y;
            *reinterpret_cast<double2 *>(Ysm + row * WS + c2) = yv[u];
          }
          pp = warp_sum(pp);
          pr = warp_sum(pr);
          if (lane == 0) {
            kul_add_atomic(sacc + SC_PP * KUL_STRIDE, pp);
            kul_add_atomic(sacc + SC_PR * KUL_STRIDE, pr);
          }
        }
        nbar_arrive(NB_FULL + buf, TCG_THREADS);
      }
    } else {
      // ===== math warps: W strips and projection Gram on the fp64 tensor cores =====
      for (int i = 0; i < nb_local; ++i) {
        const int buf = i & 1;
        const unsigned long long b = bfirst + i, r0 = b * ST_NB;
        const double *Psm = Psm0 + buf * (ST_NB * PS);
        const double *Ysm = Ysm0 + buf * (ST_NB * WS);
        nbar_sync(NB_FULL + buf, TCG_THREADS);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          double pw = 0.0, ww = 0.0;
          const int strip = half * 8 + mw;
          const int row = 8 * strip + m;
          const unsigned long long grow = r0 + row;
          const unsigned long long hh = 2ull * b + half;
          if (hh >= h0 && hh < h1) {
            double acc[4][2];
            strip_apply(st.A + (size_t)b * ST_NB * ST_NB, Psm, Ssm, strip, lane, acc);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int col = 8 * t + 2 * j;
              const double p0v = Psm[row * PS + col], p1v = Psm[row * PS + col + 1];
              pw = fma(p0v, acc[t][0], pw); pw = fma(p1v, acc[t][1], pw);
              ww = fma(acc[t][0], acc[t][0], ww); ww = fma(acc[t][1], acc[t][1], ww);
              const double2 wv = make_double2(acc[t][0], acc[t][1]);
              *reinterpret_cast<double2 *>(Wsm + row * WS + col) = wv;
              if (grow < st.n_rows) stcg2(a.Hp + (size_t)grow * ST_P + col, wv);
            }
            pw = warp_sum(pw);       // unit of the exact reduction: one 8-row strip
            ww = warp_sum(ww);
            if (lane == 0) {
              kul_add_atomic(sacc + SC_PHP * KUL_STRIDE, pw);
              kul_add_atomic(sacc + SC_HPHP * KUL_STRIDE, ww);
            }
          }
        }
        nbar_sync(NB_MSYNC, 256);
        // projection Gram, one exact unit per owned 64-row half (partition independent)
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const unsigned long long hh = 2ull * b + half;
          if (hh < h0 || hh >= h1) continue;
#pragma unroll
          for (int tt = 0; tt < 2; ++tt) {
            double g0, g1;
            gram_tile_half(Ysm + half * 64 * WS, Wsm + half * 64 * WS, 2 * mw + tt, lane, g0, g1);
            gram_accumulate(g0, g1, inv_q, gfix[tt], &ovf);
          }
        }
        nbar_sync(NB_MSYNC2, 256);
        if (i + 2 < nb_local) nbar_arrive(NB_EMPTY + buf, TCG_THREADS);
      }
      gram_flush(set, 2 * mw, lane, gfix[0], ovf);
      gram_flush(set, 2 * mw + 1, lane, gfix[1], 0);
    }
    __syncthreads();
    flush_scalars(sacc, set, 4);
    RedView rvw;
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, 0, ACC_WORDS, rvw,
                             a.dbg ? stA : nullptr)) { exit_reason = -2; break; }
    {
      // everything that depends on the reduced data is fetched in one round trip
      const u64 flag = rvw.load(ACC_FLAG_OFF);
      double c = 0.0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = tid + TCG_THREADS * h;
        const int i = e >> 5, jj = e & 31, et = jj * ST_P + i;
        const double v1 = fix2_to_double((i64)rvw.load(ACC_GRAM_OFF + 2 * e), (i64)rvw.load(ACC_GRAM_OFF + 2 * e + 1), q);
        const double v2 = fix2_to_double((i64)rvw.load(ACC_GRAM_OFF + 2 * et), (i64)rvw.load(ACC_GRAM_OFF + 2 * et + 1), q);
        const double sg = 0.5 * (v1 + v2);
        Gsm[i * WS + jj] = -sg;
        c = fma(sg, sg, c);
      }
      finalize_scalars(rvw, sh, 0, 4);
      c = warp_sum(c);
      if (lane == 0) s_part[warp] = c;
      __syncthreads();
      if (flag != 0) { exit_reason = -3; break; }
      if (tid == 0) {
        double nG2 = 0.0;
#pragma unroll
        for (int w = 0; w < TCG_WARPS; ++w) nG2 += s_part[w];
        const double nHp2 = fmax(sh.red[SC_HPHP] - nG2, 0.0);
        decide_after_A(sh, sh.red[SC_PHP], nHp2, sh.red[SC_PP], sh.red[SC_PR], a.Delta, a.epsilon);
      }
      __syncthreads();
    }
    ++phase;
    const double step = sh.step;
    if (sh.action != ACT_CONTINUE) {
      const size_t e0 = (size_t)row_lo * ST_P, e1 = (size_t)row_hi * ST_P;
      for (size_t e = e0 + 2 * (size_t)tid; e < e1; e += 2 * TCG_THREADS) {
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
    if (blockIdx.x == 0) {
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      for (int i = tid; i < ACC_WORDS; i += blockDim.x) nxt[i] = 0;
    }
    for (long long sidx = s_hi - 1 - warp; sidx >= s_lo; sidx -= TCG_WARPS) {
      const unsigned long long grow = (unsigned long long)sidx * 8ull + m;
      const bool valid = grow < st.n_rows;
      const size_t rowoff = (size_t)grow * ST_P;
      if (sidx - TCG_WARPS >= s_lo && lane < 16) {   // this warp's next strip -> L2 (5 x 2 KB)
        const size_t noff = (size_t)(sidx - TCG_WARPS) * 8 * ST_P + 16 * lane;
        prefetch_l2(a.Hp + noff); prefetch_l2(a.s + noff); prefetch_l2(p_new + noff);
        prefetch_l2(a.r + noff); prefetch_l2(st.Y + noff);
      }
      double acc[4][2];
      double2 sv[4], pv[4], rv[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int col = 8 * t + 2 * j;
        if (valid) {
          const double2 w = ldcg2(a.Hp + rowoff + col);
          acc[t][0] = w.x; acc[t][1] = w.y;
          sv[t] = ldcg2(a.s + rowoff + col);
          pv[t] = ldcg2(p_new + rowoff + col);
          rv[t] = ldcg2(a.r + rowoff + col);
        } else {
          acc[t][0] = acc[t][1] = 0.0;
          sv[t] = pv[t] = rv[t] = make_double2(0.0, 0.0);
        }
      }
      strip_rightmul(valid ? st.Y + rowoff : nullptr, Gsm, lane, acc);   // Hp = W - Y symG
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
      rr = warp_sum(rr);
      if (lane == 0) kul_add_atomic(sacc + SC_RV * KUL_STRIDE, rr);
    }
    __syncthreads();
    flush_scalars(sacc + SC_RV * KUL_STRIDE, set + SC_RV * KUL_STRIDE, 1);
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, SC_RV * KUL_STRIDE,
                             KUL_STRIDE, rvw, a.dbg ? stB : nullptr)) { exit_reason = -2; break; }
    finalize_scalars(rvw, sh, SC_RV, 1);
    __syncthreads();
    if (tid == 0) {
      update_after_B(sh, sh.red[SC_RV]);
      const int e = gram_exponent(st.op_norm_bound * sqrt(sh.pk_M_2) * 4.0);
      s_invq = scalbn(1.0, 90 - e);
      s_q = scalbn(1.0, e - 90);
    }
    __syncthreads();
    ++phase;
    if (a.dbg && tid == 0) {   // [work A, wait A, work B, wait B]; work = previous release -> arrival
      if (dbg_prev) atomicAdd(a.dbg + 4 * blockIdx.x + 0, stA[0] - dbg_prev);
      atomicAdd(a.dbg + 4 * blockIdx.x + 1, stA[1] - stA[0]);
      atomicAdd(a.dbg + 4 * blockIdx.x + 2, stB[0] - stA[1]);
      atomicAdd(a.dbg + 4 * blockIdx.x + 3, stB[1] - stB[0]);
      dbg_prev = stB[1];
    }
  }

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

// ---------------------------------------------------------------------------
// Stand-alone model kernels (one 128-row block per CTA iteration; any grid).
// ---------------------------------------------------------------------------
// W = A V - V S (S may be null => W = A V); optionally Gram(Y, W) into `set`
// (fixed point, quantum from inv_q) and <V, W>, <W, W> into the scalar slots.
__global__ void __launch_bounds__(TCG_THREADS, 1)
stiefel_apply_kernel(unsigned long long n_rows, const unsigned short *A, const double *V,
                     const double *S /* nullable */, const double *Y /* nullable */, double *Wout,
                     u64 *set, double inv_q) {
  double *Psm = reinterpret_cast<double *>(st_smem + SM_P);
  double *Wsm = reinterpret_cast<double *>(st_smem + SM_W);
  double *Ysm = reinterpret_cast<double *>(st_smem + SM_Y);
  double *Ssm = reinterpret_cast<double *>(st_smem + SM_S);
  u64 *sacc = reinterpret_cast<u64 *>(st_smem + SM_ACC);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, m = lane >> 2, j = lane & 3;
  for (int i = tid; i < ACC_NSCAL * KUL_STRIDE; i += blockDim.x) sacc[i] = 0;
  if (S) for (int e = tid; e < ST_P * ST_P; e += blockDim.x) Ssm[(e >> 5) * WS + (e & 31)] = -S[e];
  __syncthreads();
  const unsigned long long nblk = (n_rows + ST_NB - 1) / ST_NB;
  i64 gfix[4] = {0, 0, 0, 0};
  unsigned ovf = 0;
  for (unsigned long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const unsigned long long r0 = b * ST_NB;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx2 = tid + TCG_THREADS * i;
      const int row = idx2 >> 4, c2 = (idx2 & 15) * 2;
      const unsigned long long grow = r0 + row;
      double2 pv = make_double2(0.0, 0.0), yv = make_double2(0.0, 0.0);
      if (grow < n_rows) {
        const size_t off = (size_t)grow * ST_P + c2;
        pv = ldcg2(V + off);
        if (Y) yv = ldcg2(Y + off);
      }
      Psm[row * PS + c2] = pv.x;
      Psm[row * PS + c2 + 1] = pv.y;
      *reinterpret_cast<double2 *>(Ysm + row * WS + c2) = yv;
    }
    __syncthreads();
    double acc[4][2];
    strip_apply(A + (size_t)b * ST_NB * ST_NB, Psm, S ? Ssm : nullptr, warp, lane, acc);
    {
      const int row = 8 * warp + m;
      const unsigned long long grow = r0 + row;
      double pw = 0.0, ww = 0.0;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int col = 8 * t + 2 * j;
        const double p0v = Psm[row * PS + col], p1v = Psm[row * PS + col + 1];
        pw = fma(p0v, acc[t][0], pw); pw = fma(p1v, acc[t][1], pw);
        ww = fma(acc[t][0], acc[t][0], ww); ww = fma(acc[t][1], acc[t][1], ww);
        const double2 wv = make_double2(acc[t][0], acc[t][1]);
        *reinterpret_cast<double2 *>(Wsm + row * WS + col) = wv;
        if (grow < n_rows) stcg2(Wout + (size_t)grow * ST_P + col, wv);
      }
      pw = warp_sum(pw);
      ww = warp_sum(ww);
      if (lane == 0) {
        kul_add_atomic(sacc + SC_PHP * KUL_STRIDE, pw);
        kul_add_atomic(sacc + SC_HPHP * KUL_STRIDE, ww);
      }
    }
    __syncthreads();
    if (Y) {
      double g0, g1;
      gram_tile(Ysm, Wsm, warp, lane, g0, g1);
      gram_accumulate(g0, g1, inv_q, gfix, &ovf);
    }
    __syncthreads();
  }
  if (Y) gram_flush(set, warp, lane, gfix, ovf);
  flush_scalars(sacc, set, 2);
}

// Gram G = X^T Z over row blocks, fixed point into `set`.
__global__ void __launch_bounds__(TCG_THREADS, 1)
stiefel_gram_kernel(unsigned long long n_rows, const double *X, const double *Z, u64 *set, double inv_q) {
  double *Wsm = reinterpret_cast<double *>(st_smem + SM_W);
  double *Ysm = reinterpret_cast<double *>(st_smem + SM_Y);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long nblk = (n_rows + ST_NB - 1) / ST_NB;
  i64 gfix[4] = {0, 0, 0, 0};
  unsigned ovf = 0;
  for (unsigned long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const unsigned long long r0 = b * ST_NB;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx2 = tid + TCG_THREADS * i;
      const int row = idx2 >> 4, c2 = (idx2 & 15) * 2;
      const unsigned long long grow = r0 + row;
      double2 xv = make_double2(0.0, 0.0), zv = make_double2(0.0, 0.0);
      if (grow < n_rows) {
        const size_t off = (size_t)grow * ST_P + c2;
        xv = ldcg2(X + off);
        zv = ldcg2(Z + off);
      }
      *reinterpret_cast<double2 *>(Ysm + row * WS + c2) = xv;
      *reinterpret_cast<double2 *>(Wsm + row * WS + c2) = zv;
    }
    __syncthreads();
    double g0, g1;
    gram_tile(Ysm, Wsm, warp, lane, g0, g1);
    gram_accumulate(g0, g1, inv_q, gfix, &ovf);
    __syncthreads();
  }
  gram_flush(set, warp, lane, gfix, ovf);
}

// out = cW * W + X * M   (W nullable; M: p x p device, row-major)
__global__ void __launch_bounds__(TCG_THREADS, 1)
stiefel_rowgemm_kernel(unsigned long long n_rows, const double *W, double cW, const double *X,
                       const double *M, double *out) {
  double *Gsm = reinterpret_cast<double *>(st_smem + SM_G);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, m = lane >> 2, j = lane & 3;
  for (int e = tid; e < ST_P * ST_P; e += blockDim.x) Gsm[(e >> 5) * WS + (e & 31)] = M[e];
  __syncthreads();
  const unsigned long long nblk = (n_rows + ST_NB - 1) / ST_NB;
  for (unsigned long long b = blockIdx.x; b < nblk; b += gridDim.x) {
    const unsigned long long grow = b * ST_NB + 8 * warp + m;
    const bool valid = grow < n_rows;
    const size_t rowoff = (size_t)grow * ST_P;
    double acc[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int col = 8 * t + 2 * j;
      if (valid && W) {
        const double2 w = ldcg2(W + rowoff + col);
        acc[t][0] = cW * w.x; acc[t][1] = cW * w.y;
      } else {
        acc[t][0] = acc[t][1] = 0.0;
      }
    }
    strip_rightmul(valid ? X + rowoff : nullptr, Gsm, lane, acc);
    if (valid) {
#pragma unroll
      for (int t = 0; t < 4; ++t) stcg2(out + rowoff + 8 * t + 2 * j, make_double2(acc[t][0], acc[t][1]));
    }
  }
}

// max_i sum_j |A_ij| over all blocks (order independent: max of non-negative doubles as u64 bits)
__global__ void stiefel_absrowsum_kernel(const unsigned short *A, unsigned long long nrows_padded,
                                         unsigned long long *out_bits) {
  const int lane = threadIdx.x & 31;
  const unsigned long long row = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= nrows_padded) return;
  const unsigned short *Arow = A + row * ST_NB;
  double s = 0.0;
  for (int c = lane; c < ST_NB; c += 32) s += fabs(bf16_bits_to_double(Arow[c]));
  s = warp_sum(s);
  if (lane == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(s));
}

// ---- host launchers ---------------------------------------------------------
static bool g_attr_done = false;
static cudaError_t ensure_attrs() {
  if (g_attr_done) return cudaSuccess;
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(tcg_stiefel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V2_TOTAL))) return e;
  if ((e = cudaFuncSetAttribute(stiefel_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL))) return e;
  if ((e = cudaFuncSetAttribute(stiefel_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL))) return e;
  if ((e = cudaFuncSetAttribute(stiefel_rowgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL))) return e;
  g_attr_done = true;
  return cudaSuccess;
}

cudaError_t launch_tcg_stiefel(const TcgCommon &a, unsigned long long n_rows, const unsigned short *A,
                               const double *Y, const double *S_dev, double op_norm_bound, int grid,
                               cudaStream_t stm) {
  cudaError_t e = ensure_attrs();
  if (e) return e;
  TcgCommon ac = a;
  StiefelArgs sa{n_rows, A, Y, S_dev, op_norm_bound};
  void *args[] = {(void *)&ac, (void *)&sa};
  return cudaLaunchCooperativeKernel((const void *)tcg_stiefel_kernel, dim3(grid), dim3(TCG_THREADS), args,
                                     V2_TOTAL, stm);
}
cudaError_t launch_stiefel_apply(unsigned long long n_rows, const unsigned short *A, const double *V,
                                 const double *S_dev, const double *Y, double *Wout, u64 *set, double inv_q,
                                 int grid, cudaStream_t stm) {
  cudaError_t e = ensure_attrs();
  if (e) return e;
  stiefel_apply_kernel<<<grid, TCG_THREADS, SM_TOTAL, stm>>>(n_rows, A, V, S_dev, Y, Wout, set, inv_q);
  return cudaGetLastError();
}
cudaError_t launch_stiefel_gram(unsigned long long n_rows, const double *X, const double *Z, u64 *set,
                                double inv_q, int grid, cudaStream_t stm) {
  cudaError_t e = ensure_attrs();
  if (e) return e;
  stiefel_gram_kernel<<<grid, TCG_THREADS, SM_TOTAL, stm>>>(n_rows, X, Z, set, inv_q);
  return cudaGetLastError();
}
cudaError_t launch_stiefel_rowgemm(unsigned long long n_rows, const double *W, double cW, const double *X,
                                   const double *M_dev, double *out, int grid, cudaStream_t stm) {
  cudaError_t e = ensure_attrs();
  if (e) return e;
  stiefel_rowgemm_kernel<<<grid, TCG_THREADS, SM_TOTAL, stm>>>(n_rows, W, cW, X, M_dev, out);
  return cudaGetLastError();
}
cudaError_t launch_stiefel_absrowsum(const unsigned short *A, unsigned long long nrows_padded,
                                     unsigned long long *out_bits, cudaStream_t stm) {
  const int warps = 8;
  const unsigned long long grid = (nrows_padded + warps - 1) / warps;
  stiefel_absrowsum_kernel<<<(unsigned)grid, warps * 32, 0, stm>>>(A, nrows_padded, out_bits);
  return cudaGetLastError();
}
int gram_exponent_host(double bound) { return gram_exponent(bound); }

}  // namespace ob200
