// Persistent fused truncated-CG for elementwise (diagonal) Hessians with an
// optional Jacobi preconditioner, plus the init kernel shared by all operator
// kinds.  See tcg.cuh for the phase structure and the reference lines.
//
// HBM layout: every vector is a flat array of N doubles.  Unit of deterministic
// reduction = run of 256 consecutive elements handled by one warp (lane l owns
// double2 #l, #l+32, #l+64, #l+96 of the run, accumulated in that order, then a
// fixed xor-shuffle tree).  Algorithmic bytes per CG step (e = 8 B):
//   phase A: read r, p_old, d (+minv) ; write p, Hp
//   phase B: read s, p, r, Hp (+minv) ; write s, r          => 10 N e + N e (d)
#include "tcg.cuh"

namespace ob200 {

// A 256-element run is handled by one warp as 4 rows of 32 double2; kernels
// process it in two halves (HALF = 0, 1) of 2 rows to bound register pressure.
template <int CNT>
__device__ __forceinline__ void load_run(const double *base, unsigned long long N,
                                         unsigned long long e0, int lane, double2 (&v)[CNT], int first = 0) {
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    const unsigned long long e = e0 + 2ull * (unsigned)(lane + 32 * (i + first));
    if (e + 1 < N) v[i] = ldcg2(base + e);
    else {
      v[i].x = (e < N) ? __ldcg(base + e) : 0.0;
      v[i].y = 0.0;
    }
  }
}
template <int CNT>
__device__ __forceinline__ void store_run(double *base, unsigned long long N,
                                          unsigned long long e0, int lane, const double2 (&v)[CNT], int first = 0) {
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    const unsigned long long e = e0 + 2ull * (unsigned)(lane + 32 * (i + first));
    if (e + 1 < N) stcg2(base + e, v[i]);
    else if (e < N) __stcg(base + e, v[i].x);
  }
}

// s = 0, r = g, partial of <r, v> with v = minv .* r  (IterativeSolvers.h:211-266)
__global__ void __launch_bounds__(TCG_THREADS) tcg_init_kernel(TcgCommon a) {
  __shared__ u64 sacc[KUL_STRIDE];
  for (int i = threadIdx.x; i < KUL_STRIDE; i += blockDim.x) sacc[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long units = (a.N + 255ull) / 256ull;
  const unsigned long long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1ull) / gridDim.x;
  for (unsigned long long u = u0 + warp; u < u1; u += TCG_WARPS) {
    const unsigned long long e0 = u * 256ull;
    double2 g[4], m[4], z[4];
    load_run<4>(a.g, a.N, e0, lane, g);
    if (a.minv) load_run<4>(a.minv, a.N, e0, lane, m);
    double part = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double vx = a.minv ? m[i].x * g[i].x : g[i].x;
      const double vy = a.minv ? m[i].y * g[i].y : g[i].y;
      part = fma(g[i].x, vx, part);
      part = fma(g[i].y, vy, part);
      z[i].x = 0.0 * g[i].x;   // l.211: s = 0 * g
      z[i].y = 0.0 * g[i].y;
    }
    store_run<4>(a.r, a.N, e0, lane, g);
    store_run<4>(a.s, a.N, e0, lane, z);
    part = warp_sum(part);
    if (lane == 0) kul_add_atomic(sacc, part);
    if (a.blk_stats) {   // max |r| per 128 x 32 block (16 runs of 256 elements), order independent
      double mx = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) mx = fmax(mx, fmax(fabs(g[i].x), fabs(g[i].y)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) atomicMax(a.blk_stats + (u >> 4), (unsigned long long)__double_as_longlong(mx));
    }
  }
  __syncthreads();
  flush_scalars(sacc, a.acc + SC_RV * KUL_STRIDE, 1);
}

// Finalize scalar slot `slot` of accumulator set 0 into result->final_rv (host reads it).
__global__ void tcg_finalize_kernel(const u64 *acc, int slot, double *out) {
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const u64 *p = acc + slot * KUL_STRIDE;
    const double v = kul_finalize_warp([p](int j) { return p[j]; });
    if (threadIdx.x == 0) *out = v;
  }
}

__global__ void __launch_bounds__(TCG_THREADS, 1) tcg_diag_kernel(TcgCommon a, const double *hdiag) {
  __shared__ CgShared sh;
  __shared__ u64 sacc[ACC_NSCAL * KUL_STRIDE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ACC_NSCAL * KUL_STRIDE; i += blockDim.x) sacc[i] = 0;
  if (threadIdx.x == 0) {
    sh.rv = a.rv0;
    sh.sk_M_pk = 0.0;        // l.259
    sh.sk_M_2 = 0.0;         // l.263
    sh.pk_M_2 = a.rv0;       // l.266
    sh.alpha = sh.beta = sh.kappa = sh.step = 0.0;
    sh.k = 0;
    sh.action = ACT_CONTINUE;
    sh.status = 0;
  }
  __syncthreads();

  const unsigned long long units = (a.N + 255ull) / 256ull;
  const unsigned long long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1ull) / gridDim.x;
  unsigned gen = 0;
  unsigned phase = 0;
  int exit_reason = -1;

  for (;;) {
    const unsigned long long k = sh.k;
    if (k >= a.max_iterations) { exit_reason = 1; break; }                  // l.285
    if (sqrt(sh.rv) <= a.target) { exit_reason = 0; break; }                // l.290
    const double beta = sh.beta;
    const double *p_old = (k & 1ull) ? a.p1 : a.p0;
    double *p_new = (k & 1ull) ? a.p0 : a.p1;

    // ------------------------------ phase A ------------------------------
    u64 *set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    if (blockIdx.x == 0) {  // recycle the set used two phases from now
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      for (int i = threadIdx.x; i < ACC_WORDS; i += blockDim.x) nxt[i] = 0;
    }
    for (unsigned long long u = u0 + warp; u < u1; u += TCG_WARPS) {
      const unsigned long long e0 = u * 256ull;
      double php = 0.0, hphp = 0.0, pp = 0.0, pr = 0.0;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double2 r[2], po[2], d[2], m[2], pn[2], hp[2];
        load_run<2>(a.r, a.N, e0, lane, r, 2 * half);
        load_run<2>(hdiag, a.N, e0, lane, d, 2 * half);
        if (a.minv) load_run<2>(a.minv, a.N, e0, lane, m, 2 * half);
        if (k) load_run<2>(p_old, a.N, e0, lane, po, 2 * half);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double vx = a.minv ? m[i].x * r[i].x : r[i].x;
          const double vy = a.minv ? m[i].y * r[i].y : r[i].y;
          pn[i].x = k ? fma(beta, po[i].x, -vx) : -vx;                        // l.256 / l.420
          pn[i].y = k ? fma(beta, po[i].y, -vy) : -vy;
          hp[i].x = d[i].x * pn[i].x;                                         // l.294
          hp[i].y = d[i].y * pn[i].y;
          php = fma(pn[i].x, hp[i].x, php);  php = fma(pn[i].y, hp[i].y, php);
          hphp = fma(hp[i].x, hp[i].x, hphp); hphp = fma(hp[i].y, hp[i].y, hphp);
          pp = fma(pn[i].x, pn[i].x, pp);    pp = fma(pn[i].y, pn[i].y, pp);
          pr = fma(pn[i].x, r[i].x, pr);     pr = fma(pn[i].y, r[i].y, pr);
        }
        store_run<2>(p_new, a.N, e0, lane, pn, 2 * half);
        store_run<2>(a.Hp, a.N, e0, lane, hp, 2 * half);
      }
      php = warp_sum(php); hphp = warp_sum(hphp); pp = warp_sum(pp); pr = warp_sum(pr);
      if (lane == 0) {
        kul_add_atomic(sacc + SC_PHP * KUL_STRIDE, php);
        kul_add_atomic(sacc + SC_HPHP * KUL_STRIDE, hphp);
        kul_add_atomic(sacc + SC_PP * KUL_STRIDE, pp);
        kul_add_atomic(sacc + SC_PR * KUL_STRIDE, pr);
      }
    }
    __syncthreads();
    flush_scalars(sacc, set, 4);
    RedView rvw;
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, 0, 4 * KUL_STRIDE, rvw)) {
      exit_reason = -2;
      break;
    }
    finalize_scalars(rvw, sh, 0, 4);
    __syncthreads();
    if (threadIdx.x == 0)
      decide_after_A(sh, sh.red[SC_PHP], sh.red[SC_HPHP], sh.red[SC_PP], sh.red[SC_PR], a.Delta, a.epsilon);
    __syncthreads();
    ++phase;
    const double step = sh.step;
    if (sh.action != ACT_CONTINUE) {
      // boundary / kernel exit: s += sigma * p   (l.336 / l.360)
      for (unsigned long long u = u0 + warp; u < u1; u += TCG_WARPS) {
        const unsigned long long e0 = u * 256ull;
        double2 s[4], p[4];
        load_run<4>(a.s, a.N, e0, lane, s);
        load_run<4>(p_new, a.N, e0, lane, p);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[i].x = fma(step, p[i].x, s[i].x); s[i].y = fma(step, p[i].y, s[i].y); }
        store_run<4>(a.s, a.N, e0, lane, s);
      }
      exit_reason = sh.action - 1;
      break;
    }

    // ------------------------------ phase B ------------------------------
    set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    if (blockIdx.x == 0) {
      u64 *nxt = a.acc + ((phase + 1) % ACC_SETS) * ACC_WORDS;
      for (int i = threadIdx.x; i < ACC_WORDS; i += blockDim.x) nxt[i] = 0;
    }
    for (unsigned long long u = u0 + warp; u < u1; u += TCG_WARPS) {
      const unsigned long long e0 = u * 256ull;
      double rv = 0.0;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double2 s[2], p[2], r[2], hp[2], m[2];
        load_run<2>(a.s, a.N, e0, lane, s, 2 * half);
        load_run<2>(p_new, a.N, e0, lane, p, 2 * half);
        load_run<2>(a.r, a.N, e0, lane, r, 2 * half);
        load_run<2>(a.Hp, a.N, e0, lane, hp, 2 * half);
        if (a.minv) load_run<2>(a.minv, a.N, e0, lane, m, 2 * half);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          s[i].x = fma(step, p[i].x, s[i].x);  s[i].y = fma(step, p[i].y, s[i].y);     // l.374
          r[i].x = fma(step, hp[i].x, r[i].x); r[i].y = fma(step, hp[i].y, r[i].y);    // l.377
          const double vx = a.minv ? m[i].x * r[i].x : r[i].x;                          // l.383/386
          const double vy = a.minv ? m[i].y * r[i].y : r[i].y;
          rv = fma(r[i].x, vx, rv); rv = fma(r[i].y, vy, rv);                           // l.408
        }
        store_run<2>(a.s, a.N, e0, lane, s, 2 * half);
        store_run<2>(a.r, a.N, e0, lane, r, 2 * half);
      }
      rv = warp_sum(rv);
      if (lane == 0) kul_add_atomic(sacc + SC_RV * KUL_STRIDE, rv);
    }
    __syncthreads();
    flush_scalars(sacc + SC_RV * KUL_STRIDE, set + SC_RV * KUL_STRIDE, 1);
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, SC_RV * KUL_STRIDE,
                             KUL_STRIDE, rvw)) {
      exit_reason = -2;
      break;
    }
    finalize_scalars(rvw, sh, SC_RV, 1);
    __syncthreads();
    if (threadIdx.x == 0) update_after_B(sh, sh.red[SC_RV]);
    __syncthreads();
    ++phase;
  }

  if (blockIdx.x == 0 && threadIdx.x == 0) {
    TcgDeviceResult *res = a.result;
    res->num_iterations = sh.k;
    res->final_rv = sh.rv;
    res->phases = phase;
    if (exit_reason == -2) {
      res->status = 5;  // OB200_ABORTED
      res->exit_reason = -1;
      res->update_step_M_norm = 0.0;
    } else {
      res->status = 0;
      res->exit_reason = exit_reason;
      res->update_step_M_norm = (exit_reason >= 2) ? a.Delta : sqrt(sh.sk_M_2);   // l.334/359/424
    }
  }
}

// Host-driven reductions (init <r,v>, level-1 dots, Grams) across ranks: one CTA
// publishes set[off, off+count) to every rank, waits for all ranks and folds the
// integer total back into `set` in place.  mode 1: element-wise max instead of sum
// (for order-independent maxima stored as non-negative double bit patterns).
__global__ void __launch_bounds__(TCG_THREADS) tcg_exchange_kernel(CommDev cm, unsigned long long gphase, u64 *set,
                                                                   int off, int count, int mode, int *abort_flag) {
  const int slot = (int)(gphase % ACC_SLOTS);
  const size_t slot_off = (size_t)(slot * MAX_RANKS) * cm.words_per_set;
  for (int r = 0; r < cm.world; ++r) {
    u64 *dst = cm.inbox[r] + slot_off + (size_t)cm.rank * cm.words_per_set + off;
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = __ldcg(set + off + i);
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < cm.world) {
    st_release_sys_u64(cm.flags[threadIdx.x] + slot * MAX_RANKS + cm.rank, gphase + 1ull);
    const unsigned long long *f = cm.flags[cm.rank] + slot * MAX_RANKS + threadIdx.x;
    unsigned long long spins = 0;
    while (ld_acquire_sys_u64(f) < gphase + 1ull) {
      if (++spins > (1ull << 26)) { atomicExch(abort_flag, 1); break; }
      if (spins > 256) __nanosleep(64);
    }
    __threadfence();
  }
  __syncthreads();
  const u64 *in = cm.inbox[cm.rank] + slot_off + off;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    u64 v = __ldcg(in + i);
    for (int r = 1; r < cm.world; ++r) {
      const u64 x = __ldcg(in + (size_t)r * cm.words_per_set + i);
      v = mode ? (x > v ? x : v) : v + x;
    }
    set[off + i] = v;
  }
}

// ---- host launchers ---------------------------------------------------------
cudaError_t launch_tcg_exchange(const CommDev &cm, unsigned long long gphase, u64 *set, int off, int count,
                                int mode, int *abort_flag, cudaStream_t st) {
  tcg_exchange_kernel<<<1, TCG_THREADS, 0, st>>>(cm, gphase, set, off, count, mode, abort_flag);
  return cudaGetLastError();
}
cudaError_t launch_tcg_init(const TcgCommon &a, int grid, cudaStream_t st) {
  tcg_init_kernel<<<grid, TCG_THREADS, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_tcg_finalize(const u64 *acc, int slot, double *out, cudaStream_t st) {
  tcg_finalize_kernel<<<1, 32, 0, st>>>(acc, slot, out);
  return cudaGetLastError();
}
cudaError_t launch_tcg_diag(const TcgCommon &a, const double *hdiag, int grid, cudaStream_t st) {
  TcgCommon ac = a;
  const double *hd = hdiag;
  void *args[] = {(void *)&ac, (void *)&hd};
  return cudaLaunchCooperativeKernel((const void *)tcg_diag_kernel, dim3(grid), dim3(TCG_THREADS), args, 0, st);
}
int tcg_diag_max_grid(int sm_count) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tcg_diag_kernel, TCG_THREADS, 0);
  return per_sm > 0 ? sm_count : 0;   // one persistent CTA per SM
}

}  // namespace ob200
