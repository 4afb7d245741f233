// Persistent fused truncated-CG for the Rayleigh-quotient Hessian on the sphere
// (BASELINE config C2):
//   H v = 2 P_x(A v) - 2 (x^T A x) v ,  P_x(z) = z - x (x^T z) ,
//   A = diag(d) + U diag(sigma) U^T      (n x k low-rank part, k <= 16: a "dense
//   symmetric" action without an explicit n x n matrix)
// plus the stand-alone kernels of the sphere model (A v, Hessian combine).
//
// Same loop, scalar logic and exact reductions as tcg_diag_kernel (tcg.cuh maps the
// statements to reference IterativeSolvers.h:285-422); the low-rank coupling costs one
// extra grid-wide reduction per iteration:
//   phase A1 (l.420 + first half of l.294): p = -v + beta p (written back);
//            partials of t_j = <U_j, p> (k sums), <w, p> (w = A x, precomputed once per
//            base point, so x^T A p = <w, p> needs no second pass), <p,p>, <p,r>
//   reduction 1 -> st_j = sigma_j t_j, c = <w, p>
//   phase A2 (second half of l.294 + l.300 + l.305-306): z = d .* p + sum_j U_j st_j (fma chain in
//            the order of the CPU oracle), Hp = 2 (z - c x) - 2 lambda p (written), partials of
//            <p,Hp>, <Hp,Hp>
//   reduction 2 -> scalar step (decide_after_A)
//   phase B  (l.374 + l.377 + l.383/386 + l.408) as in tcg_diag_kernel
//   reduction 3 -> update_after_B
//
// HBM layout: vectors are flat arrays of n doubles; U is stored TRANSPOSED (k x n row-major:
// column j of U is a contiguous run of n doubles), so every access of the kernel is a
// coalesced 16-byte-per-lane stream.  Unit of deterministic reduction = 256-element run
// handled by one warp.  Four columns of U (8 KB per warp, 128 KB per CTA) are in flight at a time.
// Algorithmic bytes per CG step (e = 8): A1 reads r, p_old, w, U and writes p; A2 reads p, d, x, U
// and writes Hp; B reads s, p, r, Hp and writes s, r  =>  (14 + 2k) n e  (+ 2 n e with Jacobi).
#include "tcg.cuh"

namespace ob200 {

constexpr int SPH_KMAX = 16;
#ifndef SPH_CH_FUSED
#define SPH_CH_FUSED 2   // columns of U in flight per warp inside the persistent kernel (register budget)
#endif
// accumulator slots: the k low-rank sums live in the (otherwise unused) Gram region of the set
enum { SC_WP = 5 };
constexpr int SPH_T_OFF = ACC_GRAM_OFF;                       // t_j at SPH_T_OFF + j * KUL_STRIDE
constexpr int SPH_SACC_WORDS = (ACC_NSCAL + SPH_KMAX) * KUL_STRIDE;
static_assert(SPH_KMAX * KUL_STRIDE <= ACC_GRAM_WORDS, "low-rank sums must fit the Gram region");

struct SphereArgs {
  const double *d;      // n
  const double *Ut;     // k x ldu (row j = column j of U; ldu even so that every row is 16-byte aligned)
  const double *x;      // n, unit vector
  const double *w;      // n, A x
  double sigma[SPH_KMAX];
  double lambda;        // x^T A x
  int k;
  unsigned long long ldu;
};

template <int CNT>
__device__ __forceinline__ void sph_load(const double *base, unsigned long long N, unsigned long long e0, int lane,
                                         double2 (&v)[CNT]) {
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    const unsigned long long e = e0 + 2ull * (unsigned)(lane + 32 * i);
    if (e + 1 < N) v[i] = ldcg2(base + e);
    else {
      v[i].x = (e < N) ? __ldcg(base + e) : 0.0;
      v[i].y = 0.0;
    }
  }
}
template <int CNT>
__device__ __forceinline__ void sph_store(double *base, unsigned long long N, unsigned long long e0, int lane,
                                          const double2 (&v)[CNT]) {
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    const unsigned long long e = e0 + 2ull * (unsigned)(lane + 32 * i);
    if (e + 1 < N) stcg2(base + e, v[i]);
    else if (e < N) __stcg(base + e, v[i].x);
  }
}
// read-only operator data (never written during a solve): streaming (evict-first) loads
template <int CNT>
__device__ __forceinline__ void sph_load_ro(const double *base, unsigned long long N, unsigned long long e0, int lane,
                                            double2 (&v)[CNT]) {
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    const unsigned long long e = e0 + 2ull * (unsigned)(lane + 32 * i);
    if (e + 1 < N) v[i] = __ldcs(reinterpret_cast<const double2 *>(base + e));
    else {
      v[i].x = (e < N) ? __ldcs(base + e) : 0.0;
      v[i].y = 0.0;
    }
  }
}

// Partial sums <U_j, v> of one 256-element run, j = 0 .. k-1, added to the CTA accumulators: the warp
// totals are formed with the fixed xor-shuffle tree and lane j adds total j (exact accumulation).
template <int CH>
__device__ __forceinline__ void sph_run_tdot(const SphereArgs &sp, unsigned long long N, unsigned long long e0, int lane,
                                             const double2 (&v)[4], u64 *sacc_t) {
  double mine = 0.0;
#pragma unroll 1
  for (int j0 = 0; j0 < sp.k; j0 += CH) {
    double2 u[CH][4];
#pragma unroll
    for (int jj = 0; jj < CH; ++jj)
      if (j0 + jj < sp.k) sph_load_ro<4>(sp.Ut + (size_t)(j0 + jj) * sp.ldu, N, e0, lane, u[jj]);
#pragma unroll
    for (int jj = 0; jj < CH; ++jj) {
      if (j0 + jj < sp.k) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { t = fma(u[jj][i].x, v[i].x, t); t = fma(u[jj][i].y, v[i].y, t); }
        t = warp_sum(t);
        if (lane == j0 + jj) mine = t;
      }
    }
  }
  if (lane < sp.k) kul_add_atomic(sacc_t + lane * KUL_STRIDE, mine);
}

// z = d .* v + sum_j U_j st_j for one run (fma chain over j in ascending order, as the CPU oracle)
template <int CH>
__device__ __forceinline__ void sph_run_apply(const SphereArgs &sp, const double *st /* shared */, unsigned long long N,
                                              unsigned long long e0, int lane, const double2 (&v)[4], double2 (&z)[4]) {
  double2 dv[4];
  sph_load_ro<4>(sp.d, N, e0, lane, dv);
#pragma unroll
  for (int i = 0; i < 4; ++i) { z[i].x = dv[i].x * v[i].x; z[i].y = dv[i].y * v[i].y; }
#pragma unroll 1
  for (int j0 = 0; j0 < sp.k; j0 += CH) {
    double2 u[CH][4];
#pragma unroll
    for (int jj = 0; jj < CH; ++jj)
      if (j0 + jj < sp.k) sph_load_ro<4>(sp.Ut + (size_t)(j0 + jj) * sp.ldu, N, e0, lane, u[jj]);
#pragma unroll
    for (int jj = 0; jj < CH; ++jj) {
      if (j0 + jj < sp.k) {
        const double s = st[j0 + jj];
#pragma unroll
        for (int i = 0; i < 4; ++i) { z[i].x = fma(u[jj][i].x, s, z[i].x); z[i].y = fma(u[jj][i].y, s, z[i].y); }
      }
    }
  }
}

__global__ void __launch_bounds__(TCG_THREADS, 1) tcg_sphere_kernel(TcgCommon a, SphereArgs sp) {
  __shared__ CgShared sh;
  __shared__ u64 sacc[SPH_SACC_WORDS];
  __shared__ double s_st[SPH_KMAX];
  __shared__ double s_c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SPH_SACC_WORDS; i += blockDim.x) sacc[i] = 0;
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
  u64 *sacc_t = sacc + ACC_NSCAL * KUL_STRIDE;

  const unsigned long long N = a.N;
  const unsigned long long units = (N + 255ull) / 256ull;
  const unsigned long long u0 = units * blockIdx.x / gridDim.x, u1 = units * (blockIdx.x + 1ull) / gridDim.x;
  unsigned gen = 0;
  unsigned phase = 0;
  int exit_reason = -1;
  const double lam2 = 2.0 * sp.lambda;

  auto recycle = [&](unsigned ph) {   // clear the set used two phases from now (a grid barrier intervenes)
    u64 *nxt = a.acc + ((ph + 1) % ACC_SETS) * ACC_WORDS;
    const int per = (ACC_WORDS + gridDim.x - 1) / gridDim.x;
    const int z0 = per * blockIdx.x;
    for (int i = threadIdx.x; i < per && z0 + i < ACC_WORDS; i += blockDim.x) nxt[z0 + i] = 0;
  };

  for (;;) {
    const unsigned long long k = sh.k;
    if (k >= a.max_iterations) { exit_reason = 1; break; }                  // l.285
    if (sqrt(sh.rv) <= a.target) { exit_reason = 0; break; }                // l.290
    const double beta = sh.beta;
    const double *p_old = (k & 1ull) ? a.p1 : a.p0;
    double *p_new = (k & 1ull) ? a.p0 : a.p1;

    // ------------------------------ phase A1 ------------------------------
    u64 *set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    recycle(phase);
    for (unsigned long long u = u0 + warp; u < u1; u += TCG_WARPS) {
      const unsigned long long e0 = u * 256ull;
      double2 r[4], po[4], m[4], wv[4], pn[4];
      sph_load<4>(a.r, N, e0, lane, r);
      if (k) sph_load<4>(p_old, N, e0, lane, po);
      if (a.minv) sph_load_ro<4>(a.minv, N, e0, lane, m);
      sph_load_ro<4>(sp.w, N, e0, lane, wv);
      double pp = 0.0, pr = 0.0, wp = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double vx = a.minv ? m[i].x * r[i].x : r[i].x;
        const double vy = a.minv ? m[i].y * r[i].y : r[i].y;
        pn[i].x = k ? fma(beta, po[i].x, -vx) : -vx;                        // l.256 / l.420
        pn[i].y = k ? fma(beta, po[i].y, -vy) : -vy;
        pp = fma(pn[i].x, pn[i].x, pp); pp = fma(pn[i].y, pn[i].y, pp);
        pr = fma(pn[i].x, r[i].x, pr);  pr = fma(pn[i].y, r[i].y, pr);
        wp = fma(wv[i].x, pn[i].x, wp); wp = fma(wv[i].y, pn[i].y, wp);
      }
      sph_store<4>(p_new, N, e0, lane, pn);
      pp = warp_sum(pp); pr = warp_sum(pr); wp = warp_sum(wp);
      if (lane == 0) kul_add_atomic(sacc + SC_PP * KUL_STRIDE, pp);
      if (lane == 1) kul_add_atomic(sacc + SC_PR * KUL_STRIDE, pr);
      if (lane == 2) kul_add_atomic(sacc + SC_WP * KUL_STRIDE, wp);
      sph_run_tdot<SPH_CH_FUSED>(sp, N, e0, lane, pn, sacc_t);
    }
    __syncthreads();
    // scalars 2, 3, 5 and the k low-rank sums
    for (int i = threadIdx.x; i < SPH_SACC_WORDS; i += blockDim.x) {
      const u64 v = sacc[i];
      if (v) {
        const int slot = i / KUL_STRIDE;
        atomicAdd(set + (slot < ACC_NSCAL ? i : SPH_T_OFF + (i - ACC_NSCAL * KUL_STRIDE)), v);
        sacc[i] = 0;
      }
    }
    RedView rvw;
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, 0,
                             SPH_T_OFF + SPH_KMAX * KUL_STRIDE, rvw)) { exit_reason = -2; break; }
    {
      // warps 0..k-1: low-rank sums; warps 0..2 also take <p,p>, <p,r>, <w,p> afterwards (k <= 16 warps)
      if (warp < sp.k) {
        const int o = SPH_T_OFF + warp * KUL_STRIDE;
        const double t = kul_finalize_warp([&rvw, o](int j) { return rvw.load(o + j); });
        if (lane == 0) s_st[warp] = __dmul_rn(t, sp.sigma[warp]);
      }
      if (warp < 3) {
        const int slot = warp == 0 ? SC_PP : (warp == 1 ? SC_PR : SC_WP);
        const int o = slot * KUL_STRIDE;
        const double t = kul_finalize_warp([&rvw, o](int j) { return rvw.load(o + j); });
        if (lane == 0) { if (warp == 2) s_c = t; else sh.red[slot] = t; }
      }
    }
    __syncthreads();
    ++phase;
    const double c = s_c;

    // ------------------------------ phase A2 ------------------------------
    set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    recycle(phase);
    for (unsigned long long u = u0 + warp; u < u1; u += TCG_WARPS) {
      const unsigned long long e0 = u * 256ull;
      double2 pn[4], z[4], xv[4], hp[4];
      sph_load<4>(p_new, N, e0, lane, pn);
      sph_run_apply<SPH_CH_FUSED>(sp, s_st, N, e0, lane, pn, z);
      sph_load_ro<4>(sp.x, N, e0, lane, xv);
      double php = 0.0, hphp = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // 2 (A p - (x^T A p) x) - 2 lambda p, every operation rounded separately (the oracle's expression)
        hp[i].x = __dsub_rn(__dmul_rn(2.0, __dsub_rn(z[i].x, __dmul_rn(c, xv[i].x))), __dmul_rn(lam2, pn[i].x));
        hp[i].y = __dsub_rn(__dmul_rn(2.0, __dsub_rn(z[i].y, __dmul_rn(c, xv[i].y))), __dmul_rn(lam2, pn[i].y));
        php = fma(pn[i].x, hp[i].x, php);  php = fma(pn[i].y, hp[i].y, php);
        hphp = fma(hp[i].x, hp[i].x, hphp); hphp = fma(hp[i].y, hp[i].y, hphp);
      }
      sph_store<4>(a.Hp, N, e0, lane, hp);
      php = warp_sum(php); hphp = warp_sum(hphp);
      if (lane == 0) kul_add_atomic(sacc + SC_PHP * KUL_STRIDE, php);
      if (lane == 1) kul_add_atomic(sacc + SC_HPHP * KUL_STRIDE, hphp);
    }
    __syncthreads();
    flush_scalars(sacc, set, 2);
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, 0, 2 * KUL_STRIDE, rvw)) {
      exit_reason = -2;
      break;
    }
    finalize_scalars(rvw, sh, 0, 2);
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
        sph_load<4>(a.s, N, e0, lane, s);
        sph_load<4>(p_new, N, e0, lane, p);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s[i].x = fma(step, p[i].x, s[i].x); s[i].y = fma(step, p[i].y, s[i].y); }
        sph_store<4>(a.s, N, e0, lane, s);
      }
      exit_reason = sh.action - 1;
      break;
    }

    // ------------------------------ phase B ------------------------------
    set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    recycle(phase);
    for (unsigned long long u = u0 + warp; u < u1; u += TCG_WARPS) {
      const unsigned long long e0 = u * 256ull;
      double2 s[4], p[4], r[4], hp[4], m[4];
      sph_load<4>(a.s, N, e0, lane, s);
      sph_load<4>(p_new, N, e0, lane, p);
      sph_load<4>(a.r, N, e0, lane, r);
      sph_load<4>(a.Hp, N, e0, lane, hp);
      if (a.minv) sph_load_ro<4>(a.minv, N, e0, lane, m);
      double rv = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i].x = fma(step, p[i].x, s[i].x);  s[i].y = fma(step, p[i].y, s[i].y);     // l.374
        r[i].x = fma(step, hp[i].x, r[i].x); r[i].y = fma(step, hp[i].y, r[i].y);    // l.377
        const double vx = a.minv ? m[i].x * r[i].x : r[i].x;                          // l.383/386
        const double vy = a.minv ? m[i].y * r[i].y : r[i].y;
        rv = fma(r[i].x, vx, rv); rv = fma(r[i].y, vy, rv);                           // l.408
      }
      sph_store<4>(a.s, N, e0, lane, s);
      sph_store<4>(a.r, N, e0, lane, r);
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

// ---- stand-alone pieces of the sphere model ---------------------------------------
// low-rank sums t_j = <U_j, v> into set[SPH_T_OFF + j * KUL_STRIDE]
__global__ void __launch_bounds__(TCG_THREADS) sphere_tdot_kernel(unsigned long long N, SphereArgs sp, const double *v,
                                                                  u64 *set) {
  __shared__ u64 sacc_t[SPH_KMAX * KUL_STRIDE];
  for (int i = threadIdx.x; i < SPH_KMAX * KUL_STRIDE; i += blockDim.x) sacc_t[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long units = (N + 255ull) / 256ull;
  for (unsigned long long u = (unsigned long long)blockIdx.x * TCG_WARPS + warp; u < units;
       u += (unsigned long long)gridDim.x * TCG_WARPS) {
    const unsigned long long e0 = u * 256ull;
    double2 x[4];
    sph_load<4>(v, N, e0, lane, x);
    sph_run_tdot<4>(sp, N, e0, lane, x, sacc_t);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < sp.k * KUL_STRIDE; i += blockDim.x) {
    const u64 w = sacc_t[i];
    if (w) atomicAdd(set + SPH_T_OFF + i, w);
  }
}
// st_j = sigma_j * t_j  (one warp per j)
__global__ void sphere_scale_t_kernel(const u64 *set, SphereArgs sp, double *st) {
  const int warp = threadIdx.x >> 5;
  if (warp < sp.k) {
    const u64 *p = set + SPH_T_OFF + warp * KUL_STRIDE;
    const double t = kul_finalize_warp([p](int j) { return p[j]; });
    if ((threadIdx.x & 31) == 0) st[warp] = __dmul_rn(t, sp.sigma[warp]);
  }
}
// out = A v = d .* v + sum_j U_j st_j
__global__ void __launch_bounds__(TCG_THREADS) sphere_apply_kernel(unsigned long long N, SphereArgs sp, const double *st_dev,
                                                                   const double *v, double *out) {
  __shared__ double s_st[SPH_KMAX];
  if ((int)threadIdx.x < sp.k) s_st[threadIdx.x] = st_dev[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long units = (N + 255ull) / 256ull;
  for (unsigned long long u = (unsigned long long)blockIdx.x * TCG_WARPS + warp; u < units;
       u += (unsigned long long)gridDim.x * TCG_WARPS) {
    const unsigned long long e0 = u * 256ull;
    double2 x[4], z[4];
    sph_load<4>(v, N, e0, lane, x);
    sph_run_apply<4>(sp, s_st, N, e0, lane, x, z);
    sph_store<4>(out, N, e0, lane, z);
  }
}
// out = 2 (Av - c x) - 2 lambda v   (c = x^T A v; with v = x and c = lambda: the Riemannian gradient)
__global__ void __launch_bounds__(TCG_THREADS) sphere_combine_kernel(unsigned long long N, const double *Av, const double *x,
                                                                     const double *v, double c, double lambda, double *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long units = (N + 255ull) / 256ull;
  const double lam2 = 2.0 * lambda;
  for (unsigned long long u = (unsigned long long)blockIdx.x * TCG_WARPS + warp; u < units;
       u += (unsigned long long)gridDim.x * TCG_WARPS) {
    const unsigned long long e0 = u * 256ull;
    double2 a[4], xv[4], vv[4], o[4];
    sph_load<4>(Av, N, e0, lane, a);
    sph_load<4>(x, N, e0, lane, xv);
    if (v) sph_load<4>(v, N, e0, lane, vv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i].x = __dmul_rn(2.0, __dsub_rn(a[i].x, __dmul_rn(c, xv[i].x)));
      o[i].y = __dmul_rn(2.0, __dsub_rn(a[i].y, __dmul_rn(c, xv[i].y)));
      if (v) {
        o[i].x = __dsub_rn(o[i].x, __dmul_rn(lam2, vv[i].x));
        o[i].y = __dsub_rn(o[i].y, __dmul_rn(lam2, vv[i].y));
      }
    }
    sph_store<4>(out, N, e0, lane, o);
  }
}

// ---- host launchers -------------------------------------------------------------------
static int sph_grid(unsigned long long N, int sm_count) {
  const unsigned long long units = (N + 255ull) / 256ull;
  unsigned long long g = (units + TCG_WARPS - 1) / TCG_WARPS;
  const unsigned long long cap = (unsigned long long)sm_count * 2ull;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
static SphereArgs make_args(const double *d, const double *Ut, unsigned long long ldu, const double *x, const double *w,
                            const double *sigma, int k, double lambda) {
  SphereArgs sp;
  sp.d = d; sp.Ut = Ut; sp.ldu = ldu; sp.x = x; sp.w = w; sp.k = k; sp.lambda = lambda;
  for (int j = 0; j < SPH_KMAX; ++j) sp.sigma[j] = j < k ? sigma[j] : 0.0;
  return sp;
}
cudaError_t launch_tcg_sphere(const TcgCommon &a, const double *d, const double *Ut, unsigned long long ldu, const double *x,
                              const double *w, const double *sigma_host, int k, double lambda, int grid, cudaStream_t st) {
  TcgCommon ac = a;
  SphereArgs sp = make_args(d, Ut, ldu, x, w, sigma_host, k, lambda);
  void *args[] = {(void *)&ac, (void *)&sp};
  return cudaLaunchCooperativeKernel((const void *)tcg_sphere_kernel, dim3(grid), dim3(TCG_THREADS), args, 0, st);
}
cudaError_t launch_sphere_tdot(unsigned long long N, const double *Ut, unsigned long long ldu, const double *sigma_host, int k,
                               const double *v, u64 *set, int sm_count, cudaStream_t st) {
  SphereArgs sp = make_args(nullptr, Ut, ldu, nullptr, nullptr, sigma_host, k, 0.0);
  sphere_tdot_kernel<<<sph_grid(N, sm_count), TCG_THREADS, 0, st>>>(N, sp, v, set);
  return cudaGetLastError();
}
cudaError_t launch_sphere_scale_t(const u64 *set, const double *sigma_host, int k, double *st_dev, cudaStream_t st) {
  SphereArgs sp = make_args(nullptr, nullptr, 0, nullptr, nullptr, sigma_host, k, 0.0);
  sphere_scale_t_kernel<<<1, 32 * SPH_KMAX, 0, st>>>(set, sp, st_dev);
  return cudaGetLastError();
}
cudaError_t launch_sphere_apply(unsigned long long N, const double *d, const double *Ut, unsigned long long ldu, int k,
                                const double *st_dev, const double *v, double *out, int sm_count, cudaStream_t st) {
  double zero[SPH_KMAX] = {0};
  SphereArgs sp = make_args(d, Ut, ldu, nullptr, nullptr, zero, k, 0.0);
  sphere_apply_kernel<<<sph_grid(N, sm_count), TCG_THREADS, 0, st>>>(N, sp, st_dev, v, out);
  return cudaGetLastError();
}
cudaError_t launch_sphere_combine(unsigned long long N, const double *Av, const double *x, const double *v, double c,
                                  double lambda, double *out, int sm_count, cudaStream_t st) {
  sphere_combine_kernel<<<sph_grid(N, sm_count), TCG_THREADS, 0, st>>>(N, Av, x, v, c, lambda, out);
  return cudaGetLastError();
}

}  // namespace ob200
