// Persistent fused truncated-CG for sparse, locally coupled Hessians (BASELINE configs C5 and C4 as tCG operators):
//
//   OB200_OP_BLOCK_CSR3 : rotation synchronisation on SO(3)^N relaxed to St(3, r)^N (SE-Sync's rank-r relaxation):
//                         X in R^{3N x r} row-major, pose i = rows 3i .. 3i+2 with X_i X_i^T = I_3,
//                           Hess f(X)[V] = Proj_X(2 Q V - Lambda V),   Proj_X(Z)_i = Z_i - sym(Z_i X_i^T) X_i,
//                         Q = connection Laplacian, symmetric, 3 x 3 blocks in block-CSR, Lambda_i = sym((2 Q X)_i X_i^T).
//                         The projection is per pose: no global Gram, the only global sums are the CG scalars.
//   OB200_OP_STENCIL7   : 7-point Dirichlet Laplacian on a gx x gy x gz grid (x fastest) applied to the p columns of an
//                         n x p matrix (n = gx gy gz): the operator of config C4 as a Hessian `LinearOperator`.
//
// Same loop, scalar logic and exact reductions as the other kernels (tcg.cuh maps the statements to reference
// IterativeSolvers.h:285-422).  The operator needs p at OTHER rows, so a CG iteration is three streaming phases:
//   phase A1 (l.420 / l.256)            p = -v + beta p (v = M^-1 r with Jacobi), written back; <p,p>, <p,r>
//   grid barrier                        (p complete everywhere)
//   phase A2 (l.294 + l.300 + l.305-306) Hp = H(p): gathers of p from L2, operator data streamed once; <p,Hp>, <Hp,Hp>
//   reduction -> scalar step (decide_after_A)
//   phase B  (l.374 + l.377 + l.383/386 + l.408) as in tcg_diag_kernel
//   reduction -> update_after_B
// Units of deterministic reduction: 256-element runs (A1, B, stencil A2) and groups of 8 (r <= 4) or 4 (r <= 8) poses
// (CSR3 A2), one per warp iteration; partial sums are added exactly (Kulisch accumulators).
// Algorithmic bytes per CG step (e = 8): 11 N e  (A1 reads r, p_old, writes p; A2 reads p, writes Hp; B reads s, p, r,
// Hp, writes s, r)  + B_op,  B_op(CSR3) = nnz (72 + 4) + N_poses (8 + 72 + 3 r e) (row pointer, Lambda, X),
// B_op(STENCIL7) = 0; the gathered neighbour rows of p are L2 traffic, not HBM traffic.
#include "tcg.cuh"

namespace ob200 {

struct SparseArgs {
  int kind;                          // 4 = BLOCK_CSR3, 5 = STENCIL7 (OB200_OP_* values)
  int r;                             // columns
  unsigned long long units;          // CSR3: poses; STENCIL7: grid points
  const unsigned long long *rowptr;  // CSR3
  const unsigned *colidx;
  const double *blocks, *lambda, *X;
  unsigned gx, gy, gz;               // STENCIL7
  // row-sharded CSR3 (one process per GPU): column indices >= units address the halo buffer, which the owning
  // ranks fill with plain stores over NVLink in phase A1 (before the machine-wide barrier)
  unsigned long long n_halo;
  const double *halo;
  const unsigned *send_idx;
  unsigned long long send_ptr[MAX_RANKS + 1];
  double *peer_halo[MAX_RANKS];
};

template <int CNT>
__device__ __forceinline__ void sp_load(const double *base, unsigned long long N, unsigned long long e0, int lane,
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
__device__ __forceinline__ void sp_store(double *base, unsigned long long N, unsigned long long e0, int lane,
                                         const double2 (&v)[CNT]) {
#pragma unroll
  for (int i = 0; i < CNT; ++i) {
    const unsigned long long e = e0 + 2ull * (unsigned)(lane + 32 * i);
    if (e + 1 < N) stcg2(base + e, v[i]);
    else if (e < N) __stcg(base + e, v[i].x);
  }
}

// sum over the lanes of a pose (LPP = 4 or 8 consecutive lanes), xor tree: ((0+1)+(2+3)) [+ ((4+5)+(6+7))]
template <int LPP>
__device__ __forceinline__ double pose_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  if (LPP == 8) v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// Nine consecutive doubles at an 8-byte aligned address (a 3x3 block of the CSR array, Lambda_i) as five 16-byte loads
// from the enclosing 16-byte granules: phase A2 is bound by L1 wavefronts (every load instruction of a warp touches
// one line per pose, eight lines), so five instructions instead of nine is 4/9 fewer wavefronts for these operands.
// The granules lie inside the same pages as the nine doubles, whatever their parity.
__device__ __forceinline__ void load9(const double *p, double (&b)[9]) {
  const unsigned long long addr = (unsigned long long)p;
  const double2 *q = reinterpret_cast<const double2 *>(addr & ~15ull);
  const bool odd = (addr & 8ull) != 0;
  const double2 d0 = __ldg(q), d1 = __ldg(q + 1), d2 = __ldg(q + 2), d3 = __ldg(q + 3), d4 = __ldg(q + 4);
  b[0] = odd ? d0.y : d0.x; b[1] = odd ? d1.x : d0.y; b[2] = odd ? d1.y : d1.x; b[3] = odd ? d2.x : d1.y;
  b[4] = odd ? d2.y : d2.x; b[5] = odd ? d3.x : d2.y; b[6] = odd ? d3.y : d3.x; b[7] = odd ? d4.x : d3.y;
  b[8] = odd ? d4.y : d4.x;
}

// One pose, one column per lane (c < r active): w = (2 Q V - Lambda V)(pose, :, c), then the tangent projection.
// V is read with L2 loads (other CTAs wrote it in the previous phase).  Returns hp[3]; vi[3] = V(pose, :, c).
template <int LPP>
__device__ __forceinline__ void csr3_pose_apply(const SparseArgs &sp, unsigned long long pose, int c, bool active,
                                                const double *V, double (&hp)[3], double (&vi)[3]) {
  const int r = sp.r;
  double z0 = 0.0, z1 = 0.0, z2 = 0.0;
  hp[0] = hp[1] = hp[2] = 0.0;
  vi[0] = vi[1] = vi[2] = 0.0;
  double x0 = 0.0, x1 = 0.0, x2 = 0.0;
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  if (active) {
    const unsigned long long e0 = __ldg(sp.rowptr + pose), e1 = __ldg(sp.rowptr + pose + 1);
    // Edges in chunks of four: the four column indices first, then the twelve gathered values of V (the index -> gather
    // dependency is the latency chain of this phase: one chain per chunk instead of one per edge), then the blocks and the
    // FMAs in edge order (same arithmetic, same order as an edge-by-edge loop).
    for (unsigned long long e = e0; e < e1; e += 4) {
      unsigned long long jc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) jc[q] = (e + q < e1) ? (unsigned long long)__ldg(sp.colidx + e + q) : pose;
      double vv[4][3];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double *Vj = (jc[q] < sp.units ? V + (size_t)3 * jc[q] * r : sp.halo + (size_t)3 * (jc[q] - sp.units) * r) + c;
        vv[q][0] = __ldcg(Vj); vv[q][1] = __ldcg(Vj + r); vv[q][2] = __ldcg(Vj + 2 * r);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (e + q < e1) {
          double b[9];
          load9(sp.blocks + 9 * (e + q), b);
          const double v0 = vv[q][0], v1 = vv[q][1], v2 = vv[q][2];
          z0 = fma(b[0], v0, z0); z0 = fma(b[1], v1, z0); z0 = fma(b[2], v2, z0);
          z1 = fma(b[3], v0, z1); z1 = fma(b[4], v1, z1); z1 = fma(b[5], v2, z1);
          z2 = fma(b[6], v0, z2); z2 = fma(b[7], v1, z2); z2 = fma(b[8], v2, z2);
        }
      }
    }
    const double *Vi = V + (size_t)3 * pose * r + c;
    vi[0] = __ldcg(Vi); vi[1] = __ldcg(Vi + r); vi[2] = __ldcg(Vi + 2 * r);
    double L[9];
    load9(sp.lambda + 9 * pose, L);
    w0 = __dmul_rn(2.0, z0); w1 = __dmul_rn(2.0, z1); w2 = __dmul_rn(2.0, z2);
    w0 = fma(-L[0], vi[0], w0); w0 = fma(-L[1], vi[1], w0); w0 = fma(-L[2], vi[2], w0);
    w1 = fma(-L[3], vi[0], w1); w1 = fma(-L[4], vi[1], w1); w1 = fma(-L[5], vi[2], w1);
    w2 = fma(-L[6], vi[0], w2); w2 = fma(-L[7], vi[1], w2); w2 = fma(-L[8], vi[2], w2);
    const double *Xi = sp.X + (size_t)3 * pose * r + c;
    x0 = __ldg(Xi); x1 = __ldg(Xi + r); x2 = __ldg(Xi + 2 * r);
  }
  // M = W_i X_i^T (3 x 3): sums over the columns = over the lanes of the pose (inactive lanes contribute zeros)
  const double m00 = pose_sum<LPP>(__dmul_rn(w0, x0)), m01 = pose_sum<LPP>(__dmul_rn(w0, x1)), m02 = pose_sum<LPP>(__dmul_rn(w0, x2));
  const double m10 = pose_sum<LPP>(__dmul_rn(w1, x0)), m11 = pose_sum<LPP>(__dmul_rn(w1, x1)), m12 = pose_sum<LPP>(__dmul_rn(w1, x2));
  const double m20 = pose_sum<LPP>(__dmul_rn(w2, x0)), m21 = pose_sum<LPP>(__dmul_rn(w2, x1)), m22 = pose_sum<LPP>(__dmul_rn(w2, x2));
  const double s01 = __dmul_rn(0.5, __dadd_rn(m01, m10)), s02 = __dmul_rn(0.5, __dadd_rn(m02, m20)),
               s12 = __dmul_rn(0.5, __dadd_rn(m12, m21));
  const double s00 = __dmul_rn(0.5, __dadd_rn(m00, m00)), s11 = __dmul_rn(0.5, __dadd_rn(m11, m11)),
               s22 = __dmul_rn(0.5, __dadd_rn(m22, m22));
  if (active) {
    double h = w0; h = fma(-s00, x0, h); h = fma(-s01, x1, h); h = fma(-s02, x2, h); hp[0] = h;
    h = w1; h = fma(-s01, x0, h); h = fma(-s11, x1, h); h = fma(-s12, x2, h); hp[1] = h;
    h = w2; h = fma(-s02, x0, h); h = fma(-s12, x1, h); h = fma(-s22, x2, h); hp[2] = h;
  }
}

// One element of the stencil operator: e = point * p + c
__device__ __forceinline__ double stencil_elem(const SparseArgs &sp, unsigned long long e, unsigned long long N,
                                               const double *V, double &v_self) {
  v_self = 0.0;
  if (e >= N) return 0.0;
  const unsigned p = (unsigned)sp.r;
  const unsigned long long pt = e / p;
  const unsigned x = (unsigned)(pt % sp.gx), y = (unsigned)((pt / sp.gx) % sp.gy), z = (unsigned)(pt / ((unsigned long long)sp.gx * sp.gy));
  const unsigned long long sx = p, sy = (unsigned long long)sp.gx * p, sz = sy * sp.gy;
  v_self = __ldcg(V + e);
  double h = __dmul_rn(6.0, v_self);
  if (x > 0) h = __dsub_rn(h, __ldcg(V + e - sx));
  if (x + 1 < sp.gx) h = __dsub_rn(h, __ldcg(V + e + sx));
  if (y > 0) h = __dsub_rn(h, __ldcg(V + e - sy));
  if (y + 1 < sp.gy) h = __dsub_rn(h, __ldcg(V + e + sy));
  if (z > 0) h = __dsub_rn(h, __ldcg(V + e - sz));
  if (z + 1 < sp.gz) h = __dsub_rn(h, __ldcg(V + e + sz));
  return h;
}

#ifndef OB200_C5_DYN
#define OB200_C5_DYN 1
#endif
// out = H(V) over the units this CTA owns ([it0, it1) warp-iterations, round-robin over the warps), then -- fused kernel
// only -- over warp-iterations [dyn0, dyn1) handed out one at a time through `ticket` (levels the tail: the duration of
// a warp-iteration varies from SM to SM); optional partial sums <V,HV>, <HV,HV>
template <int LPP>
__device__ __forceinline__ void sparse_apply_range(const SparseArgs &sp, unsigned long long N, const double *V, double *out,
                                                   unsigned long long it0, unsigned long long it1, int warp, int lane,
                                                   u64 *sacc /* nullable */, unsigned *ticket = nullptr,
                                                   unsigned long long dyn0 = 0, unsigned long long dyn1 = 0) {
  if (sp.kind == 4) {
    constexpr int PPW = 32 / LPP;                           // poses per warp iteration
    constexpr unsigned long long NONE = ~0ull;
    const int pl = lane / LPP, c = lane % LPP;
    // next warp-iteration of this warp: the static round-robin first, then tickets
    auto successor = [&](unsigned long long it) -> unsigned long long {
      if (it < it1 && it + TCG_WARPS < it1) return it + TCG_WARPS;
      if (!OB200_C5_DYN || ticket == nullptr) return NONE;
      unsigned t = 0;
      if (lane == 0) t = atomicAdd(ticket, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      return dyn0 + t < dyn1 ? dyn0 + t : NONE;
    };
    unsigned long long it = it0 + warp < it1 ? it0 + warp : successor(NONE - TCG_WARPS);
    while (it != NONE) {
      const unsigned long long nxt = successor(it);         // (ticket latency hidden behind this iteration)
      const unsigned long long pose = it * PPW + pl;
      const bool active = pose < sp.units && c < sp.r;
      double hp[3], vi[3];
      csr3_pose_apply<LPP>(sp, pose, c, active, V, hp, vi);
      double php = 0.0, hphp = 0.0;
      if (active) {
        double *o = out + (size_t)3 * pose * sp.r + c;
        __stcg(o, hp[0]); __stcg(o + sp.r, hp[1]); __stcg(o + 2 * sp.r, hp[2]);
        php = fma(vi[0], hp[0], php); php = fma(vi[1], hp[1], php); php = fma(vi[2], hp[2], php);
        hphp = fma(hp[0], hp[0], hphp); hphp = fma(hp[1], hp[1], hphp); hphp = fma(hp[2], hp[2], hphp);
      }
      if (sacc) {
        php = warp_sum(php); hphp = warp_sum(hphp);
        if (lane == 0) kul_add_atomic(sacc + SC_PHP * KUL_STRIDE, php);
        if (lane == 1) kul_add_atomic(sacc + SC_HPHP * KUL_STRIDE, hphp);
      }
      it = nxt;
    }
  } else {
    for (unsigned long long it = it0 + warp; it < it1; it += TCG_WARPS) {
      const unsigned long long e0 = it * 256ull;
      double php = 0.0, hphp = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const unsigned long long e = e0 + (unsigned)(lane + 32 * i);
        double vs;
        const double h = stencil_elem(sp, e, N, V, vs);
        if (e < N) __stcg(out + e, h);
        php = fma(vs, h, php);
        hphp = fma(h, h, hphp);
      }
      if (sacc) {
        php = warp_sum(php); hphp = warp_sum(hphp);
        if (lane == 0) kul_add_atomic(sacc + SC_PHP * KUL_STRIDE, php);
        if (lane == 1) kul_add_atomic(sacc + SC_HPHP * KUL_STRIDE, hphp);
      }
    }
  }
}

__device__ __forceinline__ unsigned long long sparse_iterations(const SparseArgs &sp, unsigned long long N, int lpp) {
  if (sp.kind == 4) { const unsigned ppw = 32 / lpp; return (sp.units + ppw - 1) / ppw; }
  return (N + 255ull) / 256ull;
}

// Two CTAs per SM (64 registers per thread): phase A2 is a chain of dependent gathers (row pointer -> column index ->
// row of p), so its rate is set by the number of warps in flight, not by the width of one warp's loads; the streaming
// phases use runs of SP_RUN elements per warp to stay inside that register budget.
constexpr int SP_CTAS_PER_SM = 2;
constexpr int SP_CH = 2;                                   // double2 per lane and array in the streaming phases
constexpr unsigned long long SP_RUN = 64ull * SP_CH;      // elements per warp iteration there
template <int LPP>
__global__ void __launch_bounds__(TCG_THREADS, SP_CTAS_PER_SM) tcg_sparse_kernel(TcgCommon a, SparseArgs sp) {
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

  const unsigned long long N = a.N;
  const unsigned long long runs = (N + SP_RUN - 1ull) / SP_RUN;
  const unsigned long long u0 = runs * blockIdx.x / gridDim.x, u1 = runs * (blockIdx.x + 1ull) / gridDim.x;
  const unsigned long long its = sparse_iterations(sp, N, LPP);
  // phase A2 of the block-CSR operator: 7/8 of the warp-iterations are owned statically, the rest is handed out by
  // ticket (two counters behind the barrier word, used alternately; CTA 0 clears the idle one in phase A1)
  const unsigned long long its_static = (OB200_C5_DYN && sp.kind == 4) ? its - its / 8 : its;
  const unsigned long long i0 = its_static * blockIdx.x / gridDim.x, i1 = its_static * (blockIdx.x + 1ull) / gridDim.x;
  unsigned *const tickets = a.barrier + 4;
  unsigned gen = 0, phase = 0;
  int exit_reason = -1;

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
    if (blockIdx.x == 0 && threadIdx.x == 0) tickets[(k + 1ull) & 1ull] = 0u;   // idle until phase A2 of iteration k + 1
    for (unsigned long long u = u0 + warp; u < u1; u += TCG_WARPS) {
      const unsigned long long e0 = u * SP_RUN;
      double2 r[SP_CH], po[SP_CH], m[SP_CH], pn[SP_CH];
      sp_load<SP_CH>(a.r, N, e0, lane, r);
      if (k) sp_load<SP_CH>(p_old, N, e0, lane, po);
      if (a.minv) sp_load<SP_CH>(a.minv, N, e0, lane, m);
      double pp = 0.0, pr = 0.0;
#pragma unroll
      for (int i = 0; i < SP_CH; ++i) {
        const double vx = a.minv ? m[i].x * r[i].x : r[i].x;
        const double vy = a.minv ? m[i].y * r[i].y : r[i].y;
        pn[i].x = k ? fma(beta, po[i].x, -vx) : -vx;                        // l.256 / l.420
        pn[i].y = k ? fma(beta, po[i].y, -vy) : -vy;
        pp = fma(pn[i].x, pn[i].x, pp); pp = fma(pn[i].y, pn[i].y, pp);
        pr = fma(pn[i].x, r[i].x, pr);  pr = fma(pn[i].y, r[i].y, pr);
      }
      sp_store<SP_CH>(p_new, N, e0, lane, pn);
      pp = warp_sum(pp); pr = warp_sum(pr);
      if (lane == 0) kul_add_atomic(sacc + SC_PP * KUL_STRIDE, pp);
      if (lane == 1) kul_add_atomic(sacc + SC_PR * KUL_STRIDE, pr);
    }
    if (sp.kind == 4 && a.cm.world > 1) {
      // halo push: the rows of p that other ranks' rows refer to, recomputed from r / p_old (no dependence on other
      // CTAs) and stored straight into the peers' halo buffers over NVLink
      const unsigned long long per = 3ull * sp.r, total = sp.send_ptr[a.cm.world] * per;
      const unsigned long long t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1ull) / gridDim.x;
      for (unsigned long long t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
        const unsigned long long entry = t / per, el = t % per;
        int q = 0;
        while (entry >= sp.send_ptr[q + 1]) ++q;
        const size_t src = (size_t)3 * __ldg(sp.send_idx + entry) * sp.r + el;
        const double rr = __ldcg(a.r + src);
        const double v = a.minv ? __ldg(a.minv + src) * rr : rr;
        const double pv = k ? fma(beta, __ldcg(p_old + src), -v) : -v;
        sp.peer_halo[q][(entry - sp.send_ptr[q]) * per + el] = pv;
      }
      __threadfence_system();
    }
    __syncthreads();
    // <p,p>, <p,r> ride in the set of phase A2 (same reduction); this barrier only orders p
    RedView rvw;
#ifdef OB200_TIMELINE_BUILD
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && sh.k == 3) a.dbg[4096 + 0] = globaltimer_ns();
#endif
    if (!grid_reduce_barrier(a.barrier, gen, a.abort_flag, a.cm, a.cm.epoch + phase, set, 0, 0, rvw)) { exit_reason = -2; break; }
    ++phase;

#ifdef OB200_TIMELINE_BUILD
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && sh.k == 3) a.dbg[4096 + 1] = globaltimer_ns();
#endif
    // ------------------------------ phase A2 ------------------------------
    set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    recycle(phase);
    sparse_apply_range<LPP>(sp, N, p_new, a.Hp, i0, i1, warp, lane, sacc, tickets + (k & 1ull), its_static, its);
    __syncthreads();
    flush_scalars(sacc, set, 4);
#ifdef OB200_TIMELINE_BUILD
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && sh.k == 3) a.dbg[4096 + 2] = globaltimer_ns();
#endif
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
        const unsigned long long e0 = u * SP_RUN;
        double2 s[SP_CH], p[SP_CH];
        sp_load<SP_CH>(a.s, N, e0, lane, s);
        sp_load<SP_CH>(p_new, N, e0, lane, p);
#pragma unroll
        for (int i = 0; i < SP_CH; ++i) { s[i].x = fma(step, p[i].x, s[i].x); s[i].y = fma(step, p[i].y, s[i].y); }
        sp_store<SP_CH>(a.s, N, e0, lane, s);
      }
      exit_reason = sh.action - 1;
      break;
    }

#ifdef OB200_TIMELINE_BUILD
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && sh.k == 3) a.dbg[4096 + 3] = globaltimer_ns();
#endif
    // ------------------------------ phase B ------------------------------
    set = a.acc + (phase % ACC_SETS) * ACC_WORDS;
    recycle(phase);
    for (unsigned long long u = u0 + warp; u < u1; u += TCG_WARPS) {
      const unsigned long long e0 = u * SP_RUN;
      double2 s[SP_CH], p[SP_CH], r[SP_CH], hp[SP_CH], m[SP_CH];
      sp_load<SP_CH>(a.s, N, e0, lane, s);
      sp_load<SP_CH>(p_new, N, e0, lane, p);
      sp_load<SP_CH>(a.r, N, e0, lane, r);
      sp_load<SP_CH>(a.Hp, N, e0, lane, hp);
      if (a.minv) sp_load<SP_CH>(a.minv, N, e0, lane, m);
      double rv = 0.0;
#pragma unroll
      for (int i = 0; i < SP_CH; ++i) {
        s[i].x = fma(step, p[i].x, s[i].x);  s[i].y = fma(step, p[i].y, s[i].y);     // l.374
        r[i].x = fma(step, hp[i].x, r[i].x); r[i].y = fma(step, hp[i].y, r[i].y);    // l.377
        const double vx = a.minv ? m[i].x * r[i].x : r[i].x;                          // l.383/386
        const double vy = a.minv ? m[i].y * r[i].y : r[i].y;
        rv = fma(r[i].x, vx, rv); rv = fma(r[i].y, vy, rv);                           // l.408
      }
      sp_store<SP_CH>(a.s, N, e0, lane, s);
      sp_store<SP_CH>(a.r, N, e0, lane, r);
      rv = warp_sum(rv);
      if (lane == 0) kul_add_atomic(sacc + SC_RV * KUL_STRIDE, rv);
    }
    __syncthreads();
    flush_scalars(sacc + SC_RV * KUL_STRIDE, set + SC_RV * KUL_STRIDE, 1);
#ifdef OB200_TIMELINE_BUILD
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && sh.k == 3) a.dbg[4096 + 4] = globaltimer_ns();
#endif
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

// ---- stand-alone pieces ---------------------------------------------------------------------------------------
template <int LPP>
__global__ void __launch_bounds__(TCG_THREADS, SP_CTAS_PER_SM) sparse_apply_kernel(unsigned long long N, SparseArgs sp, const double *V,
                                                                   double *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long its = sparse_iterations(sp, N, LPP);
  const unsigned long long i0 = its * blockIdx.x / gridDim.x, i1 = its * (blockIdx.x + 1ull) / gridDim.x;
  sparse_apply_range<LPP>(sp, N, V, out, i0, i1, warp, lane, nullptr);
}

// Model of the rotation-synchronisation cost at X: Lambda_i = sym(G_i X_i^T) with G = 2 Q X, optional Riemannian
// gradient G - Lambda X, f = tr(X^T Q X) = 1/2 sum_i tr(G_i X_i^T) accumulated exactly into scalar slot 0 of `set`.
template <int LPP>
__global__ void __launch_bounds__(TCG_THREADS) csr3_model_kernel(SparseArgs sp, const double *X, double *lambda_out,
                                                                 double *grad /* nullable */, u64 *set) {
  __shared__ u64 sacc[KUL_STRIDE];
  for (int i = threadIdx.x; i < KUL_STRIDE; i += blockDim.x) sacc[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int PPW = 32 / LPP;
  const int pl = lane / LPP, c = lane % LPP, r = sp.r;
  const unsigned long long its = (sp.units + PPW - 1) / PPW;
  for (unsigned long long it = (unsigned long long)blockIdx.x * TCG_WARPS + warp; it < its;
       it += (unsigned long long)gridDim.x * TCG_WARPS) {
    const unsigned long long pose = it * PPW + pl;
    const bool active = pose < sp.units && c < r;
    double g0 = 0.0, g1 = 0.0, g2 = 0.0, x0 = 0.0, x1 = 0.0, x2 = 0.0;
    if (active) {
      double z0 = 0.0, z1 = 0.0, z2 = 0.0;
      const unsigned long long e0 = __ldg(sp.rowptr + pose), e1 = __ldg(sp.rowptr + pose + 1);
      for (unsigned long long e = e0; e < e1; ++e) {
        const double *B = sp.blocks + 9 * e;
        const double *Xj = X + (size_t)3 * __ldg(sp.colidx + e) * r + c;
        const double v0 = __ldg(Xj), v1 = __ldg(Xj + r), v2 = __ldg(Xj + 2 * r);
        z0 = fma(__ldg(B + 0), v0, z0); z0 = fma(__ldg(B + 1), v1, z0); z0 = fma(__ldg(B + 2), v2, z0);
        z1 = fma(__ldg(B + 3), v0, z1); z1 = fma(__ldg(B + 4), v1, z1); z1 = fma(__ldg(B + 5), v2, z1);
        z2 = fma(__ldg(B + 6), v0, z2); z2 = fma(__ldg(B + 7), v1, z2); z2 = fma(__ldg(B + 8), v2, z2);
      }
      g0 = __dmul_rn(2.0, z0); g1 = __dmul_rn(2.0, z1); g2 = __dmul_rn(2.0, z2);
      const double *Xi = X + (size_t)3 * pose * r + c;
      x0 = __ldg(Xi); x1 = __ldg(Xi + r); x2 = __ldg(Xi + 2 * r);
    }
    const double m00 = pose_sum<LPP>(__dmul_rn(g0, x0)), m01 = pose_sum<LPP>(__dmul_rn(g0, x1)), m02 = pose_sum<LPP>(__dmul_rn(g0, x2));
    const double m10 = pose_sum<LPP>(__dmul_rn(g1, x0)), m11 = pose_sum<LPP>(__dmul_rn(g1, x1)), m12 = pose_sum<LPP>(__dmul_rn(g1, x2));
    const double m20 = pose_sum<LPP>(__dmul_rn(g2, x0)), m21 = pose_sum<LPP>(__dmul_rn(g2, x1)), m22 = pose_sum<LPP>(__dmul_rn(g2, x2));
    const double s01 = __dmul_rn(0.5, __dadd_rn(m01, m10)), s02 = __dmul_rn(0.5, __dadd_rn(m02, m20)),
                 s12 = __dmul_rn(0.5, __dadd_rn(m12, m21));
    double fpart = 0.0;
    if (active) {
      if (c == 0) {
        double *L = lambda_out + 9 * pose;
        L[0] = m00; L[1] = s01; L[2] = s02; L[3] = s01; L[4] = m11; L[5] = s12; L[6] = s02; L[7] = s12; L[8] = m22;
        fpart = __dmul_rn(0.5, __dadd_rn(__dadd_rn(m00, m11), m22));
      }
      if (grad) {
        double *o = grad + (size_t)3 * pose * r + c;
        double h = g0; h = fma(-m00, x0, h); h = fma(-s01, x1, h); h = fma(-s02, x2, h); o[0] = h;
        h = g1; h = fma(-s01, x0, h); h = fma(-m11, x1, h); h = fma(-s12, x2, h); o[r] = h;
        h = g2; h = fma(-s02, x0, h); h = fma(-s12, x1, h); h = fma(-m22, x2, h); o[2 * r] = h;
      }
    }
    fpart = warp_sum(fpart);
    if (lane == 0) kul_add_atomic(sacc, fpart);
  }
  __syncthreads();
  flush_scalars(sacc, set, 1);
}

// Retraction on St(3, r)^N: out_i = rows of X_i + V_i re-orthonormalised by Gram-Schmidt in row order (the Q factor
// of the QR decomposition of (X_i + V_i)^T with positive diagonal).  One thread per pose.
__global__ void csr3_retract_kernel(unsigned long long N, int r, const double *X, const double *V, double *out, int *bad) {
  const unsigned long long pose = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pose >= N) return;
  double z[3][8];
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 8; ++c) z[a][c] = c < r ? X[(size_t)(3 * pose + a) * r + c] + V[(size_t)(3 * pose + a) * r + c] : 0.0;
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < a; ++b) {
      double d = 0.0;
      for (int c = 0; c < 8; ++c) d = fma(z[a][c], z[b][c], d);
      for (int c = 0; c < 8; ++c) z[a][c] = fma(-d, z[b][c], z[a][c]);
    }
    double nn = 0.0;
    for (int c = 0; c < 8; ++c) nn = fma(z[a][c], z[a][c], nn);
    if (!(nn > 0.0)) { atomicExch(bad, 1); nn = 1.0; }
    const double inv = 1.0 / sqrt(nn);
    for (int c = 0; c < 8; ++c) z[a][c] *= inv;
  }
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < r; ++c) out[(size_t)(3 * pose + a) * r + c] = z[a][c];
}

// ---- host launchers -------------------------------------------------------------------------------------------
static int sparse_grid(unsigned long long work_items, int ctas) {
  unsigned long long g = (work_items + TCG_WARPS - 1) / TCG_WARPS;
  if (g > (unsigned long long)ctas) g = ctas;
  if (g < 1) g = 1;
  return (int)g;
}
cudaError_t launch_tcg_sparse(const TcgCommon &a, const SparseArgs &sp, int sm_count, cudaStream_t st) {
  TcgCommon ac = a;
  SparseArgs sa = sp;
  void *args[] = {(void *)&ac, (void *)&sa};
  const unsigned long long runs = (a.N + SP_RUN - 1ull) / SP_RUN;
  const void *fn = (sp.kind == 4 && sp.r > 4) ? (const void *)tcg_sparse_kernel<8> : (const void *)tcg_sparse_kernel<4>;
  int per_sm = 0;   // co-resident CTAs per SM (cooperative launch: the grid must fit at once)
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, TCG_THREADS, 0);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  if (per_sm > SP_CTAS_PER_SM) per_sm = SP_CTAS_PER_SM;
  const int grid = sparse_grid(runs, per_sm * sm_count);
  return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(TCG_THREADS), args, 0, st);
}
cudaError_t launch_sparse_apply(unsigned long long N, const SparseArgs &sp, const double *V, double *out, int sm_count,
                                cudaStream_t st) {
  const int grid = 2 * sm_count;
  if (sp.kind == 4 && sp.r > 4) sparse_apply_kernel<8><<<grid, TCG_THREADS, 0, st>>>(N, sp, V, out);
  else sparse_apply_kernel<4><<<grid, TCG_THREADS, 0, st>>>(N, sp, V, out);
  return cudaGetLastError();
}
cudaError_t launch_csr3_model(const SparseArgs &sp, const double *X, double *lambda_out, double *grad, u64 *set,
                              int sm_count, cudaStream_t st) {
  const int grid = 2 * sm_count;
  if (sp.r > 4) csr3_model_kernel<8><<<grid, TCG_THREADS, 0, st>>>(sp, X, lambda_out, grad, set);
  else csr3_model_kernel<4><<<grid, TCG_THREADS, 0, st>>>(sp, X, lambda_out, grad, set);
  return cudaGetLastError();
}
cudaError_t launch_csr3_retract(unsigned long long N, int r, const double *X, const double *V, double *out, int *bad,
                                cudaStream_t st) {
  csr3_retract_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(N, r, X, V, out, bad);
  return cudaGetLastError();
}

}  // namespace ob200
