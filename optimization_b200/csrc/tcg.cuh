// Shared definitions for the persistent fused truncated-CG kernels.
//
// One cooperative kernel runs the WHOLE Steihaug-Toint loop of reference
// IterativeSolvers.h:285-422 on the device.  Each CG iteration is two fused
// streaming phases separated by a grid barrier:
//   phase A  (l.420 of the previous iteration + l.294 + l.300 + l.305-306 + l.320)
//            p = -v + beta p ; Hp = H(p) ; partials of <p,Hp>, <Hp,Hp>, <p,p>, <p,r>
//   phase B  (l.374 + l.377 + l.383/386 + l.408)
//            s += alpha p ; r += alpha Hp ; v = P(r) ; partial of <r,v>
// All scalar logic (l.290, 305-362, 408-417, 424) is evaluated redundantly and
// identically by thread 0 of every CTA from exactly-reduced sums.
#pragma once
#include "common.cuh"

namespace ob200 {

constexpr int TCG_THREADS = 512;
constexpr int TCG_WARPS = TCG_THREADS / 32;

// accumulator set layout (u64 words)
constexpr int ACC_NSCAL = 8;                                  // Kulisch scalars per set
constexpr int ACC_SCAL_WORDS = ACC_NSCAL * KUL_STRIDE;        // 576
constexpr int ACC_GRAM_OFF = ACC_SCAL_WORDS;                  // 32x32 entries x {hi,lo}
constexpr int ACC_GRAM_WORDS = 2048;
constexpr int ACC_FLAG_OFF = ACC_GRAM_OFF + ACC_GRAM_WORDS;   // [0] fixed-point overflow
constexpr int ACC_WORDS = ACC_FLAG_OFF + 8;                   // 2632
constexpr int ACC_SETS = ACC_SLOTS;

// scalar slots
enum { SC_PHP = 0, SC_HPHP = 1, SC_PP = 2, SC_PR = 3, SC_RV = 4 };

struct TcgDeviceResult {
  double update_step_M_norm;
  double final_rv;
  double r0_norm;
  unsigned long long num_iterations;
  int exit_reason;
  int status;      // OB200_OK / OB200_NUMERIC_RANGE / OB200_ABORTED
  unsigned phases; // reduction phases consumed (advances the global exchange epoch)
  int pad[1];
};

struct TcgCommon {
  unsigned long long N;        // scalars in a (local) tangent vector
  const double *g;
  double *s, *r, *p0, *p1, *Hp;
  const double *minv;          // Jacobi preconditioner or nullptr
  double rv0, target, Delta, epsilon;
  unsigned long long max_iterations;
  u64 *acc;                    // ACC_SETS * ACC_WORDS
  unsigned *barrier;           // monotonically increasing arrival counter
  int *abort_flag;
  TcgDeviceResult *result;
  CommDev cm;                  // multi-GPU exchange (world == 1: unused)
  unsigned long long *dbg;     // optional: [0..3] ns spent in phase A / A-sync / phase B / B-sync (max over CTAs)
  // optional (Stiefel v6 kernel): per-128-row-block maxima as bit patterns of non-negative doubles,
  // [R0 | R1 : nblk each] max |r| (ping-pong over iterations), [P0 | P1 : 4 nblk each] max |p| per L warp
  unsigned long long *blk_stats;
  unsigned long long nblk_stats;
};


// CTA-shared solver scalars (written by thread 0 only, between barriers)
struct CgShared {
  double rv, sk_M_pk, sk_M_2, pk_M_2;
  double alpha, beta, kappa, step;   // step: alpha (continue) or signed sigma (exit)
  double red[ACC_NSCAL];
  unsigned long long k;
  int action;                         // 0 continue, else exit reason + 1
  int status;
};

enum { ACT_CONTINUE = 0 };

// Flush the CTA-local Kulisch accumulators (shared memory) to the global set.
__device__ __forceinline__ void flush_scalars(u64 *sacc, u64 *gacc, int nscal) {
  for (int i = threadIdx.x; i < nscal * KUL_STRIDE; i += blockDim.x) {
    const u64 v = sacc[i];
    if (v) {
      atomicAdd(gacc + i, v);
      sacc[i] = 0;
    }
  }
}

// After the barrier: warp w < count finalizes scalar first+w into sh.red[] (callers
// __syncthreads() before reading sh.red).
__device__ __forceinline__ void finalize_scalars(const RedView &v, CgShared &sh, int first, int count) {
  const int warp = threadIdx.x >> 5;
  if (warp < count) {
    const int o = (first + warp) * KUL_STRIDE;
    const double x = kul_finalize_warp([&v, o](int j) { return v.load(o + j); });
    if ((threadIdx.x & 31) == 0) sh.red[first + warp] = x;
  }
}

// Scalar logic after phase A.  Reference IterativeSolvers.h:300-362.
// `kappa`, `nHp2`, `np2`, `pr` are the reduced inner products.
__device__ __forceinline__ void decide_after_A(CgShared &sh, double kappa, double nHp2, double np2,
                                               double pr, double Delta, double epsilon) {
  const double Delta_2 = __dmul_rn(Delta, Delta);                       // l.271
  sh.kappa = kappa;                                                     // l.300
  if (__ddiv_rn(sqrt(nHp2), sqrt(np2)) < epsilon) {                     // l.305-307
    double sgn = 1.0;
    if (pr < 0) {                                                       // l.320-326
      sgn = -1.0;
      sh.sk_M_pk = -sh.sk_M_pk;
    }
    const double disc = __dadd_rn(__dmul_rn(sh.sk_M_pk, sh.sk_M_pk),
                                  __dmul_rn(sh.pk_M_2, __dsub_rn(Delta_2, sh.sk_M_2)));
    const double sigma = __ddiv_rn(__dadd_rn(-sh.sk_M_pk, sqrt(disc)), sh.pk_M_2);  // l.330-332
    sh.step = __dmul_rn(sgn, sigma);
    sh.action = 2 /*OB200_EXIT_KERNEL*/ + 1;
    return;
  }
  const double alpha = __ddiv_rn(sh.rv, kappa);                         // l.341
  const double skp1 = __dadd_rn(__dadd_rn(sh.sk_M_2, __dmul_rn(__dmul_rn(2.0, alpha), sh.sk_M_pk)),
                                __dmul_rn(__dmul_rn(alpha, alpha), sh.pk_M_2));  // l.344-345
  if (kappa <= 0 || skp1 > Delta_2) {                                   // l.347
    const double disc = __dadd_rn(__dmul_rn(sh.sk_M_pk, sh.sk_M_pk),
                                  __dmul_rn(sh.pk_M_2, __dsub_rn(Delta_2, sh.sk_M_2)));
    sh.step = __ddiv_rn(__dadd_rn(-sh.sk_M_pk, sqrt(disc)), sh.pk_M_2);  // l.355-357
    sh.action = 3 /*OB200_EXIT_BOUNDARY*/ + 1;
    return;
  }
  sh.alpha = alpha;
  sh.step = alpha;
  sh.sk_M_2 = skp1;  // l.415 (value is consumed only after phase B)
  sh.action = ACT_CONTINUE;
}

// Same decisions with the three long-latency operations (two square roots, one division) already
// evaluated -- by different lanes, concurrently; values and rounding are identical to decide_after_A.
__device__ __forceinline__ void decide_after_A_pre(CgShared &sh, double kappa, double sq_nHp2, double sq_np2,
                                                   double alpha, double pr, double Delta, double epsilon) {
  const double Delta_2 = __dmul_rn(Delta, Delta);                       // l.271
  sh.kappa = kappa;                                                     // l.300
  if (__ddiv_rn(sq_nHp2, sq_np2) < epsilon) {                           // l.305-307
    double sgn = 1.0;
    if (pr < 0) {                                                       // l.320-326
      sgn = -1.0;
      sh.sk_M_pk = -sh.sk_M_pk;
    }
    const double disc = __dadd_rn(__dmul_rn(sh.sk_M_pk, sh.sk_M_pk),
                                  __dmul_rn(sh.pk_M_2, __dsub_rn(Delta_2, sh.sk_M_2)));
    const double sigma = __ddiv_rn(__dadd_rn(-sh.sk_M_pk, sqrt(disc)), sh.pk_M_2);  // l.330-332
    sh.step = __dmul_rn(sgn, sigma);
    sh.action = 2 /*OB200_EXIT_KERNEL*/ + 1;
    return;
  }
  const double skp1 = __dadd_rn(__dadd_rn(sh.sk_M_2, __dmul_rn(__dmul_rn(2.0, alpha), sh.sk_M_pk)),
                                __dmul_rn(__dmul_rn(alpha, alpha), sh.pk_M_2));  // l.344-345
  if (kappa <= 0 || skp1 > Delta_2) {                                   // l.347
    const double disc = __dadd_rn(__dmul_rn(sh.sk_M_pk, sh.sk_M_pk),
                                  __dmul_rn(sh.pk_M_2, __dsub_rn(Delta_2, sh.sk_M_2)));
    sh.step = __ddiv_rn(__dadd_rn(-sh.sk_M_pk, sqrt(disc)), sh.pk_M_2);  // l.355-357
    sh.action = 3 /*OB200_EXIT_BOUNDARY*/ + 1;
    return;
  }
  sh.alpha = alpha;
  sh.step = alpha;
  sh.sk_M_2 = skp1;  // l.415 (value is consumed only after phase B)
  sh.action = ACT_CONTINUE;
}

// Scalar logic after phase B.  Reference IterativeSolvers.h:408-417.
// NOTE: sh.sk_M_2 already holds skplus1_M_2; sk_M_pk / pk_M_2 are updated here.
__device__ __forceinline__ void update_after_B(CgShared &sh, double rk_vk) {
  const double beta = __ddiv_rn(rk_vk, __dmul_rn(sh.alpha, sh.kappa));            // l.412
  sh.sk_M_pk = __dmul_rn(beta, __dadd_rn(sh.sk_M_pk, __dmul_rn(sh.alpha, sh.pk_M_2)));  // l.416
  sh.pk_M_2 = __dadd_rn(rk_vk, __dmul_rn(__dmul_rn(beta, beta), sh.pk_M_2));      // l.417
  sh.beta = beta;
  sh.rv = rk_vk;
  sh.k += 1;
}

}  // namespace ob200
