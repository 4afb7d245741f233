// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// oracle/_ref/libref_oracle.so: the UNMODIFIED reference headers
//   /root/reference/include/Optimization/LinearAlgebra/IterativeSolvers.h (STPCG, l.166-426)
//   /root/reference/include/Optimization/Riemannian/TNT.h              (TNT,   l.242-689)
//   /root/reference/include/Optimization/Riemannian/GradientDescent.h  (l.124-398)
// included where they lie (never copied) and instantiated with oracle::HostMat
// (hostmat.hpp; Eigen is absent from this image).  The functors below define
// the synthetic problems of SURVEY.md section 8(d); the reference itself ships
// no Stiefel / Rayleigh code, only the S^2 test problem
// (tests/TNT_unit_test.cpp:63-122), which is restated literally here.
//
// Build: see oracle/Makefile (g++ -std=c++17 -O3 -march=x86-64-v3
// -ffp-contract=off -fopenmp -I/root/reference/include).  Outputs go to
// oracle/_ref/ only.  Exposed as a plain C ABI for ctypes.
#include "Optimization/LinearAlgebra/IterativeSolvers.h"
#include "Optimization/Riemannian/GradientDescent.h"
#include "Optimization/Riemannian/TNLS.h"
#include "Optimization/Riemannian/TNT.h"

#include "hostmat.hpp"
#include "dense_lu.hpp"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <memory>

using oracle::HostMat;
using namespace Optimization;

namespace {

inline double bf16_to_double(uint16_t b) {
  uint32_t u = uint32_t(b) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return double(f);
}

// ---------------------------------------------------------------------------
// Block-diagonal dense symmetric operator A (blocks of nb x nb, stored bf16).
// ---------------------------------------------------------------------------
struct BlockDiag {
  size_t n = 0, p = 0, nb = 0, nblk = 0;
  std::vector<double> A; // nblk * nb * nb, row-major per block, zero padded

  // out = A * V   (V, out: n x p row-major)
  void apply(const double *V, double *out) const {
    const size_t P = p;
    auto body = [&](size_t b) {
      const double *Ab = &A[b * nb * nb];
      const size_t r0 = b * nb;
      const size_t rows = std::min(nb, n - r0);
      double acc[64];
      for (size_t i = 0; i < rows; ++i) {
        for (size_t j = 0; j < P; ++j) acc[j] = 0.0;
        const double *Ai = Ab + i * nb;
        for (size_t k = 0; k < rows; ++k) {
          const double a = Ai[k];
          const double *Vk = V + (r0 + k) * P;
          for (size_t j = 0; j < P; ++j) acc[j] = std::fma(a, Vk[j], acc[j]);
        }
        double *o = out + (r0 + i) * P;
        for (size_t j = 0; j < P; ++j) o[j] = acc[j];
      }
    };
    int T = oracle::threads();
    if (T <= 1) {
      for (size_t b = 0; b < nblk; ++b) body(b);
    } else {
#pragma omp parallel for schedule(static) num_threads(T)
      for (long b = 0; b < long(nblk); ++b) body(size_t(b));
    }
  }
};

// G = X^T Z  (p x p), X, Z: n x p row-major.  Thread partials combined in order.
void gram(const double *X, const double *Z, size_t n, size_t p, double *G) {
  int T = std::max(1, oracle::threads());
  if (n * p < 65536) T = 1;
  std::vector<double> part(size_t(T) * p * p, 0.0);
  auto body = [&](int t, size_t lo, size_t hi) {
    double *g = &part[size_t(t) * p * p];
    for (size_t r = lo; r < hi; ++r) {
      const double *x = X + r * p, *z = Z + r * p;
      for (size_t i = 0; i < p; ++i) {
        const double xi = x[i];
        double *gi = g + i * p;
        for (size_t j = 0; j < p; ++j) gi[j] = std::fma(xi, z[j], gi[j]);
      }
    }
  };
  if (T == 1)
    body(0, 0, n);
  else {
#pragma omp parallel for schedule(static) num_threads(T)
    for (int t = 0; t < T; ++t)
      body(t, n * size_t(t) / size_t(T), n * size_t(t + 1) / size_t(T));
  }
  for (size_t e = 0; e < p * p; ++e) {
    double s = 0;
    for (int t = 0; t < T; ++t) s += part[size_t(t) * p * p + e];
    G[e] = s;
  }
}

void symmetrize(double *G, size_t p) {
  for (size_t i = 0; i < p; ++i)
    for (size_t j = i; j < p; ++j) {
      double s = 0.5 * (G[i * p + j] + G[j * p + i]);
      G[i * p + j] = s;
      G[j * p + i] = s;
    }
}

// out = W - X * M   (X: n x p, M: p x p), row-parallel
void sub_right_mul(const double *W, const double *X, const double *M, size_t n,
                   size_t p, double *out) {
  oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
    double acc[64];
    for (size_t r = lo; r < hi; ++r) {
      const double *x = X + r * p;
      for (size_t j = 0; j < p; ++j) acc[j] = 0.0;
      for (size_t k = 0; k < p; ++k) {
        const double xk = x[k];
        const double *Mk = M + k * p;
        for (size_t j = 0; j < p; ++j) acc[j] = std::fma(xk, Mk[j], acc[j]);
      }
      const double *w = W + r * p;
      double *o = out + r * p;
      for (size_t j = 0; j < p; ++j) o[j] = w[j] - acc[j];
    }
  });
}

// ---------------------------------------------------------------------------
// Stiefel trace minimisation  f(Y) = 1/2 tr(Y^T A Y)  on St(n, p)
//   grad f(Y)    = A Y - Y sym(Y^T A Y)
//   Hess f(Y)[V] = P_Y(A V - V sym(Y^T A Y)),  P_Y(Z) = Z - Y sym(Y^T Z)
// The user cache (the reference's `Args&...` channel, TNT.h:249) carries
// S = sym(Y^T A Y), written by QM and read by the Hessian functor -- the same
// pattern SE-Sync uses for its cached Euclidean gradient.
// ---------------------------------------------------------------------------
struct StiefelCache {
  const BlockDiag *op = nullptr;
  std::vector<double> S; // p x p
};

void stiefel_S(const BlockDiag &op, const HostMat &Y, std::vector<double> &S,
               HostMat *AY_out = nullptr) {
  HostMat AY(Y.size());
  op.apply(Y.data(), AY.data());
  S.assign(op.p * op.p, 0.0);
  gram(Y.data(), AY.data(), op.n, op.p, S.data());
  symmetrize(S.data(), op.p);
  if (AY_out) *AY_out = std::move(AY);
}

HostMat stiefel_hess(const BlockDiag &op, const HostMat &Y,
                     const std::vector<double> &S, const HostMat &V) {
  const size_t n = op.n, p = op.p;
  HostMat W(V.size());
  op.apply(V.data(), W.data());
  sub_right_mul(W.data(), V.data(), S.data(), n, p, W.data()); // W = AV - V S
  std::vector<double> G(p * p);
  gram(Y.data(), W.data(), n, p, G.data());
  symmetrize(G.data(), p);
  HostMat out(V.size());
  sub_right_mul(W.data(), Y.data(), G.data(), n, p, out.data());
  return out;
}

// Cholesky-QR retraction: Z = Y + V, Z^T Z = R^T R, Q = Z R^{-1}
HostMat stiefel_retract(size_t n, size_t p, const HostMat &Y, const HostMat &V) {
  HostMat Z(Y.size());
  {
    double *z = Z.data();
    const double *y = Y.data(), *v = V.data();
    oracle::parallel_ranges(Z.size(), [&](int, size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i) z[i] = y[i] + v[i];
    });
  }
  std::vector<double> M(p * p), R(p * p, 0.0);
  gram(Z.data(), Z.data(), n, p, M.data());
  symmetrize(M.data(), p);
  // upper Cholesky, M = R^T R
  for (size_t j = 0; j < p; ++j) {
    double s = M[j * p + j];
    for (size_t k = 0; k < j; ++k) s -= R[k * p + j] * R[k * p + j];
    const double rjj = std::sqrt(s);
    R[j * p + j] = rjj;
    for (size_t i = j + 1; i < p; ++i) {
      double t = M[j * p + i];
      for (size_t k = 0; k < j; ++k) t -= R[k * p + j] * R[k * p + i];
      R[j * p + i] = t / rjj;
    }
  }
  // Q = Z R^{-1}: q_j = (z_j - sum_{k<j} q_k R[k][j]) / R[j][j]
  double *z = Z.data();
  oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
    for (size_t r = lo; r < hi; ++r) {
      double *q = z + r * p;
      for (size_t j = 0; j < p; ++j) {
        double t = q[j];
        for (size_t k = 0; k < j; ++k) t -= q[k] * R[k * p + j];
        q[j] = t / R[j * p + j];
      }
    }
  });
  return Z;
}

// ---------------------------------------------------------------------------
// Sphere Rayleigh quotient  f(x) = x^T A x,  A = diag(d) + U diag(sigma) U^T
// ---------------------------------------------------------------------------
struct SphereOp {
  size_t n = 0, k = 0;
  const double *d = nullptr, *U = nullptr, *sigma = nullptr; // U: n x k row-major
  void apply(const double *v, double *out) const {
    std::vector<double> t(k, 0.0);
    if (k) {
      // t = U^T v  (as a 1-column gram, sequential/threads-ordered)
      int T = std::max(1, oracle::threads());
      if (n < 65536) T = 1;
      std::vector<double> part(size_t(T) * k, 0.0);
      auto body = [&](int th, size_t lo, size_t hi) {
        double *pt = &part[size_t(th) * k];
        for (size_t r = lo; r < hi; ++r) {
          const double vr = v[r];
          const double *u = U + r * k;
          for (size_t j = 0; j < k; ++j) pt[j] = std::fma(u[j], vr, pt[j]);
        }
      };
      if (T == 1)
        body(0, 0, n);
      else {
#pragma omp parallel for schedule(static) num_threads(T)
        for (int th = 0; th < T; ++th)
          body(th, n * size_t(th) / size_t(T), n * size_t(th + 1) / size_t(T));
      }
      for (size_t j = 0; j < k; ++j) {
        double s = 0;
        for (int th = 0; th < T; ++th) s += part[size_t(th) * k + j];
        t[j] = s * sigma[j];
      }
    }
    oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
      for (size_t r = lo; r < hi; ++r) {
        double acc = d[r] * v[r];
        const double *u = U + r * k;
        for (size_t j = 0; j < k; ++j) acc = std::fma(u[j], t[j], acc);
        out[r] = acc;
      }
    });
  }
};

struct SphereCache {
  double xAx = 0;
};

template <class Result>
void export_tnt(const Result &res, int *status, uint64_t *n_outer,
                uint64_t *n_trace, double *scalars, uint64_t cap,
                uint64_t *inner_iterations, double *radius, double *rho,
                double *fvals, double *gradnorms, double *step_norms,
                double *step_M_norms) {
  *status = int(res.status);
  *n_outer = res.inner_iterations.size();
  *n_trace = res.trust_region_radius.size();
  scalars[0] = res.f;
  scalars[1] = res.gradfx_norm;
  scalars[2] = res.preconditioned_grad_f_x_norm;
  scalars[3] = res.elapsed_time;
  for (size_t i = 0; i < res.inner_iterations.size() && i < cap; ++i) {
    inner_iterations[i] = res.inner_iterations[i];
    rho[i] = res.gain_ratios[i];
    step_norms[i] = res.update_step_norms[i];
    step_M_norms[i] = res.update_step_M_norms[i];
  }
  for (size_t i = 0; i < res.trust_region_radius.size() && i < cap; ++i) {
    radius[i] = res.trust_region_radius[i];
    fvals[i] = res.objective_values[i];
    gradnorms[i] = res.gradient_norms[i];
  }
}

using NoP = std::optional<LinearAlgebra::STPCGPreconditioner<HostMat, std::nullptr_t>>;
using NoAt = std::optional<LinearAlgebra::LinearOperator<std::nullptr_t, HostMat>>;
using NoUser = std::optional<LinearAlgebra::STPCGUserFunction<HostMat, std::nullptr_t, double>>;

Riemannian::TNTParams<double> make_params(const double *prm) {
  // prm: [max_iterations, gradient_tolerance, relative_decrease_tolerance,
  //       stepsize_tolerance, preconditioned_gradient_tolerance,
  //       Delta_tolerance, Delta0, eta1, eta2, alpha1, alpha2,
  //       max_TPCG_iterations, kappa_fgr, theta, max_computation_time]
  Riemannian::TNTParams<double> P;
  P.max_iterations = size_t(prm[0]);
  P.gradient_tolerance = prm[1];
  P.relative_decrease_tolerance = prm[2];
  P.stepsize_tolerance = prm[3];
  P.preconditioned_gradient_tolerance = prm[4];
  P.Delta_tolerance = prm[5];
  P.Delta0 = prm[6];
  P.eta1 = prm[7];
  P.eta2 = prm[8];
  P.alpha1 = prm[9];
  P.alpha2 = prm[10];
  P.max_TPCG_iterations = size_t(prm[11]);
  P.kappa_fgr = prm[12];
  P.theta = prm[13];
  P.max_computation_time = prm[14];
  return P;
}

} // namespace

// ---- STPCG over an operator given as a plain C callback (sparse Hessian families of configs C5 / C4) -------------
// The reference's STPCG (IterativeSolvers.h:166-426) with HostMat vectors; the Hessian functor calls the shared
// C restatement of the operator (oracle/sparse_ops.h), the preconditioner is the optional pointwise Jacobi scaling.
extern "C" {
#include "sparse_ops.h"
}
namespace {
template <class Apply>
int ref_stpcg_callback(uint64_t n, const double *g, const double *minv, Apply apply, double Delta,
                       uint64_t max_iterations, double kappa_fgr, double theta, double epsilon, double *s_out,
                       double *update_step_M_norm, uint64_t *num_iterations) {
  using V = HostMat;
  using M = std::nullptr_t;
  V G(g, n);
  LinearAlgebra::SymmetricLinearOperator<V> H = [&](const V &x) {
    V out(n);
    apply(x.data(), out.data());
    return out;
  };
  LinearAlgebra::InnerProduct<V> ip = [](const V &a, const V &b) { return oracle::dot(a, b); };
  std::optional<LinearAlgebra::STPCGPreconditioner<V, M>> P;
  if (minv)
    P = [&](const V &x) {
      V out(n);
      const double *xd = x.data();
      double *o = out.data();
      for (size_t i = 0; i < n; ++i) o[i] = minv[i] * xd[i];
      return std::make_pair(std::move(out), M());
    };
  try {
    size_t iters = 0;
    double mnorm = 0;
    V s = LinearAlgebra::STPCG<V, M>(G, H, ip, mnorm, iters, Delta, size_t(max_iterations), kappa_fgr, theta, P,
                                     NoAt(), NoUser(), epsilon);
    std::memcpy(s_out, s.data(), n * sizeof(double));
    *update_step_M_norm = mnorm;
    *num_iterations = iters;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}
}  // namespace

extern "C" {

void ref_set_threads(int t) { oracle::threads() = t < 1 ? 1 : t; }
int ref_get_threads() { return oracle::threads(); }
int ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- STPCG, diagonal Hessian, optional Jacobi preconditioner --------------
// Mirrors the shape of tests/IterativeSolvers_unit_test.cpp:138-310.
// returns 0 on success, 1 if the reference threw std::invalid_argument.
int ref_stpcg_diag(uint64_t n, const double *g, const double *hdiag,
                   const double *minv /* nullable */, double Delta,
                   uint64_t max_iterations, double kappa_fgr, double theta,
                   double epsilon, double *s_out, double *update_step_M_norm,
                   uint64_t *num_iterations) {
  using V = HostMat;
  using M = std::nullptr_t;
  V G(g, n);
  LinearAlgebra::SymmetricLinearOperator<V> H = [&](const V &x) {
    V out(n);
    const double *xd = x.data();
    double *o = out.data();
    oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i) o[i] = hdiag[i] * xd[i];
    });
    return out;
  };
  LinearAlgebra::InnerProduct<V> ip = [](const V &a, const V &b) {
    return oracle::dot(a, b);
  };
  std::optional<LinearAlgebra::STPCGPreconditioner<V, M>> P;
  if (minv)
    P = [&](const V &x) {
      V out(n);
      const double *xd = x.data();
      double *o = out.data();
      oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) o[i] = minv[i] * xd[i];
      });
      return std::make_pair(std::move(out), M());
    };
  try {
    size_t iters = 0;
    double mnorm = 0;
    V s = LinearAlgebra::STPCG<V, M>(G, H, ip, mnorm, iters, Delta,
                                     size_t(max_iterations), kappa_fgr, theta,
                                     P, NoAt(), NoUser(), epsilon);
    std::memcpy(s_out, s.data(), n * sizeof(double));
    *update_step_M_norm = mnorm;
    *num_iterations = iters;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

int ref_csr3_stpcg(uint64_t N, uint64_t r, const uint64_t *rowptr, const uint32_t *colidx,
                              const double *blocks, const double *lambda, const double *X, const double *g,
                              const double *minv, double Delta, uint64_t max_iterations, double kappa_fgr,
                              double theta, double epsilon, double *s_out, double *mnorm, uint64_t *iters) {
  return ref_stpcg_callback(3 * N * r, g, minv, [&](const double *v, double *out) {
    csr3_hess_apply(N, int(r), rowptr, colidx, blocks, lambda, X, v, out);
  }, Delta, max_iterations, kappa_fgr, theta, epsilon, s_out, mnorm, iters);
}
int ref_stencil7_stpcg(uint32_t gx, uint32_t gy, uint32_t gz, uint64_t p, const double *g,
                                  const double *minv, double Delta, uint64_t max_iterations, double kappa_fgr,
                                  double theta, double epsilon, double *s_out, double *mnorm, uint64_t *iters) {
  return ref_stpcg_callback(uint64_t(gx) * gy * gz * p, g, minv, [&](const double *v, double *out) {
    stencil7_apply(gx, gy, gz, int(p), v, out);
  }, Delta, max_iterations, kappa_fgr, theta, epsilon, s_out, mnorm, iters);
}

// ---- Stiefel problem handle -------------------------------------------------
struct RefStiefel {
  BlockDiag op;
};

void *ref_stiefel_create(uint64_t n, uint64_t p, uint64_t nb,
                         const uint16_t *A_bf16 /* nblk*nb*nb */) {
  auto *h = new RefStiefel;
  h->op.n = n;
  h->op.p = p;
  h->op.nb = nb;
  h->op.nblk = (n + nb - 1) / nb;
  h->op.A.resize(h->op.nblk * nb * nb);
  for (size_t i = 0; i < h->op.A.size(); ++i) h->op.A[i] = bf16_to_double(A_bf16[i]);
  return h;
}
void ref_stiefel_destroy(void *h) { delete static_cast<RefStiefel *>(h); }

// S = sym(Y^T A Y); also returns f = 1/2 <Y, AY> and grad = AY - Y S
void ref_stiefel_model(void *hh, const double *Y, double *S_out, double *f_out,
                       double *grad_out /* nullable */) {
  auto *h = static_cast<RefStiefel *>(hh);
  const size_t N = h->op.n * h->op.p;
  HostMat Ym(Y, N), AY;
  std::vector<double> S;
  stiefel_S(h->op, Ym, S, &AY);
  std::memcpy(S_out, S.data(), S.size() * sizeof(double));
  if (f_out) *f_out = 0.5 * oracle::dot(Ym, AY);
  if (grad_out)
    sub_right_mul(AY.data(), Ym.data(), S.data(), h->op.n, h->op.p, grad_out);
}

void ref_stiefel_hess(void *hh, const double *Y, const double *S,
                      const double *V, double *out) {
  auto *h = static_cast<RefStiefel *>(hh);
  const size_t N = h->op.n * h->op.p;
  HostMat Ym(Y, N), Vm(V, N);
  std::vector<double> Sv(S, S + h->op.p * h->op.p);
  HostMat r = stiefel_hess(h->op, Ym, Sv, Vm);
  std::memcpy(out, r.data(), N * sizeof(double));
}

void ref_stiefel_retract(void *hh, const double *Y, const double *V, double *out) {
  auto *h = static_cast<RefStiefel *>(hh);
  const size_t N = h->op.n * h->op.p;
  HostMat Ym(Y, N), Vm(V, N);
  HostMat r = stiefel_retract(h->op.n, h->op.p, Ym, Vm);
  std::memcpy(out, r.data(), N * sizeof(double));
}

// Stand-alone tCG on the Stiefel Hessian at Y:
// reference STPCG with H = Hess f(Y), Frobenius inner product, optional Jacobi.
// projected != 0: P(x) = P_Y(minv o x), P_Y(Z) = Z - Y sym(Y^T Z) (what TNT's adapter hands to STPCG for the
// tangent-space preserving preconditioner of ref_stiefel_tnt_precon).
int ref_stiefel_stpcg_precon(void *hh, const double *Y, const double *g,
                             const double *minv /* nullable, n*p */, int projected, double Delta,
                             uint64_t max_iterations, double kappa_fgr, double theta,
                             double epsilon, double *s_out, double *update_step_M_norm,
                             uint64_t *num_iterations) {
  auto *h = static_cast<RefStiefel *>(hh);
  using V = HostMat;
  using M = std::nullptr_t;
  const size_t N = h->op.n * h->op.p;
  V Ym(Y, N), G(g, N);
  std::vector<double> S;
  stiefel_S(h->op, Ym, S);
  LinearAlgebra::SymmetricLinearOperator<V> H = [&](const V &x) {
    return stiefel_hess(h->op, Ym, S, x);
  };
  LinearAlgebra::InnerProduct<V> ip = [](const V &a, const V &b) {
    return oracle::dot(a, b);
  };
  std::optional<LinearAlgebra::STPCGPreconditioner<V, M>> P;
  if (minv)
    P = [&](const V &x) {
      V out(N);
      const double *xd = x.data();
      double *o = out.data();
      oracle::parallel_ranges(N, [&](int, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) o[i] = minv[i] * xd[i];
      });
      if (projected) {
        const size_t n = h->op.n, p = h->op.p;
        std::vector<double> Gm(p * p);
        gram(Ym.data(), out.data(), n, p, Gm.data());
        symmetrize(Gm.data(), p);
        V proj(N);
        sub_right_mul(out.data(), Ym.data(), Gm.data(), n, p, proj.data());
        out = std::move(proj);
      }
      return std::make_pair(std::move(out), M());
    };
  try {
    size_t iters = 0;
    double mnorm = 0;
    V s = LinearAlgebra::STPCG<V, M>(G, H, ip, mnorm, iters, Delta,
                                     size_t(max_iterations), kappa_fgr, theta,
                                     P, NoAt(), NoUser(), epsilon);
    std::memcpy(s_out, s.data(), N * sizeof(double));
    *update_step_M_norm = mnorm;
    *num_iterations = iters;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

int ref_stiefel_stpcg(void *hh, const double *Y, const double *g,
                      const double *minv /* nullable, n*p */, double Delta,
                      uint64_t max_iterations, double kappa_fgr, double theta,
                      double epsilon, double *s_out, double *update_step_M_norm,
                      uint64_t *num_iterations) {
  return ref_stiefel_stpcg_precon(hh, Y, g, minv, 0, Delta, max_iterations, kappa_fgr, theta, epsilon, s_out,
                                  update_step_M_norm, num_iterations);
}

// End-to-end reference TNT on the Stiefel trace-min problem.
// minv (nullable): elementwise scaling of the tangent-space preserving preconditioner  precon(Y, V) = P_Y(minv o V),
// P_Y(Z) = Z - Y sym(Y^T Z), handed to the reference's TNT as its `precon` argument (TNT.h:247, adapter l.413-426).
int ref_stiefel_tnt_precon(void *hh, const double *Y0, const double *minv, const double *prm, double *Y_out,
                           int *status, uint64_t *n_outer, uint64_t *n_trace,
                           double *scalars, uint64_t cap, uint64_t *inner_iterations,
                           double *radius, double *rho, double *fvals, double *gradnorms,
                           double *step_norms, double *step_M_norms) {
  auto *h = static_cast<RefStiefel *>(hh);
  using V = HostMat;
  const BlockDiag &op = h->op;
  const size_t N = op.n * op.p;
  StiefelCache cache;
  cache.op = &op;

  Objective<V, double, StiefelCache> f = [&](const V &Y, StiefelCache &) {
    V AY(N);
    op.apply(Y.data(), AY.data());
    return 0.5 * oracle::dot(Y, AY);
  };
  Riemannian::QuadraticModel<V, V, StiefelCache> QM =
      [&](const V &Y, V &grad, Riemannian::LinearOperator<V, V, StiefelCache> &Hess,
          StiefelCache &c) {
        V AY;
        stiefel_S(op, Y, c.S, &AY);
        grad = V(N);
        sub_right_mul(AY.data(), Y.data(), c.S.data(), op.n, op.p, grad.data());
        Hess = [&op](const V &Yc, const V &Vt, StiefelCache &cc) {
          return stiefel_hess(op, Yc, cc.S, Vt);
        };
      };
  Riemannian::RiemannianMetric<V, V, double, StiefelCache> metric =
      [](const V &, const V &a, const V &b, StiefelCache &) {
        return oracle::dot(a, b);
      };
  Riemannian::Retraction<V, V, StiefelCache> retract =
      [&](const V &Y, const V &Vt, StiefelCache &) {
        return stiefel_retract(op.n, op.p, Y, Vt);
      };
  std::optional<Riemannian::LinearOperator<V, V, StiefelCache>> precon;
  if (minv)
    precon = [&op, minv, N](const V &Y, const V &Vt, StiefelCache &) {
      V Z(N);
      double *z = Z.data();
      const double *v = Vt.data();
      oracle::parallel_ranges(N, [&](int, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) z[i] = minv[i] * v[i];
      });
      std::vector<double> G(op.p * op.p);
      gram(Y.data(), Z.data(), op.n, op.p, G.data());
      symmetrize(G.data(), op.p);
      V out(N);
      sub_right_mul(Z.data(), Y.data(), G.data(), op.n, op.p, out.data());
      return out;
    };
  try {
    V Y0m(Y0, N);
    auto res = Riemannian::TNT<V, V, double, StiefelCache>(
        f, QM, metric, retract, Y0m, cache, precon, make_params(prm));
    std::memcpy(Y_out, res.x.data(), N * sizeof(double));
    export_tnt(res, status, n_outer, n_trace, scalars, cap, inner_iterations,
               radius, rho, fvals, gradnorms, step_norms, step_M_norms);
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

int ref_stiefel_tnt(void *hh, const double *Y0, const double *prm, double *Y_out,
                    int *status, uint64_t *n_outer, uint64_t *n_trace,
                    double *scalars, uint64_t cap, uint64_t *inner_iterations,
                    double *radius, double *rho, double *fvals, double *gradnorms,
                    double *step_norms, double *step_M_norms) {
  return ref_stiefel_tnt_precon(hh, Y0, nullptr, prm, Y_out, status, n_outer, n_trace, scalars, cap, inner_iterations,
                                radius, rho, fvals, gradnorms, step_norms, step_M_norms);
}

// ---- Sphere Rayleigh quotient (configs C1 / C2) ----------------------------
int ref_sphere_stpcg(uint64_t n, uint64_t k, const double *d, const double *U,
                     const double *sigma, const double *x, const double *g,
                     double Delta, uint64_t max_iterations, double kappa_fgr,
                     double theta, double epsilon, double *s_out,
                     double *update_step_M_norm, uint64_t *num_iterations) {
  using V = HostMat;
  using M = std::nullptr_t;
  SphereOp op{n, k, d, U, sigma};
  V X(x, n), G(g, n);
  V AX(n);
  op.apply(X.data(), AX.data());
  const double xAx = oracle::dot(X, AX);
  LinearAlgebra::SymmetricLinearOperator<V> H = [&](const V &v) {
    V Av(n);
    op.apply(v.data(), Av.data());
    const double xAv = oracle::dot(X, Av);
    V out(n);
    double *o = out.data();
    const double *av = Av.data(), *xd = X.data(), *vd = v.data();
    oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i)
        o[i] = 2.0 * (av[i] - xAv * xd[i]) - 2.0 * xAx * vd[i];
    });
    return out;
  };
  LinearAlgebra::InnerProduct<V> ip = [](const V &a, const V &b) {
    return oracle::dot(a, b);
  };
  try {
    size_t iters = 0;
    double mnorm = 0;
    V s = LinearAlgebra::STPCG<V, M>(G, H, ip, mnorm, iters, Delta,
                                     size_t(max_iterations), kappa_fgr, theta,
                                     NoP(), NoAt(), NoUser(), epsilon);
    std::memcpy(s_out, s.data(), n * sizeof(double));
    *update_step_M_norm = mnorm;
    *num_iterations = iters;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

// minv (nullable): pointwise Jacobi scaling  precon(x, v) = minv o v  as the reference TNT's `precon` argument
int ref_sphere_tnt_precon(uint64_t n, uint64_t k, const double *d, const double *U,
                   const double *sigma, const double *x0, const double *minv, const double *prm,
                   double *x_out, int *status, uint64_t *n_outer,
                   uint64_t *n_trace, double *scalars, uint64_t cap,
                   uint64_t *inner_iterations, double *radius, double *rho,
                   double *fvals, double *gradnorms, double *step_norms,
                   double *step_M_norms) {
  using V = HostMat;
  SphereOp op{n, k, d, U, sigma};
  SphereCache cache;
  Objective<V, double, SphereCache> f = [&](const V &x, SphereCache &) {
    V Ax(n);
    op.apply(x.data(), Ax.data());
    return oracle::dot(x, Ax);
  };
  Riemannian::QuadraticModel<V, V, SphereCache> QM =
      [&](const V &x, V &grad, Riemannian::LinearOperator<V, V, SphereCache> &Hess,
          SphereCache &c) {
        V Ax(n);
        op.apply(x.data(), Ax.data());
        c.xAx = oracle::dot(x, Ax);
        grad = V(n);
        double *gd = grad.data();
        const double *ax = Ax.data(), *xd = x.data();
        const double xAx = c.xAx;
        oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
          for (size_t i = lo; i < hi; ++i) gd[i] = 2.0 * (ax[i] - xAx * xd[i]);
        });
        Hess = [&op, n](const V &xc, const V &v, SphereCache &cc) {
          V Av(n);
          op.apply(v.data(), Av.data());
          const double xAv = oracle::dot(xc, Av);
          V out(n);
          double *o = out.data();
          const double *av = Av.data(), *xd2 = xc.data(), *vd = v.data();
          const double c2 = cc.xAx;
          oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i)
              o[i] = 2.0 * (av[i] - xAv * xd2[i]) - 2.0 * c2 * vd[i];
          });
          return out;
        };
      };
  Riemannian::RiemannianMetric<V, V, double, SphereCache> metric =
      [](const V &, const V &a, const V &b, SphereCache &) {
        return oracle::dot(a, b);
      };
  Riemannian::Retraction<V, V, SphereCache> retract =
      [n](const V &x, const V &v, SphereCache &) {
        V z(n);
        double *zd = z.data();
        const double *xd = x.data(), *vd = v.data();
        oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
          for (size_t i = lo; i < hi; ++i) zd[i] = xd[i] + vd[i];
        });
        const double nrm = std::sqrt(oracle::dot(z, z));
        const double inv = 1.0 / nrm;
        oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
          for (size_t i = lo; i < hi; ++i) zd[i] *= inv;
        });
        return z;
      };
  std::optional<Riemannian::LinearOperator<V, V, SphereCache>> precon;
  if (minv)
    precon = [minv, n](const V &, const V &v, SphereCache &) {
      V out(n);
      double *o = out.data();
      const double *vd = v.data();
      oracle::parallel_ranges(n, [&](int, size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) o[i] = minv[i] * vd[i];
      });
      return out;
    };
  try {
    V X0(x0, n);
    auto res = Riemannian::TNT<V, V, double, SphereCache>(
        f, QM, metric, retract, X0, cache, precon, make_params(prm));
    std::memcpy(x_out, res.x.data(), n * sizeof(double));
    export_tnt(res, status, n_outer, n_trace, scalars, cap, inner_iterations,
               radius, rho, fvals, gradnorms, step_norms, step_M_norms);
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

int ref_sphere_tnt(uint64_t n, uint64_t k, const double *d, const double *U,
                   const double *sigma, const double *x0, const double *prm,
                   double *x_out, int *status, uint64_t *n_outer,
                   uint64_t *n_trace, double *scalars, uint64_t cap,
                   uint64_t *inner_iterations, double *radius, double *rho,
                   double *fvals, double *gradnorms, double *step_norms,
                   double *step_M_norms) {
  return ref_sphere_tnt_precon(n, k, d, U, sigma, x0, nullptr, prm, x_out, status, n_outer, n_trace, scalars, cap,
                               inner_iterations, radius, rho, fvals, gradnorms, step_norms, step_M_norms);
}

// ---- The reference's own S^2 test problem (tests/TNT_unit_test.cpp:63-122) --
// f(X;P) = |X - P|^2 on S^2, P = north pole; projection retraction; optional
// preconditioner diag(1,2,3).  The Hessian keeps the reference's quirk
// (X.dot(gradF(X,P)) with the *Riemannian* gradient, TNT_unit_test.cpp:96).
int ref_s2_tnt(const double *x0, const double *Ppt, int use_precon,
               const double *prm, double *x_out, int *status, uint64_t *n_outer,
               uint64_t *n_trace, double *scalars, uint64_t cap,
               uint64_t *inner_iterations, double *radius, double *rho,
               double *fvals, double *gradnorms, double *step_norms,
               double *step_M_norms) {
  using V = HostMat;
  auto project = [](const V &X, const V &W) {
    V out(3);
    const double c = oracle::dot(X, W);
    for (int i = 0; i < 3; ++i) out.d[i] = W.d[i] - c * X.d[i];
    return out;
  };
  Objective<V, double, V> F = [](const V &X, V &P) {
    double s = 0;
    for (int i = 0; i < 3; ++i) s += (X.d[i] - P.d[i]) * (X.d[i] - P.d[i]);
    return s;
  };
  Riemannian::VectorField<V, V, V> gradF = [project](const V &X, V &P) {
    V nabla(3);
    for (int i = 0; i < 3; ++i) nabla.d[i] = 2 * (X.d[i] - P.d[i]);
    return project(X, nabla);
  };
  Riemannian::LinearOperatorConstructor<V, V, V> HessCon =
      [project, gradF](const V &, V &) {
        Riemannian::LinearOperator<V, V, V> Hessian =
            [project, gradF](const V &X, const V &Xdot, V &P) {
              V EH(3);
              for (int i = 0; i < 3; ++i) EH.d[i] = 2.0 * Xdot.d[i];
              V out = project(X, EH);
              const double c = oracle::dot(X, gradF(X, P));
              for (int i = 0; i < 3; ++i) out.d[i] -= c * Xdot.d[i];
              return out;
            };
        return Hessian;
      };
  Riemannian::RiemannianMetric<V, V, double, V> metric =
      [](const V &, const V &a, const V &b, V &) { return oracle::dot(a, b); };
  Riemannian::Retraction<V, V, V> retract = [](const V &X, const V &Vt, V &) {
    V z(3);
    for (int i = 0; i < 3; ++i) z.d[i] = X.d[i] + Vt.d[i];
    const double nrm = std::sqrt(oracle::dot(z, z));
    for (int i = 0; i < 3; ++i) z.d[i] /= nrm;
    return z;
  };
  std::optional<Riemannian::LinearOperator<V, V, V>> precon;
  if (use_precon)
    precon = [](const V &, const V &Vt, V &) {
      V out(3);
      out.d[0] = 1.0 * Vt.d[0];
      out.d[1] = 2.0 * Vt.d[1];
      out.d[2] = 3.0 * Vt.d[2];
      return out;
    };
  try {
    V X0(x0, 3), P(Ppt, 3);
    auto res = Riemannian::TNT<V, V, double, V>(F, gradF, HessCon, metric, retract,
                                               X0, P, precon, make_params(prm));
    std::memcpy(x_out, res.x.data(), 3 * sizeof(double));
    export_tnt(res, status, n_outer, n_trace, scalars, cap, inner_iterations,
               radius, rho, fvals, gradnorms, step_norms, step_M_norms);
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

// Reference GradientDescent on the same S^2 problem
// (tests/GradientDescent_unit_test.cpp:76-148 shape).
int ref_s2_gd(const double *x0, const double *Ppt, uint64_t max_iterations,
              double gradient_tolerance, double *x_out, int *status,
              uint64_t *n_iter, double *f_out, double *gradnorm_out) {
  using V = HostMat;
  auto project = [](const V &X, const V &W) {
    V out(3);
    const double c = oracle::dot(X, W);
    for (int i = 0; i < 3; ++i) out.d[i] = W.d[i] - c * X.d[i];
    return out;
  };
  Objective<V, double, V> F = [](const V &X, V &P) {
    double s = 0;
    for (int i = 0; i < 3; ++i) s += (X.d[i] - P.d[i]) * (X.d[i] - P.d[i]);
    return s;
  };
  Riemannian::VectorField<V, V, V> gradF = [project](const V &X, V &P) {
    V nabla(3);
    for (int i = 0; i < 3; ++i) nabla.d[i] = 2 * (X.d[i] - P.d[i]);
    return project(X, nabla);
  };
  Riemannian::RiemannianMetric<V, V, double, V> metric =
      [](const V &, const V &a, const V &b, V &) { return oracle::dot(a, b); };
  Riemannian::Retraction<V, V, V> retract = [](const V &X, const V &Vt, V &) {
    V z(3);
    for (int i = 0; i < 3; ++i) z.d[i] = X.d[i] + Vt.d[i];
    const double nrm = std::sqrt(oracle::dot(z, z));
    for (int i = 0; i < 3; ++i) z.d[i] /= nrm;
    return z;
  };
  Riemannian::GradientDescentParams<double> params;
  params.max_iterations = size_t(max_iterations);
  params.gradient_tolerance = gradient_tolerance;
  try {
    V X0(x0, 3), P(Ppt, 3);
    auto res = Riemannian::GradientDescent<V, V, double, V>(F, gradF, metric,
                                                           retract, X0, P, params);
    std::memcpy(x_out, res.x.data(), 3 * sizeof(double));
    *status = int(res.status);
    *n_iter = res.gradient_norms.size();
    *f_out = res.f;
    *gradnorm_out = res.gradfx_norm;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

// Reference GradientDescent on the sphere Rayleigh-quotient model (same functors as ref_sphere_tnt).
int ref_sphere_gd(uint64_t n, uint64_t k, const double *d, const double *U,
                  const double *sigma, const double *x0, uint64_t max_iterations,
                  double gradient_tolerance, double *x_out, int *status,
                  uint64_t *n_iter, uint64_t *ls_total, double *f_out,
                  double *gradnorm_out) {
  using V = HostMat;
  SphereOp op{n, k, d, U, sigma};
  Objective<V, double> f = [&](const V &x) {
    V Ax(n);
    op.apply(x.data(), Ax.data());
    return oracle::dot(x, Ax);
  };
  Riemannian::VectorField<V, V> grad = [&](const V &x) {
    V Ax(n);
    op.apply(x.data(), Ax.data());
    const double xAx = oracle::dot(x, Ax);
    V g(n);
    double *gd = g.data();
    const double *ax = Ax.data(), *xd = x.data();
    for (size_t i = 0; i < n; ++i) gd[i] = 2.0 * (ax[i] - xAx * xd[i]);
    return g;
  };
  Riemannian::RiemannianMetric<V, V, double> metric =
      [](const V &, const V &a, const V &b) { return oracle::dot(a, b); };
  Riemannian::Retraction<V, V> retract = [n](const V &x, const V &v) {
    V z(n);
    double *zd = z.data();
    const double *xd = x.data(), *vd = v.data();
    for (size_t i = 0; i < n; ++i) zd[i] = xd[i] + vd[i];
    const double inv = 1.0 / std::sqrt(oracle::dot(z, z));
    for (size_t i = 0; i < n; ++i) zd[i] *= inv;
    return z;
  };
  Riemannian::GradientDescentParams<double> params;
  params.max_iterations = size_t(max_iterations);
  params.gradient_tolerance = gradient_tolerance;
  try {
    V X0(x0, n);
    auto res = Riemannian::GradientDescent<V, V, double>(f, grad, metric, retract, X0, params);
    std::memcpy(x_out, res.x.data(), n * sizeof(double));
    *status = int(res.status);
    *n_iter = res.gradient_norms.size();
    uint64_t ls = 0;
    for (size_t v : res.linesearch_iterations) ls += v;
    *ls_total = ls;
    *f_out = res.f;
    *gradnorm_out = res.gradfx_norm;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

// Constraint-preconditioned (projected) STPCG, the shape of the reference's tests
// tests/IterativeSolvers_unit_test.cpp:316-496: H = diag(h), M = diag(m), constraint A s = 0
// (A: mc x n row-major), preconditioner = KKT solve [M A^T; A 0][v; lambda] = [r; 0] (dense LU stand-in
// for UmfPackLU), At(lambda) = A^T lambda.
int ref_stpcg_projected(uint64_t n, uint64_t mc, const double *hdiag, const double *mdiag,
                        const double *A, const double *g, double Delta,
                        uint64_t max_iterations, double kappa_fgr, double theta,
                        double *s_out, double *update_step_M_norm, uint64_t *num_iterations) {
  using V = HostMat;
  oracle::DenseLU lu;
  if (!lu.factor(oracle::kkt_matrix(mdiag, A, n, mc), n + mc)) return 2;
  LinearAlgebra::SymmetricLinearOperator<V> H = [&](const V &v) {
    V out(n);
    for (size_t i = 0; i < n; ++i) out.d[i] = hdiag[i] * v.d[i];
    return out;
  };
  LinearAlgebra::InnerProduct<V> ip = [](const V &a, const V &b) { return oracle::dot(a, b); };
  LinearAlgebra::STPCGPreconditioner<V, V> P = [&](const V &r) {
    std::vector<double> w(n + mc, 0.0);
    for (size_t i = 0; i < n; ++i) w[i] = r.d[i];
    const std::vector<double> z = lu.solve(w);
    V x(n), l(mc);
    for (size_t i = 0; i < n; ++i) x.d[i] = z[i];
    for (size_t c = 0; c < mc; ++c) l.d[c] = z[n + c];
    return std::make_pair(x, l);
  };
  LinearAlgebra::LinearOperator<V, V> At = [&](const V &l) {
    V out(n);
    for (size_t i = 0; i < n; ++i) {
      double acc = 0;
      for (size_t c = 0; c < mc; ++c) acc += A[c * n + i] * l.d[c];
      out.d[i] = acc;
    }
    return out;
  };
  try {
    V G(g, n);
    size_t iters = 0;
    double mnorm = 0;
    std::optional<LinearAlgebra::STPCGPreconditioner<V, V>> Pop(P);
    std::optional<LinearAlgebra::LinearOperator<V, V>> Atop(At);
    V s = LinearAlgebra::STPCG<V, V>(G, H, ip, mnorm, iters, Delta, size_t(max_iterations),
                                     kappa_fgr, theta, Pop, Atop);
    std::memcpy(s_out, s.data(), n * sizeof(double));
    *update_step_M_norm = mnorm;
    *num_iterations = iters;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

// Reference LSQR (IterativeSolvers.h:552-855) on a dense m x n matrix (row-major), the shape of the
// reference's tests/IterativeSolvers_unit_test.cpp:498-700.
int ref_lsqr(uint64_t m, uint64_t n, const double *A, const double *b, uint64_t max_iterations,
             double lambda, double btol, double Atol, double cond_limit, double Delta,
             double *x_out, double *xnorm_out, uint64_t *num_iterations) {
  using V = HostMat;
  LinearAlgebra::LinearOperator<V, V> Aop = [&](const V &x) {
    V out(m);
    for (size_t i = 0; i < m; ++i) {
      double acc = 0;
      for (size_t j = 0; j < n; ++j) acc += A[i * n + j] * x.d[j];
      out.d[i] = acc;
    }
    return out;
  };
  LinearAlgebra::LinearOperator<V, V> Atop = [&](const V &y) {
    V out(n);
    for (size_t j = 0; j < n; ++j) {
      double acc = 0;
      for (size_t i = 0; i < m; ++i) acc += A[i * n + j] * y.d[i];
      out.d[j] = acc;
    }
    return out;
  };
  LinearAlgebra::InnerProduct<V> ip = [](const V &a, const V &c) { return oracle::dot(a, c); };
  try {
    V B(b, m);
    double xnorm = 0;
    size_t iters = 0;
    V x = LinearAlgebra::LSQR<V>(Aop, Atop, B, ip, xnorm, iters, size_t(max_iterations), lambda, btol,
                                 Atol, cond_limit, Delta);
    if (x.size() == n) std::memcpy(x_out, x.data(), n * sizeof(double));
    else std::memset(x_out, 0, n * sizeof(double));
    *xnorm_out = xnorm;
    *num_iterations = iters;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

// Reference LSQR on a DIAGONAL operator A = diag(d) (the shape the device check uses: A and A^T are pointwise scalings,
// ob200_hadamard on the device), n unknowns.
int ref_lsqr_diag(uint64_t n, const double *d, const double *b, uint64_t max_iterations,
                  double lambda, double btol, double Atol, double cond_limit, double Delta,
                  double *x_out, double *xnorm_out, uint64_t *num_iterations) {
  using V = HostMat;
  LinearAlgebra::LinearOperator<V, V> Aop = [&](const V &x) {
    V out(n);
    for (size_t i = 0; i < n; ++i) out.d[i] = d[i] * x.d[i];
    return out;
  };
  LinearAlgebra::InnerProduct<V> ip = [](const V &a, const V &c) { return oracle::dot(a, c); };
  try {
    V B(b, n);
    double xnorm = 0;
    size_t iters = 0;
    V x = LinearAlgebra::LSQR<V>(Aop, Aop, B, ip, xnorm, iters, size_t(max_iterations), lambda, btol,
                                 Atol, cond_limit, Delta);
    if (x.size() == n) std::memcpy(x_out, x.data(), n * sizeof(double));
    else std::memset(x_out, 0, n * sizeof(double));
    *xnorm_out = xnorm;
    *num_iterations = iters;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

// Reference EuclideanTNLS (Riemannian/TNLS.h:265-729) on a separable problem every operation of which is a device
// level-1 kernel:  F(x)_i = (d_i (x_i x_i) + x_i) - b_i,  DF(x) = DF(x)^T = diag(2 (d_i x_i) + 1).
int ref_tnls_elem(uint64_t n, const double *d, const double *b, const double *x0, uint64_t max_iterations,
                  double root_tol, double grad_tol, double rel_tol, double step_tol, double Delta_tol,
                  double *x_out, int *status, uint64_t *n_outer, uint64_t cap, uint64_t *inner_iterations,
                  double *rho_out, double *radius_out, double *fvals_out, double *f_out, double *gradnorm_out) {
  using V = HostMat;
  std::vector<double> jd(n, 0.0);
  Riemannian::Mapping<V, V> F = [&](const V &x) {
    V out(n);
    for (size_t i = 0; i < n; ++i) {
      const double xx = x.d[i] * x.d[i];
      const double u = d[i] * xx;
      out.d[i] = (u + x.d[i]) - b[i];
    }
    return out;
  };
  Riemannian::JacobianPairFunction<V, V, V> JF = [&](const V &x) {
    for (size_t i = 0; i < n; ++i) jd[i] = 2.0 * (d[i] * x.d[i]) + 1.0;
    Riemannian::Jacobian<V, V, V> DF = [&](const V &, const V &v) {
      V out(n);
      for (size_t i = 0; i < n; ++i) out.d[i] = jd[i] * v.d[i];
      return out;
    };
    Riemannian::JacobianAdjoint<V, V, V> DFt = [&](const V &, const V &w) {
      V out(n);
      for (size_t i = 0; i < n; ++i) out.d[i] = jd[i] * w.d[i];
      return out;
    };
    return std::make_pair(DF, DFt);
  };
  Riemannian::TNLSParams<double> params;
  params.max_iterations = size_t(max_iterations);
  params.root_tolerance = root_tol;
  params.gradient_tolerance = grad_tol;
  params.relative_decrease_tolerance = rel_tol;
  params.stepsize_tolerance = step_tol;
  params.Delta_tolerance = Delta_tol;
  try {
    V X0(x0, n);
    const std::optional<Riemannian::TNLSPreconditioner<V, V>> no_precon;
    auto res = Riemannian::EuclideanTNLS<V>(F, JF, X0, no_precon, params);
    std::memcpy(x_out, res.x.data(), n * sizeof(double));
    *status = int(res.status);
    *n_outer = res.inner_iterations.size();
    for (size_t i = 0; i < res.inner_iterations.size() && i < cap; ++i) {
      inner_iterations[i] = res.inner_iterations[i];
      rho_out[i] = res.rho[i];
    }
    for (size_t i = 0; i < res.trust_region_radius.size() && i < cap; ++i) {
      radius_out[i] = res.trust_region_radius[i];
      fvals_out[i] = res.objective_values[i];
    }
    *f_out = res.f;
    *gradnorm_out = res.gradfx_norm;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

// Reference TNLS (Riemannian/TNLS.h:265-729) on the curve-fitting problem of the reference's
// tests/TNLS_unit_test.cpp:  F(beta)_i = y_i - sin(beta_0 t_i + beta_1),  t = linspace(-pi, pi, m),
// optional right preconditioner M = diag(1 / |J column|) (diagonal stand-in for the test's R^{-1}).
int ref_tnls_sine(uint64_t m, const double *t, const double *y, const double *beta0, int use_precon,
                  uint64_t max_iterations, double root_tol, double grad_tol, double rel_tol,
                  double step_tol, double Delta_tol, double *beta_out, int *status,
                  uint64_t *n_outer, uint64_t cap, uint64_t *inner_iterations, double *rho_out,
                  double *radius_out, double *fvals_out, double *f_out, double *gradnorm_out) {
  using V = HostMat;
  std::vector<double> Jm(2 * m, 0.0), scale(2, 1.0);    // Jacobian columns, preconditioner diagonal
  Riemannian::Mapping<V, V> F = [&](const V &b) {
    V out(m);
    for (size_t i = 0; i < m; ++i) out.d[i] = y[i] - std::sin(b.d[0] * t[i] + b.d[1]);
    return out;
  };
  Riemannian::JacobianPairFunction<V, V, V> JF = [&](const V &b) {
    for (size_t i = 0; i < m; ++i) {
      Jm[m + i] = -std::cos(b.d[0] * t[i] + b.d[1]);
      Jm[i] = Jm[m + i] * t[i];
    }
    for (int c = 0; c < 2; ++c) {
      double s2 = 0;
      for (size_t i = 0; i < m; ++i) s2 += Jm[c * m + i] * Jm[c * m + i];
      scale[c] = 1.0 / std::sqrt(s2);
    }
    Riemannian::Jacobian<V, V, V> DF = [&](const V &, const V &v) {
      V out(m);
      for (size_t i = 0; i < m; ++i) out.d[i] = Jm[i] * v.d[0] + Jm[m + i] * v.d[1];
      return out;
    };
    Riemannian::JacobianAdjoint<V, V, V> DFt = [&](const V &, const V &w) {
      V out(2);
      for (int c = 0; c < 2; ++c) {
        double acc = 0;
        for (size_t i = 0; i < m; ++i) acc += Jm[c * m + i] * w.d[i];
        out.d[c] = acc;
      }
      return out;
    };
    return std::make_pair(DF, DFt);
  };
  Riemannian::LinearOperator<V, V> M = [&](const V &, const V &v) {
    V out(2);
    out.d[0] = scale[0] * v.d[0];
    out.d[1] = scale[1] * v.d[1];
    return out;
  };
  std::optional<Riemannian::TNLSPreconditioner<V, V>> precon;
  if (use_precon) precon = std::make_pair(M, M);
  Riemannian::TNLSParams<double> params;
  params.max_iterations = size_t(max_iterations);
  params.root_tolerance = root_tol;
  params.gradient_tolerance = grad_tol;
  params.relative_decrease_tolerance = rel_tol;
  params.stepsize_tolerance = step_tol;
  params.Delta_tolerance = Delta_tol;
  try {
    V B0(beta0, 2);
    auto res = Riemannian::EuclideanTNLS<V>(F, JF, B0, precon, params);
    beta_out[0] = res.x.d[0];
    beta_out[1] = res.x.d[1];
    *status = int(res.status);
    *n_outer = res.inner_iterations.size();
    for (size_t i = 0; i < res.inner_iterations.size() && i < cap; ++i) {
      inner_iterations[i] = res.inner_iterations[i];
      rho_out[i] = res.rho[i];
    }
    for (size_t i = 0; i < res.trust_region_radius.size() && i < cap; ++i) {
      radius_out[i] = res.trust_region_radius[i];
      fvals_out[i] = res.objective_values[i];
    }
    *f_out = res.f;
    *gradnorm_out = res.gradfx_norm;
  } catch (const std::invalid_argument &) {
    return 1;
  }
  return 0;
}

} // extern "C"
