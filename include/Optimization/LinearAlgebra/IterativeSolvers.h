// Drop-in header layer: Steihaug-Toint truncated preconditioned (projected) CG with the reference's
// signature and semantics (reference: include/Optimization/LinearAlgebra/IterativeSolvers.h:166-426),
// written from scratch.  Two execution paths:
//   * device-fused: when Vector is Optimization::b200::DeviceMatrix and both H and inner_product
//     wrap descriptor functors (b200::BoundHessian / b200::FrobeniusProduct), no constraint operator
//     and no per-iteration hook are supplied, and the preconditioner is absent or a b200::BoundJacobi,
//     the whole loop runs in ONE persistent CUDA kernel behind the C ABI (ob200_stpcg);
//   * generic: any Vector type with the operators listed in SURVEY.md 8(b); the loop below only does
//     control flow, all arithmetic is the Vector type's (for DeviceMatrix: level-1 CUDA kernels).
#pragma once
#include <cmath>
#include <limits>
#include <optional>
#include <stdexcept>
#include <utility>

#include "Optimization/LinearAlgebra/Concepts.h"

namespace Optimization {
namespace LinearAlgebra {

template <typename Vector, typename Multiplier, typename Scalar = double, typename... Args>
using STPCGUserFunction = std::function<bool(
    size_t k, const Vector &g, const SymmetricLinearOperator<Vector, Args...> &H,
    const std::optional<LinearOperator<Vector, std::pair<Vector, Multiplier>, Args...>> &P,
    const std::optional<LinearOperator<Multiplier, Vector, Args...>> &At, const Vector &sk, const Vector &rk,
    const Vector &vk, const Vector &pk, Scalar alpha_k, Args &...args)>;

template <typename Vector, typename Multiplier, typename... Args>
using STPCGPreconditioner = LinearOperator<Vector, std::pair<Vector, Multiplier>, Args...>;

namespace detail {
// Customisation point for device-resident vector types: returns true and fills the outputs if the
// fused device path handled the call.  The primary template declines.
template <typename Vector, typename Multiplier, typename Scalar, typename... Args>
struct FusedSTPCG {
  static bool run(const Vector &, const SymmetricLinearOperator<Vector, Args...> &,
                  const InnerProduct<Vector, Scalar, Args...> &, Scalar &, size_t &, Scalar, size_t, Scalar, Scalar,
                  const std::optional<STPCGPreconditioner<Vector, Multiplier, Args...>> &,
                  const std::optional<LinearOperator<Multiplier, Vector, Args...>> &,
                  const std::optional<STPCGUserFunction<Vector, Multiplier, Scalar, Args...>> &, Scalar, Vector &) {
    return false;
  }
};
}  // namespace detail

template <typename Vector, typename Multiplier, typename Scalar = double, typename... Args>
Vector STPCG(const Vector &g, const SymmetricLinearOperator<Vector, Args...> &H,
             const InnerProduct<Vector, Scalar, Args...> &inner_product, Args &...args, Scalar &update_step_M_norm,
             size_t &num_iterations, Scalar Delta, size_t max_iterations = 1000, Scalar kappa_fgr = .1,
             Scalar theta = .5,
             const std::optional<STPCGPreconditioner<Vector, Multiplier, Args...>> &P = std::nullopt,
             const std::optional<LinearOperator<Multiplier, Vector, Args...>> &At = std::nullopt,
             const std::optional<STPCGUserFunction<Vector, Multiplier, Scalar, Args...>> &user_function =
                 std::nullopt,
             Scalar epsilon = 1e-8) {
  // same admissible ranges (and the same std::invalid_argument) as the reference, l.183-205
  if (!(Delta > 0)) throw std::invalid_argument("Trust-region radius (Delta) must be a positive real value");
  if (kappa_fgr < 0 || kappa_fgr >= 1)
    throw std::invalid_argument(
        "Target fractional reduction of the gradient norm (kappa_fgr) must be a real value in the range [0,1)");
  if (theta < 0 || theta > 1)
    throw std::invalid_argument(
        "Target superlinear convergence rate (theta) must be a real value in the range [0,1]");
  if (epsilon <= 0 || epsilon >= 1)
    throw std::invalid_argument("Relative norm tolerance for declaring a vector to lie in the kernel of H (epsilon) "
                                "should be a small positive number in the range (0,1)");

  {
    Vector fused;
    if (detail::FusedSTPCG<Vector, Multiplier, Scalar, Args...>::run(g, H, inner_product, update_step_M_norm,
                                                                       num_iterations, Delta, max_iterations,
                                                                       kappa_fgr, theta, P, At, user_function,
                                                                       epsilon, fused))
      return fused;
  }

  // ---- generic path: control flow only --------------------------------------------------------
  const bool projected = P && At;            // constraint multipliers are only meaningful with both
  Vector s = 0 * g;
  Vector r = g;
  Vector v;
  Multiplier lambda = Multiplier();
  if (P) {
    std::pair<Vector, Multiplier> pre = (*P)(r, args...);
    v = std::move(pre.first);
    lambda = std::move(pre.second);
    if (projected) r -= (*At)(lambda, args...);
  } else {
    v = r;
  }
  Vector p = -v;

  // M-norm bookkeeping is propagated by recurrences: no product with M is ever formed
  Scalar s_M_p = 0, s_M_s = 0;
  Scalar p_M_p = inner_product(r, v, args...);
  const Scalar Delta2 = Delta * Delta;
  const Scalar r0 = std::sqrt(inner_product(r, v, args...));
  const Scalar target = r0 * std::min(kappa_fgr, std::pow(r0, theta));

  auto to_boundary = [&]() {   // positive root of |s + sigma p|_M = Delta
    return (-s_M_p + std::sqrt(s_M_p * s_M_p + p_M_p * (Delta2 - s_M_s))) / p_M_p;
  };

  Vector Hp;
  for (num_iterations = 0; num_iterations < max_iterations; ++num_iterations) {
    if (std::sqrt(inner_product(r, v, args...)) <= target) break;

    Hp = H(p, args...);
    const Scalar kappa = inner_product(p, Hp, args...);

    // p (numerically) in ker H: move to the boundary along the descent orientation of p
    if (std::sqrt(inner_product(Hp, Hp, args...)) / std::sqrt(inner_product(p, p, args...)) < epsilon) {
      if (inner_product(p, r, args...) < 0) {
        p *= -1;
        s_M_p *= -1;
      }
      s += to_boundary() * p;
      update_step_M_norm = Delta;
      return s;
    }

    const Scalar alpha = inner_product(r, v, args...) / kappa;
    const Scalar next_M = s_M_s + 2 * alpha * s_M_p + alpha * alpha * p_M_p;

    if (kappa <= 0 || next_M > Delta2) {     // negative curvature or the full step leaves the region
      s += to_boundary() * p;
      update_step_M_norm = Delta;
      return s;
    }

    if (user_function && (*user_function)(num_iterations, g, H, P, At, s, r, v, p, alpha, args...)) break;

    s = s + alpha * p;
    r += alpha * Hp;
    if (P) {
      std::pair<Vector, Multiplier> pre = (*P)(r, args...);
      v = std::move(pre.first);
      lambda = std::move(pre.second);
      if (projected) r -= (*At)(lambda, args...);
    } else {
      v = r;
    }

    const Scalar rv = inner_product(r, v, args...);
    const Scalar beta = rv / (alpha * kappa);
    s_M_s = next_M;
    s_M_p = beta * (s_M_p + alpha * p_M_p);
    p_M_p = rv + beta * beta * p_M_p;
    p = -v + beta * p;
  }
  update_step_M_norm = std::sqrt(s_M_s);
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// LSQR (Paige & Saunders) for  min_x |A x - b|^2 + lambda |x|^2  s.t. |x| <= Delta, with the reference's
// signature, stopping rules and trust-region modification (reference IterativeSolvers.h:451-875), written from
// scratch.  Control flow and scalar recurrences only; all vector arithmetic is the Vector types' own (for
// b200::DeviceMatrix: the level-1 CUDA kernels).  Vector requirements beyond STPCG's: `v /= Scalar`,
// `Vector - Scalar * Vector`.
// ---------------------------------------------------------------------------------------------------------

// Called at the END of every iteration; returning true stops the iteration.
template <typename VectorX, typename VectorY, typename Scalar = double, typename... Args>
using LSQRUserFunction =
    std::function<bool(size_t k, const LinearOperator<VectorX, VectorY, Args...> &A,
                       const LinearOperator<VectorY, VectorX, Args...> &At, const VectorY &b, const VectorX &xk,
                       Scalar xk_norm, Scalar rbar_norm, Scalar Abar_rbar_norm, Scalar Abar_norm_est,
                       Scalar Abar_cond_est, Args &...args)>;

template <typename VectorX, typename VectorY, typename Scalar = double, typename... Args>
VectorX LSQR(const LinearOperator<VectorX, VectorY, Args...> &A, const LinearOperator<VectorY, VectorX, Args...> &At,
             const VectorY &b, const InnerProduct<VectorX, Scalar, Args...> &inner_product_x,
             const InnerProduct<VectorY, Scalar, Args...> &inner_product_y, Args &...args, Scalar &xnorm,
             size_t &num_iterations, size_t max_iterations = 1000, Scalar lambda = 0, Scalar btol = 1e-6,
             Scalar Atol = 1e-6, Scalar Abar_cond_limit = 1e8,
             Scalar Delta = std::sqrt(std::numeric_limits<Scalar>::max()),
             const std::optional<LSQRUserFunction<VectorX, VectorY, Scalar, Args...>> &user_function = std::nullopt) {
  if (lambda < 0)
    throw std::invalid_argument("Tikhonov regularization parameter (lambda) must be a nonnegative real value");
  if (btol < 0) throw std::invalid_argument("Stopping tolerance btol must be a nonnegative real number");
  if (Atol < 0) throw std::invalid_argument("Stopping tolerance Atol must be a nonnegative real number");
  if (Abar_cond_limit <= 0)
    throw std::invalid_argument("Stopping tolerance Abar_cond_limit must be a positive real number");
  if (Delta <= 0) throw std::invalid_argument("Trust-region radius (Delta) must be a positive real value");

  xnorm = 0;
  num_iterations = 0;
  const Scalar sqrt_lambda = std::sqrt(lambda);

  // Golub-Kahan start: beta u = b, alpha v = A^T u
  VectorY u = b;
  VectorX v = At(u, args...);
  VectorX x = 0 * v;
  VectorX w;
  Scalar alpha = std::sqrt(inner_product_x(v, v, args...));
  Scalar beta = std::sqrt(inner_product_y(u, u, args...));
  if (beta > 0) u /= beta;
  if (alpha > 0) {
    v /= alpha;
    alpha /= beta;   // v was formed from the unnormalised u
    w = v;
  }
  Scalar Abar_rbar_norm = alpha * beta;
  if (Abar_rbar_norm == 0) return x;   // x = 0 is already stationary

  const Scalar bnorm = beta;
  Scalar rbar_norm = beta;
  Scalar Abar_norm_est = 0, Abar_cond_est = 0;
  Scalar D_fro2 = 0;         // squared Frobenius norm of the direction matrix [d_1 ... d_k]
  Scalar xx = 0;             // running |x|^2 estimate from the right rotations
  Scalar rhobar = alpha, phibar = beta;
  Scalar cs2 = -1, sn2 = 0, z = 0, res2 = 0;

  for (num_iterations = 0; num_iterations < max_iterations; ++num_iterations) {
    // next bidiagonalisation step:  beta u = A v - alpha u ;  alpha v = A^T u - beta v
    u = A(v, args...) - alpha * u;
    beta = std::sqrt(inner_product_y(u, u, args...));
    if (beta > 0) {
      u /= beta;
      Abar_norm_est = std::sqrt(Abar_norm_est * Abar_norm_est + alpha * alpha + beta * beta + lambda);
      v = At(u, args...) - beta * v;
      alpha = std::sqrt(inner_product_x(v, v, args...));
      if (alpha > 0) v /= alpha;
    }

    // rotation 1: remove the damping term from the diagonal
    const Scalar rhobar1 = std::sqrt(rhobar * rhobar + lambda);
    const Scalar cs1 = rhobar / rhobar1;
    const Scalar sn1 = sqrt_lambda / rhobar1;
    const Scalar psi = sn1 * phibar;
    phibar *= cs1;

    // rotation 2: remove the subdiagonal beta (lower -> upper bidiagonal)
    const Scalar rho = std::sqrt(rhobar1 * rhobar1 + beta * beta);
    const Scalar cs = rhobar1 / rho;
    const Scalar sn = beta / rho;
    const Scalar theta = sn * alpha;
    rhobar = -cs * alpha;
    const Scalar phi = cs * phibar;
    phibar *= sn;
    const Scalar tau = sn * phi;

    // rotation 3 (from the right): remove the superdiagonal theta; yields the |x| estimate
    const Scalar delta = sn2 * rho;
    const Scalar gammabar = -cs2 * rho;
    const Scalar rhs = phi - delta * z;
    const Scalar zbar = rhs / gammabar;
    const Scalar gamma = std::sqrt(gammabar * gammabar + theta * theta);
    cs2 = gammabar / gamma;
    sn2 = theta / gamma;
    z = rhs / gamma;

    const Scalar w2 = inner_product_x(w, w, args...);
    const Scalar d2 = w2 / (rho * rho);
    xnorm = std::sqrt(xx + zbar * zbar);   // |x| AFTER the full update
    xx += z * z;

    const Scalar t2 = -theta / rho;        // step for w
    Scalar t1;                             // step for x
    if (xnorm <= Delta) {
      t1 = phi / rho;
    } else {                               // full step leaves the region: stop on the boundary
      const Scalar xtx = inner_product_x(x, x, args...);
      const Scalar wtx = inner_product_x(w, x, args...);
      t1 = (-wtx + std::sqrt(wtx * wtx + w2 * (Delta * Delta - xtx))) / w2;
      xnorm = Delta;
    }
    x += t1 * w;
    w = v + t2 * w;
    D_fro2 += d2;

    Abar_cond_est = Abar_norm_est * std::sqrt(D_fro2);
    const Scalar res1 = phibar * phibar;
    res2 += psi * psi;
    rbar_norm = std::sqrt(res1 + res2);
    Abar_rbar_norm = alpha * std::fabs(tau);

    if (rbar_norm <= btol * bnorm + Atol * Abar_norm_est * xnorm) break;    // S1: residual (consistent systems)
    if (Abar_rbar_norm <= Atol * Abar_norm_est * rbar_norm) break;          // S2: gradient (inconsistent systems)
    if (Abar_cond_est >= Abar_cond_limit) break;                            // conditioning
    if (xnorm >= Delta) break;                                              // trust-region boundary
    if (user_function && (*user_function)(num_iterations, A, At, b, x, xnorm, rbar_norm, Abar_rbar_norm,
                                          Abar_norm_est, Abar_cond_est, args...))
      break;
  }
  return x;
}

// Same space for domain and codomain.
template <typename Vector, typename Scalar = double, typename... Args>
Vector LSQR(const LinearOperator<Vector, Vector, Args...> &A, const LinearOperator<Vector, Vector, Args...> &At,
            const Vector &b, const InnerProduct<Vector, Scalar, Args...> &inner_product, Args &...args, Scalar &xnorm,
            size_t &num_iterations, size_t max_iterations = 1000, Scalar lambda = 0, Scalar btol = 1e-6,
            Scalar Atol = 1e-6, Scalar Abar_cond_limit = 1e8,
            Scalar Delta = std::sqrt(std::numeric_limits<Scalar>::max()),
            const std::optional<LSQRUserFunction<Vector, Vector, Scalar, Args...>> &user_function = std::nullopt) {
  return LSQR<Vector, Vector, Scalar, Args...>(A, At, b, inner_product, inner_product, args..., xnorm, num_iterations,
                                               max_iterations, lambda, btol, Atol, Abar_cond_limit, Delta,
                                               user_function);
}

}  // namespace LinearAlgebra
}  // namespace Optimization
