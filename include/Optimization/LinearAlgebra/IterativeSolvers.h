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

}  // namespace LinearAlgebra
}  // namespace Optimization
