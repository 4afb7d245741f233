// Drop-in header layer: Riemannian truncated-Newton trust-region method for nonlinear least squares,
//   min_x |F(x)|,  F: X -> Y (Y a Euclidean space),
// with the reference's entry points, parameter / result types and stopping semantics (reference:
// include/Optimization/Riemannian/TNLS.h:44-768), written from scratch.  Each outer iteration solves the
// linearised problem  min_h |F(x) + DF(x) h|  s.t. |h|_M <= Delta  inexactly with LSQR
// (Optimization::LinearAlgebra::LSQR) on the (optionally right-preconditioned) Jacobian.  Host control flow only.
#pragma once
#include <algorithm>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <limits>
#include <optional>
#include <stdexcept>
#include <tuple>

#include "Optimization/LinearAlgebra/IterativeSolvers.h"
#include "Optimization/Riemannian/Concepts.h"
#include "Optimization/Util/Stopwatch.h"

namespace Optimization {
namespace Riemannian {

// Right preconditioner M (first) and its transpose M^T (second): the inner solver works on DF(x) M.
template <typename VariableX, typename TangentX, typename... Args>
using TNLSPreconditioner =
    std::pair<LinearOperator<VariableX, TangentX, Args...>, LinearOperator<VariableX, TangentX, Args...>>;

// Called once per outer iteration; returning true stops the method.
template <typename VariableX, typename TangentX, typename VectorY, typename Scalar = double, typename... Args>
using TNLSUserFunction =
    std::function<bool(size_t i, double t, const VariableX &x, VectorY Fx,
                       const Jacobian<VariableX, TangentX, VectorY, Args...> &gradFx,
                       const JacobianAdjoint<VariableX, TangentX, VectorY, Args...> &gradFxT, Scalar Delta,
                       size_t num_LSQR_iters, const TangentX &h, Scalar dL, Scalar rho, bool accepted, Args &...args)>;

template <typename Scalar = double>
struct TNLSParams : public SmoothOptimizerParams<Scalar> {
  Scalar Delta0 = 1;                  // initial trust-region radius
  Scalar eta1 = .05;                  // gain ratio of a successful step
  Scalar eta2 = .9;                   // gain ratio of a very successful step
  Scalar alpha1 = .25;                // radius shrink factor
  Scalar alpha2 = 2.5;                // radius growth factor
  size_t max_LSQR_iterations = 1000;  // inner iteration cap
  Scalar kappa_fgr = .1;              // inner target: |r| <= kappa_fgr |F(x)|
  Scalar theta = .5;                  // inner target: |r| <= |F(x)|^(1 + theta)
  Scalar lambda = 0;                  // Tikhonov regularisation of the inner problem
  Scalar Atol = 1e-6;                 // LSQR's relative gradient tolerance
  Scalar Acond_limit = 1e8;           // LSQR's conditioning limit
  Scalar root_tolerance = 1e-6;       // stop when |F(x)| falls below this
  Scalar Delta_tolerance = 1e-6;      // stop when the radius falls below this
};

enum class TNLSStatus { Root, Gradient, RelativeDecrease, Stepsize, TrustRegion, IterationLimit, ElapsedTime, UserFunction };

template <typename Variable, typename Scalar = double>
struct TNLSResult : public SmoothOptimizerResult<Variable, Scalar> {
  TNLSStatus status;
  std::vector<size_t> inner_iterations;
  std::vector<Scalar> rho;                    // gain ratio of every iteration
  std::vector<Scalar> trust_region_radius;    // radius at the START of every iteration
};

template <typename VariableX, typename TangentX, typename VectorY, typename Scalar = double, typename... Args>
TNLSResult<VariableX, Scalar>
TNLS(const Mapping<VariableX, VectorY, Args...> &F, const JacobianPairFunction<VariableX, TangentX, VectorY> &J,
     const RiemannianMetric<VariableX, TangentX, Scalar, Args...> &metric_X,
     const LinearAlgebra::InnerProduct<VectorY, Scalar, Args...> &inner_product_Y,
     const Retraction<VariableX, TangentX, Args...> &retract_X, const VariableX &x0, Args &...args,
     const std::optional<TNLSPreconditioner<VariableX, TangentX, Args...>> &precon = std::nullopt,
     const TNLSParams<Scalar> &params = TNLSParams<Scalar>(),
     const std::optional<TNLSUserFunction<VariableX, TangentX, VectorY, Scalar, Args...>> &user_function =
         std::nullopt) {
  // admissible ranges: the reference's (its TNLS.h:284-352)
  auto need = [](bool ok, const char *msg) {
    if (!ok) throw std::invalid_argument(msg);
  };
  need(params.max_computation_time >= 0, "Maximum computation time must be a nonnegative real value");
  need(params.root_tolerance >= 0, "Root tolerance must be a nonnegative real value");
  need(params.gradient_tolerance >= 0, "Gradient tolerance must be a nonnegative real value");
  need(params.relative_decrease_tolerance >= 0, "Relative decrease tolerance must be a nonnegative real value");
  need(params.stepsize_tolerance >= 0, "Stepsize tolerance must be a nonnegative real value");
  need(params.Delta_tolerance >= 0, "Trust-region radius tolerance must be a nonnegative real value");
  need(params.Delta0 > 0, "Initial trust-region radius must be a positive real value");
  need(params.eta1 > 0 && params.eta1 < 1,
       "Threshold on gain ratio for a successful iteration (eta1) must satisfy 0 < eta1 < 1");
  need(params.eta1 <= params.eta2 && params.eta2 < 1,
       "Threshold on gain ratio for a very successful iteration (eta2) must satisfy eta1 <= eta2 < 1");
  need(params.alpha1 > 0 && params.alpha1 < 1,
       "Multiplicative factor for decreasing trust-region radius (alpha1) must satisfy 0 < alpha1 < 1");
  need(params.alpha2 > 1, "Multiplicative factor for increasing trust-region radius (alpha1) must satisfy alpha2 > 1");
  need(params.kappa_fgr > 0 && params.kappa_fgr < 1,
       "Target relative decrease in predicted residual for inexact update step computation (kappa_fgr) must satisfy "
       "0 < kappa_fgr < 1");
  need(params.theta >= 0, "Target superlinear convergence rate parameter (theta) must be a nonnegative real number");
  need(params.Atol >= 0, "Relative norm stopping tolerance Atol must be a nonnegative real number");
  need(params.Acond_limit > 0, "Stopping criterion Acond_limit must be a positive real number");

  using JacobianOp = Jacobian<VariableX, TangentX, VectorY>;
  using AdjointOp = JacobianAdjoint<VariableX, TangentX, VectorY>;

  const Scalar sqrt_eps = std::sqrt(std::numeric_limits<Scalar>::epsilon());
  const bool talk = params.verbose;
  const int it_width = int(std::floor(std::log10(double(params.max_iterations)))) + 1;
  const int in_width = int(std::floor(std::log10(double(params.max_LSQR_iterations)))) + 1;

  TNLSResult<VariableX, Scalar> out;
  out.status = TNLSStatus::IterationLimit;

  // state at the current iterate
  VariableX x = x0;
  VectorY Fx = F(x, args...);
  Scalar F2 = inner_product_Y(Fx, Fx, args...);
  Scalar Fnorm = std::sqrt(F2);
  JacobianOp DF;
  AdjointOp DFt;
  std::tie(DF, DFt) = J(x);
  // the objective is L(x) = |F(x)|:  grad L(x) = DF(x)^T F(x) / |F(x)|
  TangentX gradL = DFt(x, Fx) / Fnorm;
  Scalar gnorm = std::sqrt(metric_X(x, gradL, gradL, args...));

  // x-bound views handed to LSQR (they follow x, DF, DFt as the iteration moves)
  LinearAlgebra::LinearOperator<TangentX, VectorY, Args...> A = [&](const TangentX &v, Args &...a) -> VectorY {
    return precon ? DF(x, precon->first(x, v, a...)) : DF(x, v);
  };
  LinearAlgebra::LinearOperator<VectorY, TangentX, Args...> At = [&](const VectorY &w, Args &...a) -> TangentX {
    return precon ? precon->second(x, DFt(x, w), a...) : DFt(x, w);
  };
  LinearAlgebra::InnerProduct<TangentX, Scalar, Args...> inner_product_X =
      [&](const TangentX &v1, const TangentX &v2, Args &...a) -> Scalar { return metric_X(x, v1, v2, a...); };

  Scalar Delta = params.Delta0;
  Scalar rel_decrease = 0, hnorm = 0;

  if (talk) {
    std::cout << std::scientific << std::setprecision(int(params.precision));
    std::cout << "Truncated-Newton least squares (trust region, inner solver LSQR)\n" << std::endl;
  }

  const auto t0 = Stopwatch::tick();
  for (size_t it = 0; it < params.max_iterations; ++it) {
    const double now = Stopwatch::tock(t0);
    if (now > params.max_computation_time) {
      out.status = TNLSStatus::ElapsedTime;
      break;
    }
    out.time.push_back(now);
    out.objective_values.push_back(Fnorm);
    out.gradient_norms.push_back(gnorm);
    out.trust_region_radius.push_back(Delta);
    if (params.log_iterates) out.iterates.push_back(x);
    if (talk)
      std::cout << "[" << std::setw(it_width) << it << "] t " << now << "  |F| " << std::setw(int(params.precision) + 7) << Fnorm
                << "  |grad| " << gnorm;

    if (Fnorm < params.root_tolerance) {
      out.status = TNLSStatus::Root;
      break;
    }
    if (gnorm < params.gradient_tolerance) {
      out.status = TNLSStatus::Gradient;
      break;
    }

    // inexact Gauss-Newton step: forcing term min(|F|^theta, kappa_fgr) as LSQR's residual tolerance
    const Scalar forcing = std::min(std::pow(Fnorm, params.theta), params.kappa_fgr);
    size_t inner = 0;
    Scalar h_M_norm = 0;
    TangentX h = LinearAlgebra::LSQR<TangentX, VectorY, Scalar, Args...>(
        A, At, -Fx, inner_product_X, inner_product_Y, args..., h_M_norm, inner, params.max_LSQR_iterations,
        params.lambda, forcing, params.Atol, params.Acond_limit, Delta);
    if (precon) h = precon->first(x, h, args...);   // back to the original coordinates
    hnorm = std::sqrt(metric_X(x, h, h, args...));
    if (talk)
      std::cout << "  radius " << Delta << "  LSQR " << std::setw(in_width) << inner << "  |h| " << hnorm << "  |h|_M " << h_M_norm;

    // trial point, actual vs predicted decrease of |F|^2
    VariableX x_trial = retract_X(x, h, args...);
    VectorY F_trial = F(x_trial, args...);
    const Scalar F2_trial = inner_product_Y(F_trial, F_trial, args...);
    const Scalar Fnorm_trial = std::sqrt(F2_trial);
    const VectorY r = DF(x, h) + Fx;
    const Scalar r2 = inner_product_Y(r, r, args...);
    const Scalar predicted = F2 - r2;
    const Scalar dL = Fnorm - Fnorm_trial;
    const Scalar actual = F2 - F2_trial;
    rel_decrease = dL / (sqrt_eps + Fnorm);
    const Scalar rho = actual / predicted;
    const bool accepted = !std::isnan(rho) && rho > params.eta1;
    if (talk)
      std::cout << "  decrease " << std::setw(int(params.precision) + 7) << dL << "  gain " << std::setw(int(params.precision) + 7)
                << rho << (accepted ? "  accepted" : "  rejected");

    out.inner_iterations.push_back(inner);
    out.update_step_norms.push_back(hnorm);
    out.rho.push_back(rho);

    if (user_function && (*user_function)(it, now, x, Fx, DF, DFt, Delta, inner, h, dL, rho, accepted, args...)) {
      out.status = TNLSStatus::UserFunction;
      break;
    }

    if (accepted) {
      x = std::move(x_trial);
      Fx = std::move(F_trial);
      F2 = F2_trial;
      Fnorm = Fnorm_trial;
      if (rel_decrease < params.relative_decrease_tolerance) {
        out.status = TNLSStatus::RelativeDecrease;
        break;
      }
      if (hnorm < params.stepsize_tolerance) {
        out.status = TNLSStatus::Stepsize;
        break;
      }
      std::tie(DF, DFt) = J(x);
      gradL = DFt(x, Fx) / Fnorm;
      gnorm = std::sqrt(metric_X(x, gradL, gradL, args...));
    }

    // radius update
    if (!std::isnan(rho) && rho >= params.eta2) {
      Delta = std::max<Scalar>(params.alpha2 * h_M_norm, Delta);
    } else if (std::isnan(rho) || rho < params.eta1) {
      Delta = params.alpha1 * h_M_norm;
      if (Delta < params.Delta_tolerance) {
        out.status = TNLSStatus::TrustRegion;
        break;
      }
    }
    if (talk) std::cout << std::endl;
  }

  out.elapsed_time = Stopwatch::tock(t0);
  out.x = x;
  out.f = Fnorm;
  out.gradfx_norm = gnorm;

  if (talk) {
    static const char *const why[] = {"residual norm below root_tolerance", "gradient norm below gradient_tolerance",
                                      "relative decrease below relative_decrease_tolerance",
                                      "step length below stepsize_tolerance", "trust-region radius below Delta_tolerance",
                                      "iteration limit reached", "time limit exceeded", "stopped by the user function"};
    std::cout << "\n\nTNLS stopped: " << why[int(out.status)] << "\n  |F(x)| = " << out.f << ", |grad| = " << out.gradfx_norm
              << ", last relative decrease " << rel_decrease << ", last |h| " << hnorm << ", Delta " << Delta << ", "
              << out.elapsed_time << " s\n" << std::endl;
    std::cout << std::defaultfloat << std::setprecision(6);
  }
  return out;
}

// Euclidean convenience (reference TNLS.h:749-765): standard inner products, retraction x + v, one Vector type.
template <typename Vector, typename Scalar = double, typename... Args>
TNLSResult<Vector, Scalar>
EuclideanTNLS(const Mapping<Vector, Vector, Args...> &F, const JacobianPairFunction<Vector, Vector, Vector> &J,
              const Vector &x0, Args &...args,
              const std::optional<TNLSPreconditioner<Vector, Vector, Args...>> &precon = std::nullopt,
              const TNLSParams<Scalar> &params = TNLSParams<Scalar>(),
              const std::optional<TNLSUserFunction<Vector, Vector, Vector, Scalar, Args...>> &user_function =
                  std::nullopt) {
  return TNLS<Vector, Vector, Vector, Scalar, Args...>(F, J, EuclideanMetric<Vector, Scalar, Args...>,
                                                       EuclideanInnerProduct<Vector, Scalar, Args...>,
                                                       EuclideanRetraction<Vector, Args...>, x0, args..., precon, params,
                                                       user_function);
}

}  // namespace Riemannian
}  // namespace Optimization
