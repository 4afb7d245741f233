// Drop-in header layer: Riemannian truncated-Newton trust-region method with the reference's entry
// points, parameter / result types and stopping semantics (reference:
// include/Optimization/Riemannian/TNT.h:64-805), written from scratch around the B200 tCG path.
// The outer loop is host control flow; the inner Steihaug-Toint solve goes through
// Optimization::LinearAlgebra::STPCG, which dispatches to the fused CUDA kernel when the tangent
// type is Optimization::b200::DeviceMatrix and the functors are descriptor functors.
#pragma once
#include <algorithm>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <limits>
#include <optional>
#include <stdexcept>

#include "Optimization/LinearAlgebra/IterativeSolvers.h"
#include "Optimization/Riemannian/Concepts.h"
#include "Optimization/Util/Stopwatch.h"

namespace Optimization {
namespace Riemannian {

template <typename Variable, typename Tangent, typename Scalar = double, typename... Args>
using TNTUserFunction = std::function<bool(size_t i, double t, const Variable &x, Scalar f, const Tangent &g,
                                           const LinearOperator<Variable, Tangent, Args...> &HessOp, Scalar Delta,
                                           size_t num_STPCG_iters, const Tangent &h, Scalar df, Scalar rho,
                                           bool accepted, Args &...args)>;

template <typename Scalar = double>
struct TNTParams : public SmoothOptimizerParams<Scalar> {
  Scalar Delta0 = 1;                                  // initial trust-region radius
  Scalar eta1 = .05;                                  // gain ratio of a successful step
  Scalar eta2 = .9;                                   // gain ratio of a very successful step
  Scalar alpha1 = .25;                                // radius shrink factor
  Scalar alpha2 = 2.5;                                // radius growth factor
  size_t max_TPCG_iterations = 1000;                  // inner iteration cap
  Scalar kappa_fgr = .1;                              // inner target: fractional gradient reduction
  Scalar theta = .5;                                  // inner target: superlinear rate
  Scalar preconditioned_gradient_tolerance = 1e-6;
  Scalar Delta_tolerance = 1e-6;                      // stop when the radius falls below this
};

enum class TNTStatus {
  Gradient,
  PreconditionedGradient,
  RelativeDecrease,
  Stepsize,
  TrustRegion,
  IterationLimit,
  ElapsedTime,
  UserFunction
};

template <typename Variable, typename Scalar = double>
struct TNTResult : public SmoothOptimizerResult<Variable, Scalar> {
  Scalar preconditioned_grad_f_x_norm;
  TNTStatus status;
  std::vector<Scalar> preconditioned_gradient_norms;
  std::vector<size_t> inner_iterations;
  std::vector<Scalar> update_step_M_norms;
  std::vector<Scalar> gain_ratios;
  std::vector<Scalar> trust_region_radius;            // radius at the START of each iteration (+ final)
};

namespace detail {
// Customisation point: how the x-bound operator views handed to the inner solver are built.  The
// primary template wraps the user's functors in lambdas that always see the current iterate; device
// layers specialise it to expose descriptor functors the fused tCG path can recognise.
template <typename Variable, typename Tangent, typename Scalar, typename... Args>
struct InnerViews {
  static LinearAlgebra::SymmetricLinearOperator<Tangent, Args...>
  hessian(const Variable &x, const LinearOperator<Variable, Tangent, Args...> &Hess) {
    return [&x, &Hess](const Tangent &v, Args &...a) -> Tangent { return Hess(x, v, a...); };
  }
  static LinearAlgebra::InnerProduct<Tangent, Scalar, Args...>
  inner_product(const Variable &x, const RiemannianMetric<Variable, Tangent, Scalar, Args...> &metric) {
    return [&x, &metric](const Tangent &a1, const Tangent &a2, Args &...a) -> Scalar { return metric(x, a1, a2, a...); };
  }
  // the preconditioner in the form the inner solver takes: (v, lambda) = P(r) with an empty multiplier
  // (reference TNT.h:413-426)
  template <typename Multiplier>
  static std::optional<LinearAlgebra::STPCGPreconditioner<Tangent, Multiplier, Args...>>
  preconditioner(const Variable &x, const std::optional<LinearOperator<Variable, Tangent, Args...>> &precon) {
    if (!precon) return std::nullopt;
    return LinearAlgebra::STPCGPreconditioner<Tangent, Multiplier, Args...>(
        [&x, &precon](const Tangent &v, Args &...a) -> std::pair<Tangent, Multiplier> {
          return {(*precon)(x, v, a...), Multiplier()};
        });
  }
};

template <typename Scalar>
void check_tnt_params(const TNTParams<Scalar> &p) {
  auto bad = [](const char *msg) { throw std::invalid_argument(msg); };
  if (p.max_computation_time < 0) bad("Maximum computation time must be a nonnegative real value");
  if (p.gradient_tolerance < 0) bad("Gradient tolerance must be a nonnegative real value");
  if (p.preconditioned_gradient_tolerance < 0)
    bad("Preconditioned gradient tolerance must be a nonnegative real value");
  if (p.relative_decrease_tolerance < 0) bad("Relative decrease tolerance must be a nonnegative real value");
  if (p.stepsize_tolerance < 0) bad("Stepsize tolerance must be a nonnegative real value");
  if (p.Delta_tolerance < 0) bad("Trust-region radius tolerance must be a nonnegative real value");
  if (p.Delta0 <= 0) bad("Initial trust-region radius must be a positive real value");
  if (p.eta1 <= 0 || p.eta1 >= 1) bad("Gain-ratio threshold eta1 must satisfy 0 < eta1 < 1");
  if (p.eta1 > p.eta2 || p.eta2 >= 1) bad("Gain-ratio threshold eta2 must satisfy eta1 <= eta2 < 1");
  if (p.alpha1 <= 0 || p.alpha1 >= 1) bad("Radius shrink factor alpha1 must satisfy 0 < alpha1 < 1");
  if (p.alpha2 <= 1) bad("Radius growth factor alpha2 must satisfy alpha2 > 1");
  if (p.kappa_fgr <= 0 || p.kappa_fgr >= 1) bad("kappa_fgr must satisfy 0 < kappa_fgr < 1");
  if (p.theta < 0) bad("theta must be a nonnegative real number");
}
}  // namespace detail

// Quadratic-model form (reference TNT.h:242-254).
template <typename Variable, typename Tangent, typename Scalar = double, typename... Args>
TNTResult<Variable, Scalar>
TNT(const Objective<Variable, Scalar, Args...> &f, const QuadraticModel<Variable, Tangent, Args...> &QM,
    const RiemannianMetric<Variable, Tangent, Scalar, Args...> &metric,
    const Retraction<Variable, Tangent, Args...> &retract, const Variable &x0, Args &...args,
    const std::optional<LinearOperator<Variable, Tangent, Args...>> &precon = std::nullopt,
    const TNTParams<Scalar> &params = TNTParams<Scalar>(),
    const std::optional<TNTUserFunction<Variable, Tangent, Scalar, Args...>> &user_function = std::nullopt) {
  detail::check_tnt_params(params);
  using Multiplier = std::nullptr_t;
  namespace LA = Optimization::LinearAlgebra;

  const Scalar sqrt_eps = std::sqrt(std::numeric_limits<Scalar>::epsilon());
  TNTResult<Variable, Scalar> out;
  out.status = TNTStatus::IterationLimit;
  if (params.log_iterates) out.iterates.reserve(params.max_iterations + 1);

  Variable x = x0;
  Scalar fx = f(x, args...);
  Tangent grad;
  LinearOperator<Variable, Tangent, Args...> Hess;
  QM(x, grad, Hess, args...);

  auto norms = [&](Scalar &gnorm, Scalar &pgnorm) {
    gnorm = std::sqrt(metric(x, grad, grad, args...));
    if (precon) {
      const Tangent pg = (*precon)(x, grad, args...);
      pgnorm = std::sqrt(metric(x, pg, pg, args...));
    } else {
      pgnorm = gnorm;
    }
  };
  Scalar gnorm, pgnorm;
  norms(gnorm, pgnorm);

  // x-bound views handed to the inner solver (rebuilt whenever QM replaces the Hessian operator)
  using Views = detail::InnerViews<Variable, Tangent, Scalar, Args...>;
  LA::SymmetricLinearOperator<Tangent, Args...> H = Views::hessian(x, Hess);
  const LA::InnerProduct<Tangent, Scalar, Args...> inner = Views::inner_product(x, metric);
  const std::optional<LA::STPCGPreconditioner<Tangent, Multiplier, Args...>> P =
      Views::template preconditioner<Multiplier>(x, precon);

  Scalar Delta = params.Delta0;
  const auto t0 = Stopwatch::tick();
  if (params.verbose) std::cout << std::scientific << std::setprecision(static_cast<int>(params.precision));

  auto record = [&](double t) {
    out.time.push_back(t);
    out.objective_values.push_back(fx);
    out.gradient_norms.push_back(gnorm);
    out.preconditioned_gradient_norms.push_back(pgnorm);
    out.trust_region_radius.push_back(Delta);
    if (params.log_iterates) out.iterates.push_back(x);
  };

  for (size_t it = 0; it < params.max_iterations; ++it) {
    const double elapsed = Stopwatch::tock(t0);
    if (elapsed > params.max_computation_time) {
      out.status = TNTStatus::ElapsedTime;
      break;
    }
    record(elapsed);
    if (params.verbose)
      std::cout << "Iter: " << it << ", time: " << elapsed << ", f: " << fx << ", |g|: " << gnorm
                << ", |M^{-1}g|: " << pgnorm;

    if (gnorm < params.gradient_tolerance) {
      out.status = TNTStatus::Gradient;
      break;
    }
    if (pgnorm < params.preconditioned_gradient_tolerance) {
      out.status = TNTStatus::PreconditionedGradient;
      break;
    }

    // inner solve: Steihaug-Toint truncated CG on the model at x
    Scalar h_M = 0;
    size_t inner_its = 0;
    const Tangent h = LA::STPCG<Tangent, Multiplier, Scalar, Args...>(
        grad, H, inner, args..., h_M, inner_its, Delta, params.max_TPCG_iterations, params.kappa_fgr, params.theta, P);
    const Scalar h_norm = std::sqrt(metric(x, h, h, args...));

    Variable x_new = retract(x, h, args...);
    const Scalar f_new = f(x_new, args...);
    const Scalar predicted = -metric(x, grad, h, args...) - .5 * metric(x, h, Hess(x, h, args...), args...);
    const Scalar df = fx - f_new;
    const Scalar rel_decrease = df / (sqrt_eps + std::fabs(fx));
    const Scalar rho = df / predicted;
    const bool accepted = !std::isnan(rho) && rho > params.eta1;

    if (params.verbose)
      std::cout << ", Delta: " << Delta << ", inner iters: " << inner_its << ", |h|: " << h_norm
                << ", |h|_M: " << h_M << ", df: " << df << ", rho: " << rho << (accepted ? " *" : "") << std::endl;

    out.inner_iterations.push_back(inner_its);
    out.update_step_norms.push_back(h_norm);
    out.update_step_M_norms.push_back(h_M);
    out.gain_ratios.push_back(rho);

    if (user_function &&
        (*user_function)(it, elapsed, x, fx, grad, Hess, Delta, inner_its, h, df, rho, accepted, args...)) {
      out.status = TNTStatus::UserFunction;
      break;
    }

    if (accepted) {
      x = std::move(x_new);
      fx = f_new;
      if (rel_decrease < params.relative_decrease_tolerance) {
        out.status = TNTStatus::RelativeDecrease;     // gradient norms stay those of the previous iterate
        break;
      }
      if (h_norm < params.stepsize_tolerance) {
        out.status = TNTStatus::Stepsize;
        break;
      }
      QM(x, grad, Hess, args...);
      H = Views::hessian(x, Hess);
      norms(gnorm, pgnorm);
    }

    if (!std::isnan(rho) && rho >= params.eta2) {
      Delta = std::max<Scalar>(params.alpha2 * h_M, Delta);
    } else if (std::isnan(rho) || rho < params.eta1) {
      Delta = params.alpha1 * h_M;
      if (Delta < params.Delta_tolerance) {
        out.status = TNTStatus::TrustRegion;
        break;
      }
    }
  }

  out.elapsed_time = Stopwatch::tock(t0);
  out.x = x;
  out.f = fx;
  out.gradfx_norm = gnorm;
  out.preconditioned_grad_f_x_norm = pgnorm;
  record(out.elapsed_time);   // one trailing trace entry, as the reference (TNT.h:617-621)

  if (params.verbose) {
    static const char *why[] = {"gradient tolerance", "preconditioned gradient tolerance", "relative decrease",
                                "step size",          "trust-region radius",               "iteration limit",
                                "elapsed time",       "user function"};
    std::cout << std::endl
              << "TNT finished: " << why[static_cast<int>(out.status)] << "; f = " << out.f
              << ", |g| = " << out.gradfx_norm << ", " << out.inner_iterations.size() << " outer iterations, "
              << out.elapsed_time << " s" << std::endl;
    std::cout << std::defaultfloat << std::setprecision(6);
  }
  return out;
}

// Gradient + Hessian-constructor form (reference TNT.h:704-718).
template <typename Variable, typename Tangent, typename Scalar = double, typename... Args>
TNTResult<Variable, Scalar>
TNT(const Objective<Variable, Scalar, Args...> &f, const VectorField<Variable, Tangent, Args...> &grad_f,
    const LinearOperatorConstructor<Variable, Tangent, Args...> &HC,
    const RiemannianMetric<Variable, Tangent, Scalar, Args...> &metric,
    const Retraction<Variable, Tangent, Args...> &retract, const Variable &x0, Args &...args,
    const std::optional<LinearOperator<Variable, Tangent, Args...>> &precon = std::nullopt,
    const TNTParams<Scalar> &params = TNTParams<Scalar>(),
    const std::optional<TNTUserFunction<Variable, Tangent, Scalar, Args...>> &user_function = std::nullopt) {
  QuadraticModel<Variable, Tangent, Args...> QM = [&grad_f, &HC](const Variable &X, Tangent &g,
                                                                LinearOperator<Variable, Tangent, Args...> &Hs,
                                                                Args &...a) {
    g = grad_f(X, a...);
    Hs = HC(X, a...);
  };
  return TNT<Variable, Tangent, Scalar, Args...>(f, QM, metric, retract, x0, args..., precon, params, user_function);
}

// Euclidean conveniences (reference TNT.h:753-805): standard inner product (`Vector::dot`), retraction X + V.
template <typename Vector, typename Scalar = double, typename... Args>
using EuclideanTNTUserFunction = TNTUserFunction<Vector, Vector, Scalar, Args...>;

template <typename Vector, typename Scalar = double, typename... Args>
TNTResult<Vector, Scalar>
EuclideanTNT(const Objective<Vector, Scalar, Args...> &f, const EuclideanQuadraticModel<Vector, Args...> &QM,
             const Vector &x0, Args &...args,
             const std::optional<EuclideanLinearOperator<Vector, Args...>> &precon = std::nullopt,
             const TNTParams<Scalar> &params = TNTParams<Scalar>(),
             const std::optional<EuclideanTNTUserFunction<Vector, Scalar, Args...>> &user_function = std::nullopt) {
  return TNT<Vector, Vector, Scalar, Args...>(f, QM, EuclideanMetric<Vector, Scalar, Args...>,
                                              EuclideanRetraction<Vector, Args...>, x0, args..., precon, params,
                                              user_function);
}

// gradient + Hessian-constructor form
template <typename Vector, typename Scalar = double, typename... Args>
TNTResult<Vector, Scalar>
EuclideanTNT(const Objective<Vector, Scalar, Args...> &f, const EuclideanVectorField<Vector, Args...> &nabla_f,
             const EuclideanLinearOperatorConstructor<Vector, Args...> &HessianConstructor, const Vector &x0,
             Args &...args, const std::optional<EuclideanLinearOperator<Vector, Args...>> &precon = std::nullopt,
             const TNTParams<Scalar> &params = TNTParams<Scalar>(),
             const std::optional<EuclideanTNTUserFunction<Vector, Scalar, Args...>> &user_function = std::nullopt) {
  EuclideanQuadraticModel<Vector, Args...> QM = [&nabla_f, &HessianConstructor](
                                                    const Vector &X, Vector &grad,
                                                    EuclideanLinearOperator<Vector, Args...> &Hess, Args &...a) {
    grad = nabla_f(X, a...);
    Hess = HessianConstructor(X, a...);
  };
  return EuclideanTNT<Vector, Scalar, Args...>(f, QM, x0, args..., precon, params, user_function);
}

}  // namespace Riemannian
}  // namespace Optimization
