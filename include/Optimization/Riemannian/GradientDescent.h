// Drop-in header layer: Riemannian gradient descent with Armijo backtracking, with the reference's entry
// points, parameter / result types and stopping semantics (reference:
// include/Optimization/Riemannian/GradientDescent.h:36-436), written from scratch.  Host control flow only:
// every vector statement goes through the operators of the Tangent / Variable types, so with
// Optimization::b200::DeviceMatrix the arithmetic runs in the level-1 CUDA kernels of the C ABI
// (ob200_axpby, ob200_dot) and the model / retraction entry points.
#pragma once
#include <cmath>
#include <iomanip>
#include <iostream>
#include <limits>
#include <optional>
#include <stdexcept>
#include <vector>

#include "Optimization/Riemannian/Concepts.h"
#include "Optimization/Util/Stopwatch.h"

namespace Optimization {
namespace Riemannian {

// Called once per ACCEPTED iteration with the state before the update is applied
// (i, elapsed time, x, f(x), grad f(x), accepted step h, decrease df).
template <typename Variable, typename Tangent, typename Scalar = double, typename... Args>
using GradientDescentUserFunction = std::function<void(size_t i, double t, const Variable &x, Scalar f, const Tangent &g,
                                                       const Tangent &h, Scalar df, Args &...args)>;

template <typename Scalar = double>
struct GradientDescentParams : public SmoothOptimizerParams<Scalar> {
  Scalar alpha = 1.0;              // first trial stepsize of the Armijo search (> 0)
  Scalar beta = .5;                // backtracking shrink factor, in (0, 1)
  Scalar sigma = .5;               // sufficient-decrease fraction, in (0, 1)
  size_t max_ls_iterations = 100;  // trial stepsizes per iteration
};

enum class GradientDescentStatus {
  Gradient,          // gradient norm below gradient_tolerance
  RelativeDecrease,  // last accepted step decreased f by less than relative_decrease_tolerance (relative)
  Stepsize,          // last accepted step shorter than stepsize_tolerance
  LineSearch,        // no trial stepsize gave sufficient decrease
  IterationLimit,
  ElapsedTime
};

template <typename Variable, typename Scalar = double>
struct GradientDescentResult : public SmoothOptimizerResult<Variable, Scalar> {
  GradientDescentStatus status;
  std::vector<size_t> linesearch_iterations;   // trial stepsizes used by each accepted iteration
};

namespace detail {
// closing report of a verbose run: one line naming the stopping rule that fired and the quantity it tested
template <typename Scalar>
void report_gradient_descent(GradientDescentStatus why, Scalar gnorm, Scalar rel_decrease, Scalar step_norm, Scalar f,
                             double seconds, const GradientDescentParams<Scalar> &prm) {
  std::cout << "\n\nGradient descent stopped: ";
  switch (why) {
    case GradientDescentStatus::Gradient: std::cout << "gradient norm " << gnorm << " below " << prm.gradient_tolerance; break;
    case GradientDescentStatus::RelativeDecrease:
      std::cout << "relative decrease " << rel_decrease << " below " << prm.relative_decrease_tolerance;
      break;
    case GradientDescentStatus::Stepsize: std::cout << "step length " << step_norm << " below " << prm.stepsize_tolerance; break;
    case GradientDescentStatus::LineSearch:
      std::cout << "no sufficient decrease within " << prm.max_ls_iterations << " trial stepsizes";
      break;
    case GradientDescentStatus::IterationLimit: std::cout << "iteration limit " << prm.max_iterations << " reached"; break;
    case GradientDescentStatus::ElapsedTime: std::cout << "time limit " << prm.max_computation_time << " s exceeded"; break;
  }
  std::cout << "\n  f = " << f << ", |grad f| = " << gnorm << ", " << seconds << " s\n" << std::endl;
}
}  // namespace detail

template <typename Variable, typename Tangent, typename Scalar = double, typename... Args>
GradientDescentResult<Variable, Scalar>
GradientDescent(const Objective<Variable, Scalar, Args...> &f, const VectorField<Variable, Tangent, Args...> &grad_f,
                const RiemannianMetric<Variable, Tangent, Scalar, Args...> &metric,
                const Retraction<Variable, Tangent, Args...> &retract, const Variable &x0, Args &...args,
                const GradientDescentParams<Scalar> &params = GradientDescentParams<Scalar>(),
                const std::optional<GradientDescentUserFunction<Variable, Tangent, Scalar, Args...>> &user_function =
                    std::nullopt) {
  // the reference checks exactly these five (its GradientDescent.h:141-161)
  if (params.max_computation_time < 0)
    throw std::invalid_argument("Maximum computation time must be a nonnegative real value");
  if (params.gradient_tolerance < 0) throw std::invalid_argument("Gradient tolerance must be a nonnegative real value");
  if (params.alpha <= 0)
    throw std::invalid_argument("Initial stepsize for backtracking line-search must be a positive real value");
  if (params.beta <= 0 || params.beta >= 1)
    throw std::invalid_argument("Multiplicative shrinkage factor for stepsize in backtracking line-search must be a "
                                "value in the range (0, 1)");
  if (params.sigma <= 0 || params.sigma >= 1)
    throw std::invalid_argument("Sufficient fractional decrease parameter for step acceptance in backtracking line "
                                "search must be a value in the range (0, 1)");

  const Scalar sqrt_eps = std::sqrt(std::numeric_limits<Scalar>::epsilon());
  const bool talk = params.verbose;
  const int it_width = int(std::floor(std::log10(double(params.max_iterations)))) + 1;
  const int ls_width = int(std::floor(std::log10(double(params.max_ls_iterations)))) + 1;

  GradientDescentResult<Variable, Scalar> out;
  out.status = GradientDescentStatus::IterationLimit;

  // state of the current iterate
  Variable x = x0;
  Scalar fx = f(x, args...);
  Tangent g = grad_f(x, args...);
  Scalar gnorm = std::sqrt(metric(x, g, g, args...));
  Scalar rel_decrease = 0, step_norm = 0;

  if (talk) {
    std::cout << std::scientific << std::setprecision(int(params.precision));
    std::cout << "Riemannian gradient descent (Armijo backtracking)\n" << std::endl;
  }

  const auto t0 = Stopwatch::tick();
  for (size_t it = 0; it < params.max_iterations; ++it) {
    const double now = Stopwatch::tock(t0);
    if (now > params.max_computation_time) {
      out.status = GradientDescentStatus::ElapsedTime;
      break;
    }
    out.time.push_back(now);
    out.objective_values.push_back(fx);
    out.gradient_norms.push_back(gnorm);
    if (params.log_iterates) out.iterates.push_back(x);
    if (talk)
      std::cout << "[" << std::setw(it_width) << it << "] t " << now << "  f " << std::setw(int(params.precision) + 7) << fx
                << "  |grad| " << gnorm;

    if (gnorm < params.gradient_tolerance) {
      out.status = GradientDescentStatus::Gradient;
      break;
    }

    // Armijo backtracking along -grad: t = alpha, alpha beta, alpha beta^2, ...  (the stepsize is formed by
    // repeated multiplication starting from alpha / beta, as the reference does, so trial values agree bit for bit)
    Scalar t = params.alpha / params.beta;
    size_t trials = 0;
    bool accepted = false;
    Tangent h;
    Variable x_trial;
    Scalar f_trial = fx, df = 0;
    do {   // at least one trial even when max_ls_iterations == 0, like the reference (GradientDescent.h:270-286)
      ++trials;
      t *= params.beta;
      h = -t * g;
      x_trial = retract(x, h, args...);
      f_trial = f(x_trial, args...);
      df = fx - f_trial;
      accepted = df > params.sigma * t * gnorm * gnorm;
    } while (!accepted && trials < params.max_ls_iterations);
    if (talk) std::cout << "  trials " << std::setw(ls_width) << trials;
    if (!accepted) {
      out.status = GradientDescentStatus::LineSearch;
      break;
    }

    step_norm = t * gnorm;
    rel_decrease = df / (std::fabs(fx) + sqrt_eps);
    out.linesearch_iterations.push_back(trials);
    out.update_step_norms.push_back(step_norm);
    if (user_function) (*user_function)(it, now, x, fx, g, h, df, args...);
    if (talk) std::cout << "  |step| " << step_norm << "  decrease " << df;

    // move
    x = x_trial;
    fx = f_trial;
    g = grad_f(x, args...);
    gnorm = std::sqrt(metric(x, g, g, args...));

    if (rel_decrease < params.relative_decrease_tolerance) {
      out.status = GradientDescentStatus::RelativeDecrease;
      break;
    }
    if (step_norm < params.stepsize_tolerance) {
      out.status = GradientDescentStatus::Stepsize;
      break;
    }
    if (talk) std::cout << std::endl;
  }

  out.elapsed_time = Stopwatch::tock(t0);
  out.x = x;
  out.f = fx;
  out.gradfx_norm = gnorm;

  if (talk) detail::report_gradient_descent(out.status, gnorm, rel_decrease, step_norm, out.f, out.elapsed_time, params);
  return out;
}

// Euclidean convenience (reference GradientDescent.h:415-433): standard inner product, retraction X + V.
template <typename Vector, typename Scalar = double, typename... Args>
using EuclideanGradientDescentUserFunction = GradientDescentUserFunction<Vector, Vector, Scalar, Args...>;

template <typename Vector, typename Scalar = double, typename... Args>
GradientDescentResult<Vector, Scalar>
EuclideanGradientDescent(const Objective<Vector, Scalar, Args...> &f, const EuclideanVectorField<Vector, Args...> grad_f,
                         const Vector &x0, Args &...args,
                         const GradientDescentParams<Scalar> &params = GradientDescentParams<Scalar>(),
                         const std::optional<EuclideanGradientDescentUserFunction<Vector, Scalar, Args...>> &user_function =
                             std::nullopt) {
  return GradientDescent<Vector, Vector, Scalar, Args...>(f, grad_f, EuclideanMetric<Vector, Scalar, Args...>,
                                                          EuclideanRetraction<Vector, Args...>, x0, args..., params,
                                                          user_function);
}

}  // namespace Riemannian
}  // namespace Optimization
