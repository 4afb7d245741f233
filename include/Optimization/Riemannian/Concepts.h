// Drop-in header layer: the functor surface of the Riemannian solvers
// (reference: include/Optimization/Riemannian/Concepts.h:44-190).
#pragma once
#include <functional>
#include <utility>

#include "Optimization/Base/Concepts.h"

namespace Optimization {
namespace Riemannian {

template <typename Variable, typename Tangent, typename... Args>
using VectorField = std::function<Tangent(const Variable &X, Args &...args)>;

template <typename Variable, typename Tangent, typename... Args>
using LinearOperator = std::function<Tangent(const Variable &X, const Tangent &V, Args &...args)>;

template <typename Variable, typename Tangent, typename... Args>
using LinearOperatorConstructor =
    std::function<LinearOperator<Variable, Tangent, Args...>(const Variable &X, Args &...args)>;

template <typename Variable, typename Tangent, typename... Args>
using QuadraticModel = std::function<void(const Variable &X, Tangent &gradient,
                                          LinearOperator<Variable, Tangent, Args...> &Hessian, Args &...args)>;

template <typename VariableX, typename VariableY, typename... Args>
using Mapping = std::function<VariableY(const VariableX &X, Args &...args)>;

template <typename VariableX, typename TangentX, typename TangentY, typename... Args>
using Jacobian = std::function<TangentY(const VariableX &X, const TangentX &V, Args &...args)>;

template <typename VariableX, typename TangentX, typename TangentY, typename... Args>
using JacobianAdjoint = std::function<TangentX(const VariableX &X, const TangentY &W, Args &...args)>;

template <typename VariableX, typename TangentX, typename TangentY, typename... Args>
using JacobianPairFunction =
    std::function<std::pair<Jacobian<VariableX, TangentX, TangentY, Args...>,
                            JacobianAdjoint<VariableX, TangentX, TangentY, Args...>>(const VariableX &X, Args &...args)>;

template <typename Variable, typename Tangent, typename Scalar = double, typename... Args>
using RiemannianMetric = std::function<Scalar(const Variable &X, const Tangent &V1, const Tangent &V2, Args &...args)>;

template <typename Variable, typename Tangent, typename... Args>
using Retraction = std::function<Variable(const Variable &X, const Tangent &update, Args &...args)>;

template <typename Scalar = double>
struct SmoothOptimizerParams : public OptimizerParams {
  Scalar gradient_tolerance = 1e-6;
  Scalar relative_decrease_tolerance = 1e-6;
  Scalar stepsize_tolerance = 1e-6;
};

template <typename Variable, typename Scalar = double>
struct SmoothOptimizerResult : public OptimizerResult<Variable, Scalar> {
  Scalar gradfx_norm;
  std::vector<Scalar> gradient_norms;
  std::vector<Scalar> update_step_norms;
};

// Euclidean specialisations (reference Riemannian/Concepts.h:162-190): one type `Vector` for points and tangent
// vectors; `Vector` provides `double dot(const Vector &) const`.
template <typename Vector, typename... Args>
using EuclideanVectorField = VectorField<Vector, Vector, Args...>;

template <typename Vector, typename... Args>
using EuclideanLinearOperator = LinearOperator<Vector, Vector, Args...>;

template <typename Vector, typename... Args>
using EuclideanLinearOperatorConstructor = LinearOperatorConstructor<Vector, Vector, Args...>;

template <typename Vector, typename... Args>
using EuclideanQuadraticModel = QuadraticModel<Vector, Vector, Args...>;

template <typename Vector, typename Scalar = double, typename... Args>
Scalar EuclideanInnerProduct(const Vector &V1, const Vector &V2, Args &...) {
  return V1.dot(V2);
}

template <typename Vector, typename Scalar = double, typename... Args>
Scalar EuclideanMetric(const Vector &, const Vector &V1, const Vector &V2, Args &...args) {
  return EuclideanInnerProduct<Vector, Scalar, Args...>(V1, V2, args...);
}

template <typename Vector, typename... Args>
Vector EuclideanRetraction(const Vector &X, const Vector &V, Args &...) {
  return X + V;
}

}  // namespace Riemannian
}  // namespace Optimization
