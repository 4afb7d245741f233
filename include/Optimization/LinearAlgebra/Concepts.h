// Drop-in header layer (reference: include/Optimization/LinearAlgebra/Concepts.h:16-26).
#pragma once
#include <functional>

namespace Optimization {
namespace LinearAlgebra {

template <typename VectorX, typename VectorY, typename... Args>
using LinearOperator = std::function<VectorY(const VectorX &x, Args &...args)>;

template <typename Vector, typename... Args>
using SymmetricLinearOperator = LinearOperator<Vector, Vector, Args...>;

template <typename Vector, typename Scalar = double, typename... Args>
using InnerProduct = std::function<Scalar(const Vector &x, const Vector &y, Args &...args)>;

}  // namespace LinearAlgebra
}  // namespace Optimization
