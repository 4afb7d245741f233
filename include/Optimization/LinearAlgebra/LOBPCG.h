// Drop-in header layer: LOBPCG entry point with the reference's signature (reference:
// include/Optimization/LinearAlgebra/LOBPCG.h:101-140), written from scratch around the B200 block kernels.
// The reference implements the method on Eigen matrices (its `Matrix` parameter must provide Eigen's API and its
// Rayleigh-Ritz step uses Eigen's GeneralizedSelfAdjointEigenSolver); this layer provides the DEVICE path:
// Matrix = Optimization::b200::DeviceMatrix (row-major m x nx block vector in HBM), Vector = std::vector<double>,
// operators A, B, T given as block-operator descriptor functors (b200::BlockOperator, see
// Optimization/b200/Device.h), dispatched to ob200_lobpcg.  Other instantiations throw std::invalid_argument.
#pragma once
#include <optional>
#include <stdexcept>
#include <utility>

#include "Optimization/LinearAlgebra/Concepts.h"

namespace Optimization {
namespace LinearAlgebra {

// Per-iteration hook of the reference (i, A, B, T, nev, Theta, X, residual norms, nc); returning true stops the
// iteration.  The device path runs the whole method behind the C ABI and does not call back: passing a
// user_function makes the call throw.
template <typename Vector, typename Matrix, typename Scalar = double, typename... Args>
using LOBPCGUserFunction =
    std::function<bool(size_t i, const SymmetricLinearOperator<Matrix, Args...> &A,
                       const std::optional<SymmetricLinearOperator<Matrix, Args...>> &B,
                       const std::optional<SymmetricLinearOperator<Matrix, Args...>> &T, size_t nev, const Vector &Theta,
                       const Matrix &X, const Vector &residuals, size_t nc, Args &...args)>;

namespace detail {
// Customisation point: device layers specialise it; the primary template declines.
template <typename Vector, typename Matrix, typename Scalar, typename... Args>
struct DeviceLOBPCG {
  static bool run(const SymmetricLinearOperator<Matrix, Args...> &, const std::optional<SymmetricLinearOperator<Matrix, Args...>> &,
                  const std::optional<SymmetricLinearOperator<Matrix, Args...>> &, const Matrix &, size_t, size_t, size_t &,
                  size_t &, Scalar, std::pair<Vector, Matrix> &) {
    return false;
  }
};
}  // namespace detail

// Smallest nev eigenpairs (Theta, X) of A x = lambda B x from the initial block X0 (m x nx): returns the eigenvalue
// estimates and the B-orthonormal eigenvector estimates; num_iters / nc report iterations and converged pairs.
template <typename Vector, typename Matrix, typename Scalar = double, typename... Args>
std::pair<Vector, Matrix>
LOBPCG(const SymmetricLinearOperator<Matrix, Args...> &A, const std::optional<SymmetricLinearOperator<Matrix, Args...>> &B,
       const std::optional<SymmetricLinearOperator<Matrix, Args...>> &T, const Matrix &X0, size_t nev, size_t max_iters,
       size_t &num_iters, size_t &nc, Args &...args, Scalar tau = 1e-6,
       const std::optional<LOBPCGUserFunction<Vector, Matrix, Scalar, Args...>> &user_function = std::nullopt) {
  if (user_function)
    throw std::invalid_argument("LOBPCG: the device path runs behind the C ABI and does not call a per-iteration user function");
  std::pair<Vector, Matrix> out;
  if (detail::DeviceLOBPCG<Vector, Matrix, Scalar, Args...>::run(A, B, T, X0, nev, max_iters, num_iters, nc, tau, out))
    return out;
  throw std::invalid_argument(
      "LOBPCG: this layer provides the device path only (Matrix = Optimization::b200::DeviceMatrix with "
      "b200::BlockOperator functors); the reference's dense path needs Eigen");
}

}  // namespace LinearAlgebra
}  // namespace Optimization
