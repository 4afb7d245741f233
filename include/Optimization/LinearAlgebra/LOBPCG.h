// Drop-in header layer: locally optimal block preconditioned conjugate gradients with the reference's entry points
// (reference: include/Optimization/LinearAlgebra/LOBPCG.h -- RayleighRitz l.53-62, LOBPCGUserFunction l.86-94, LOBPCG
// l.131-337, the random-start overload l.376-390), written from scratch.  Two execution paths:
//   * device: Matrix = Optimization::b200::DeviceMatrix (row-major m x nx block vector in HBM), Vector =
//     std::vector<double>, operators given as block-operator descriptor functors (b200::BlockOperator, see
//     Optimization/b200/Device.h): the whole method runs behind the C ABI (ob200_lobpcg) on the B200 block kernels.
//     That path cannot call back into host code: a per-iteration user_function makes it throw.
//   * dense: any Matrix / Vector pair with the subset of the Eigen dense API listed under `dense_api` below (Eigen
//     itself, or the stand-in the test-suite compiles the reference against).  Same iteration as the reference --
//     soft locking of the converged leading columns, Rayleigh-Ritz with diagonal equilibration of the B-Gram, the
//     backward-stable convergence test with Gaussian-probe operator-norm estimates -- with the per-iteration hook.
//     It is only compiled when an <Eigen/Eigenvalues> header is on the include path.
#pragma once
#include <cstddef>
#include <functional>
#include <optional>
#include <random>
#include <stdexcept>
#include <tuple>
#include <type_traits>
#include <utility>

#if defined(__has_include)
#if __has_include(<Eigen/Eigenvalues>)
#include <Eigen/Core>
#include <Eigen/Eigenvalues>
#define OPTIMIZATION_B200_HAVE_DENSE_EIGENSOLVER 1
#endif
#endif

#include "Optimization/LinearAlgebra/Concepts.h"

namespace Optimization {
namespace LinearAlgebra {

// Per-iteration hook (i, A, B, T, nev, Theta, X, residual norms, nc); called after the Ritz pairs and residuals of the
// iteration are available; returning true ends the iteration.
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

// dense_api<Matrix>: the matrix type offers column blocks, transposition and column norms the way Eigen does
template <typename Matrix, typename = void>
struct dense_api : std::false_type {};
template <typename Matrix>
struct dense_api<Matrix, std::void_t<decltype(std::declval<Matrix &>().leftCols(size_t(0))),
                                     decltype(std::declval<const Matrix &>().transpose()),
                                     decltype(std::declval<const Matrix &>().colwise())>> : std::true_type {};
}  // namespace detail

#ifdef OPTIMIZATION_B200_HAVE_DENSE_EIGENSOLVER
// Rayleigh-Ritz on the pencil (A, B), B positive definite: Theta ascending and C with C'AC = diag(Theta), C'BC = I.
// The pencil is first scaled symmetrically so that B has a unit diagonal.
template <typename Vector, typename Matrix>
std::pair<Vector, Matrix> RayleighRitz(const Matrix &A, const Matrix &B) {
  const Vector scale = B.diagonal().cwiseSqrt().cwiseInverse();
  const Eigen::GeneralizedSelfAdjointEigenSolver<Matrix> pencil(scale.asDiagonal() * A * scale.asDiagonal(),
                                                               scale.asDiagonal() * B * scale.asDiagonal());
  return std::make_pair(pencil.eigenvalues(), scale.asDiagonal() * pencil.eigenvectors());
}

namespace detail {
template <typename Vector, typename Matrix, typename Scalar, typename... Args>
std::pair<Vector, Matrix>
dense_lobpcg(const SymmetricLinearOperator<Matrix, Args...> &A, const std::optional<SymmetricLinearOperator<Matrix, Args...>> &B,
             const std::optional<SymmetricLinearOperator<Matrix, Args...>> &T, const Matrix &X0, size_t nev, size_t max_iters,
             size_t &num_iters, size_t &nc, Args &...args, Scalar tau,
             const std::optional<LOBPCGUserFunction<Vector, Matrix, Scalar, Args...>> &user_function) {
  const size_t m = X0.rows(), nx = X0.cols();
  auto applyB = [&](const Matrix &V) -> Matrix { return B ? (*B)(V, args...) : V; };

  // operator 2-norm estimates from one Gaussian probe block (drawn row by row from the default engine)
  Scalar normA, normB;
  {
    std::default_random_engine engine;
    std::normal_distribution<Scalar> gauss(0, 1.0);
    Matrix probe(m, nx);
    for (size_t i = 0; i < m; ++i)
      for (size_t j = 0; j < nx; ++j) probe(i, j) = gauss(engine);
    normA = A(probe, args...).norm() / probe.norm();
    normB = B ? (*B)(probe, args...).norm() / probe.norm() : Scalar(1.0);
  }

  // start: B-orthonormalise the block by a Rayleigh-Ritz step on span(X0)
  Matrix X = X0, AX = A(X, args...), BX = applyB(X);
  Vector Theta;
  Matrix C;
  std::tie(Theta, C) = RayleighRitz<Vector, Matrix>(X.transpose() * AX, X.transpose() * BX);
  AX = AX * C;
  BX = BX * C;
  Matrix R = AX - BX * Theta.asDiagonal();
  Matrix P, basis(m, 3 * nx);
  Vector resid;
  nc = 0;

  for (num_iters = 1; num_iters < max_iters; ++num_iters) {
    // search space [X, W_active, P_active]: the nc converged leading columns stay in X only (soft locking)
    const size_t active = nx - nc;
    const Matrix W = T ? (*T)(R, args...) : R;
    basis.leftCols(nx) = X;
    basis.middleCols(nx, active) = W.rightCols(active);
    size_t ns = nx + active;
    if (num_iters > 1) {
      basis.middleCols(ns, active) = P.rightCols(active);
      ns += active;
    }
    const Matrix AS = A(basis.leftCols(ns), args...);
    const Matrix BS = B ? (*B)(basis.leftCols(ns), args...) : Matrix(basis.leftCols(ns));
    std::tie(Theta, C) = RayleighRitz<Vector, Matrix>(basis.leftCols(ns).transpose() * AS, basis.leftCols(ns).transpose() * BS);

    X = basis.leftCols(ns) * C.leftCols(nx);
    AX = A(X, args...);
    BX = applyB(X);
    R = AX - BX * Theta.head(nx).asDiagonal();
    P = basis.middleCols(nx, ns - nx) * C.bottomLeftCorner(ns - nx, nx);

    // backward-stable test per pair: |r_i| <= tau (|A| + |B| |theta_i|) |x_i|; pairs lock in order
    resid = R.colwise().norm();
    const Vector tol = tau * (normA + normB * Theta.head(nx).cwiseAbs().transpose().array()) * X.colwise().norm().array();
    const auto ok = (resid.head(nev).array() <= tol.head(nev).array());
    nc = 0;
    while (nc < nev && ok[nc]) ++nc;

    if (user_function && (*user_function)(num_iters, A, B, T, nev, Theta.head(nx), X, resid, nc, args...)) break;
    if (nc == nev) break;
  }
  Theta.conservativeResize(nev);
  X.conservativeResize(Eigen::NoChange, nev);
  return std::make_pair(Theta, X);
}
}  // namespace detail
#endif  // OPTIMIZATION_B200_HAVE_DENSE_EIGENSOLVER

// Smallest nev eigenpairs (Theta, X) of A x = lambda B x from the initial block X0 (m x nx): returns the eigenvalue
// estimates and the B-orthonormal eigenvector estimates (m x nev); num_iters / nc report iterations and converged pairs.
template <typename Vector, typename Matrix, typename Scalar = double, typename... Args>
std::pair<Vector, Matrix>
LOBPCG(const SymmetricLinearOperator<Matrix, Args...> &A, const std::optional<SymmetricLinearOperator<Matrix, Args...>> &B,
       const std::optional<SymmetricLinearOperator<Matrix, Args...>> &T, const Matrix &X0, size_t nev, size_t max_iters,
       size_t &num_iters, size_t &nc, Args &...args, Scalar tau = 1e-6,
       const std::optional<LOBPCGUserFunction<Vector, Matrix, Scalar, Args...>> &user_function = std::nullopt) {
  if (nev > size_t(X0.cols()))
    throw std::invalid_argument("Block size nx must be greater than or equal to the number nev of desired eigenpairs");
  if (size_t(X0.cols()) > size_t(X0.rows()))
    throw std::invalid_argument("Block size nx must be less than or equal to the dimension m of the problem");
  if constexpr (sizeof...(Args) == 0) {
    std::pair<Vector, Matrix> out;
    if (!user_function &&
        detail::DeviceLOBPCG<Vector, Matrix, Scalar>::run(A, B, T, X0, nev, max_iters, num_iters, nc, tau, out))
      return out;
  }
#ifdef OPTIMIZATION_B200_HAVE_DENSE_EIGENSOLVER
  if constexpr (detail::dense_api<Matrix>::value)
    return detail::dense_lobpcg<Vector, Matrix, Scalar, Args...>(A, B, T, X0, nev, max_iters, num_iters, nc, args..., tau,
                                                                 user_function);
#endif
  throw std::invalid_argument(
      "LOBPCG: no execution path for this instantiation -- the device path needs Matrix = Optimization::b200::DeviceMatrix "
      "with b200::BlockOperator functors and no user_function; the dense path needs an Eigen-style Matrix type and "
      "<Eigen/Eigenvalues> on the include path");
}

// Same with a random initial block (entries uniform in [-1, 1], Matrix::Random): reference l.376-390.
template <typename Vector, typename Matrix, typename Scalar = double, typename... Args>
std::pair<Vector, Matrix>
LOBPCG(const SymmetricLinearOperator<Matrix, Args...> &A, const std::optional<SymmetricLinearOperator<Matrix, Args...>> &B,
       const std::optional<SymmetricLinearOperator<Matrix, Args...>> &T, size_t m, size_t nx, size_t nev, size_t max_iters,
       size_t &num_iters, size_t &nc, Args &...args, Scalar tau = 1e-6,
       const std::optional<LOBPCGUserFunction<Vector, Matrix, Scalar, Args...>> &user_function = std::nullopt) {
  const Matrix X0 = Matrix::Random(m, nx);
  return LOBPCG<Vector, Matrix, Scalar, Args...>(A, B, T, X0, nev, max_iters, num_iters, nc, args..., tau, user_function);
}

}  // namespace LinearAlgebra
}  // namespace Optimization
