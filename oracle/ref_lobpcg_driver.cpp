// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// oracle/_ref/libref_lobpcg.so: the UNMODIFIED reference header
//   /root/reference/include/Optimization/LinearAlgebra/LOBPCG.h (RayleighRitz l.53-62, LOBPCG l.131-337)
// included where it lies and compiled against oracle/eigen_shim (a stand-in for the Eigen API the header uses; Eigen
// is absent from this image).  The operators are the test problems of the reference's tests/LOBPCG_unit_test.cpp
// (diagonal A, B, T) and the 7-point Laplacian of BASELINE config C4.  Plain C ABI for ctypes.
#include "Optimization/LinearAlgebra/LOBPCG.h"

#include <cstring>
#include <random>

using Eigen::Matrix;
using Eigen::Vector;
using namespace Optimization::LinearAlgebra;

namespace {
// block operators on m x k column-major matrices
struct Op {
  int kind = 0;              // 0 none, 1 diagonal, 2 scalar, 3 stencil7
  const double *diag = nullptr;
  double alpha = 0;
  uint32_t gx = 0, gy = 0, gz = 0;
  Matrix apply(const Matrix &X) const {
    Matrix Y(X.rows(), X.cols());
    for (size_t j = 0; j < X.cols(); ++j) {
      const double *x = X.col(j);
      double *y = Y.col(j);
      if (kind == 1) for (size_t i = 0; i < X.rows(); ++i) y[i] = diag[i] * x[i];
      else if (kind == 2) for (size_t i = 0; i < X.rows(); ++i) y[i] = alpha * x[i];
      else {
        const size_t sx = 1, sy = gx, sz = size_t(gx) * gy;
        for (uint32_t z = 0; z < gz; ++z) for (uint32_t yy = 0; yy < gy; ++yy) for (uint32_t xx = 0; xx < gx; ++xx) {
          const size_t o = (size_t(z) * gy + yy) * gx + xx;
          double h = 6.0 * x[o];
          if (xx > 0) h -= x[o - sx];
          if (xx + 1 < gx) h -= x[o + sx];
          if (yy > 0) h -= x[o - sy];
          if (yy + 1 < gy) h -= x[o + sy];
          if (z > 0) h -= x[o - sz];
          if (z + 1 < gz) h -= x[o + sz];
          y[o] = h;
        }
      }
    }
    return Y;
  }
};
Op make_op(int kind, const double *diag, double alpha, const uint32_t *grid) {
  Op o; o.kind = kind; o.diag = diag; o.alpha = alpha;
  if (grid) { o.gx = grid[0]; o.gy = grid[1]; o.gz = grid[2]; }
  return o;
}
}  // namespace

extern "C" {
// The Gaussian probe block the reference draws for its operator-norm estimates (LOBPCG.h:203-210): same generator,
// same order -- row-major m x nx on return (what ob200_lobpcg takes as Omega).
void ref_lobpcg_omega(uint64_t m, uint64_t nx, double *out_rowmajor) {
  std::default_random_engine gen;
  std::normal_distribution<double> normal(0, 1.0);
  for (size_t i = 0; i < m; ++i)
    for (size_t j = 0; j < nx; ++j) out_rowmajor[i * nx + j] = normal(gen);
}

// kinds: 0 = absent, 1 = diagonal (diag array), 2 = scalar (alpha), 3 = 7-point Laplacian (grid)
// X0 / X_out: row-major m x nx / m x nev.  Returns 0, or 1 on std::invalid_argument, 2 on another exception.
int ref_lobpcg(int kindA, const double *diagA, double alphaA, const uint32_t *gridA, int kindB, const double *diagB,
               double alphaB, int kindT, const double *diagT, double alphaT, uint64_t m, uint64_t nx, const double *X0,
               uint64_t nev, uint64_t max_iters, double tau, double *theta_out, double *X_out, uint64_t *num_iters,
               uint64_t *nc, double *resid_trace /* nullable: max_iters x nx residual norms */) {
  const Op a = make_op(kindA, diagA, alphaA, gridA), b = make_op(kindB, diagB, alphaB, nullptr),
           t = make_op(kindT, diagT, alphaT, nullptr);
  SymmetricLinearOperator<Matrix> A = [&a](const Matrix &X) { return a.apply(X); };
  std::optional<SymmetricLinearOperator<Matrix>> B, T;
  if (kindB) B = [&b](const Matrix &X) { return b.apply(X); };
  if (kindT) T = [&t](const Matrix &X) { return t.apply(X); };
  Matrix X(m, nx);
  for (size_t i = 0; i < m; ++i) for (size_t j = 0; j < nx; ++j) X(i, j) = X0[i * nx + j];
  std::optional<LOBPCGUserFunction<Vector, Matrix>> user;
  if (resid_trace)
    user = [&](size_t i, const SymmetricLinearOperator<Matrix> &, const std::optional<SymmetricLinearOperator<Matrix>> &,
               const std::optional<SymmetricLinearOperator<Matrix>> &, size_t, const Vector &, const Matrix &,
               const Vector &r, size_t) {
      if (i < max_iters) for (size_t j = 0; j < nx; ++j) resid_trace[i * nx + j] = r(j);
      return false;
    };
  try {
    size_t it = 0, conv = 0;
    auto out = LOBPCG<Vector, Matrix>(A, B, T, X, size_t(nev), size_t(max_iters), it, conv, tau, user);
    for (size_t j = 0; j < nev; ++j) theta_out[j] = out.first(j);
    for (size_t i = 0; i < m; ++i) for (size_t j = 0; j < nev; ++j) X_out[i * nev + j] = out.second(i, j);
    *num_iters = it;
    *nc = conv;
  } catch (const std::invalid_argument &) {
    return 1;
  } catch (const std::exception &) {
    return 2;
  }
  return 0;
}
}
