// TEST INFRASTRUCTURE ONLY.  Dense LU with partial pivoting for the small KKT systems of the
// constraint-preconditioned STPCG tests (the reference's tests use Eigen::UmfPackLU, absent here:
// tests/IterativeSolvers_unit_test.cpp:316-496).  Included by oracle/ref_driver.cpp and by
// tests/host/projected_host_check.cpp so that both sides solve the KKT systems with identical arithmetic.
#pragma once
#include <cmath>
#include <cstddef>
#include <utility>
#include <vector>

namespace oracle {
struct DenseLU {
  size_t n = 0;
  std::vector<double> a;     // row-major n x n, L (unit diagonal) and U in place
  std::vector<size_t> piv;
  bool factor(const std::vector<double> &K, size_t dim) {
    n = dim;
    a = K;
    piv.resize(n);
    for (size_t c = 0; c < n; ++c) {
      size_t best = c;
      for (size_t r = c + 1; r < n; ++r)
        if (std::fabs(a[r * n + c]) > std::fabs(a[best * n + c])) best = r;
      piv[c] = best;
      if (a[best * n + c] == 0.0) return false;
      if (best != c)
        for (size_t j = 0; j < n; ++j) std::swap(a[c * n + j], a[best * n + j]);
      for (size_t r = c + 1; r < n; ++r) {
        const double l = a[r * n + c] / a[c * n + c];
        a[r * n + c] = l;
        for (size_t j = c + 1; j < n; ++j) a[r * n + j] -= l * a[c * n + j];
      }
    }
    return true;
  }
  std::vector<double> solve(std::vector<double> b) const {
    for (size_t c = 0; c < n; ++c) {
      if (piv[c] != c) std::swap(b[c], b[piv[c]]);
      for (size_t r = c + 1; r < n; ++r) b[r] -= a[r * n + c] * b[c];
    }
    for (size_t i = n; i-- > 0;) {
      double s = b[i];
      for (size_t j = i + 1; j < n; ++j) s -= a[i * n + j] * b[j];
      b[i] = s / a[i * n + i];
    }
    return b;
  }
};
// KKT matrix [diag(m) A^T; A 0], A: mc x n row-major
inline std::vector<double> kkt_matrix(const double *mdiag, const double *A, size_t n, size_t mc) {
  const size_t N = n + mc;
  std::vector<double> K(N * N, 0.0);
  for (size_t i = 0; i < n; ++i) K[i * N + i] = mdiag[i];
  for (size_t c = 0; c < mc; ++c)
    for (size_t i = 0; i < n; ++i) {
      K[i * N + n + c] = A[c * n + i];
      K[(n + c) * N + i] = A[c * n + i];
    }
  return K;
}
}  // namespace oracle
