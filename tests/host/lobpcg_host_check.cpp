// CPU check of the drop-in header layer: OUR include/Optimization/LinearAlgebra/LOBPCG.h (dense path) compiled against
// the Eigen stand-in of the test-suite (oracle/eigen_shim), on the problems of the reference's tests/LOBPCG_unit_test.cpp
// and a small 3-D Laplacian.  Prints one JSON line per case; tests/test_headers.py compares them BIT FOR BIT with the
// reference's own header compiled against the same stand-in (oracle/_ref/libref_lobpcg.so).
// Input: binary file [u64 n][u64 nx][f64 adiag(n)][f64 bdiag(n)][f64 X0(n*nx) row-major].
#include <cstdio>
#include <string>
#include <vector>
#include "Optimization/LinearAlgebra/LOBPCG.h"

using Eigen::Matrix;
using Eigen::Vector;
using namespace Optimization::LinearAlgebra;
using Op = SymmetricLinearOperator<Matrix>;

static const char *g_dump = nullptr;   // argv[2]: prefix of the raw dumps of X (row-major doubles), one file per case
static void print_case(const char *name, const std::pair<Vector, Matrix> &res, size_t it, size_t nc, size_t hook_calls) {
  if (g_dump) {
    const std::string path = std::string(g_dump) + name + ".bin";
    if (FILE *o = fopen(path.c_str(), "wb")) {
      for (size_t i = 0; i < res.second.rows(); ++i)
        for (size_t j = 0; j < res.second.cols(); ++j) { const double v = res.second(i, j); fwrite(&v, 8, 1, o); }
      fclose(o);
    }
  }
  printf("{\"case\": \"%s\", \"num_iters\": %zu, \"nc\": %zu, \"hook_calls\": %zu, \"x_rows\": %zu, \"x_cols\": %zu, \"theta\": [",
         name, it, nc, hook_calls, res.second.rows(), res.second.cols());
  for (size_t i = 0; i < res.first.size(); ++i) printf("%s%.17g", i ? ", " : "", res.first(i));
  printf("], \"x_sum\": %.17g}\n", [&] { double s = 0; for (double v : res.second.d) s += v; return s; }());
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  if (argc > 2) g_dump = argv[2];
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  unsigned long long n = 0, nx = 0;
  if (fread(&n, 8, 1, f) != 1 || fread(&nx, 8, 1, f) != 1) return 2;
  std::vector<double> ad(n), bd(n), x0(n * nx);
  if (fread(ad.data(), 8, n, f) != n || fread(bd.data(), 8, n, f) != n || fread(x0.data(), 8, n * nx, f) != n * nx) return 2;
  fclose(f);
  Matrix X0(n, nx);
  for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < nx; ++j) X0(i, j) = x0[i * nx + j];
  auto diag_op = [n](const std::vector<double> &d, bool absolute) {
    return Op([n, &d, absolute](const Matrix &X) {
      Matrix Y(X.rows(), X.cols());
      for (size_t j = 0; j < X.cols(); ++j) for (size_t i = 0; i < n; ++i) Y(i, j) = (absolute && d[i] < 0 ? -d[i] : d[i]) * X(i, j);
      return Y;
    });
  };
  const Op A = diag_op(ad, false);
  const std::optional<Op> B(diag_op(bd, false)), T(diag_op(ad, true)), none;
  const char *names[4] = {"plain", "precon", "generalized_precon", "generalized"};
  for (int c = 0; c < 4; ++c) {
    size_t it = 0, nc = 0, calls = 0;
    const std::optional<LOBPCGUserFunction<Vector, Matrix>> hook(
        [&calls](size_t, const Op &, const std::optional<Op> &, const std::optional<Op> &, size_t, const Vector &, const Matrix &,
                 const Vector &, size_t) { ++calls; return false; });
    auto res = LOBPCG<Vector, Matrix>(A, c >= 2 ? B : none, (c == 1 || c == 2) ? T : none, X0, 5, 10 * n, it, nc, 1e-8, hook);
    print_case(names[c], res, it, nc, calls);
  }
  {   // the hook can stop the iteration (LOBPCG.h:318-320)
    size_t it = 0, nc = 0, calls = 0;
    const std::optional<LOBPCGUserFunction<Vector, Matrix>> stop(
        [&calls](size_t i, const Op &, const std::optional<Op> &, const std::optional<Op> &, size_t, const Vector &, const Matrix &,
                 const Vector &, size_t) { ++calls; return i >= 7; });
    auto res = LOBPCG<Vector, Matrix>(A, none, none, X0, 5, 10 * n, it, nc, 1e-8, stop);
    print_case("hook_stop", res, it, nc, calls);
  }
  {   // random-start overload (l.376-390) and the argument checks (l.148-155)
    size_t it = 0, nc = 0;
    auto res = LOBPCG<Vector, Matrix>(A, none, none, size_t(n), size_t(nx), size_t(3), size_t(10 * n), it, nc, 1e-8);
    print_case("random_start", res, it, nc, 0);
    int thrown = 0;
    try { LOBPCG<Vector, Matrix>(A, none, none, X0, nx + 1, 10, it, nc); } catch (const std::invalid_argument &) { ++thrown; }
    Matrix wide(2, 5);
    try { LOBPCG<Vector, Matrix>(A, none, none, wide, 1, 10, it, nc); } catch (const std::invalid_argument &) { ++thrown; }
    printf("{\"case\": \"invalid_argument\", \"thrown\": %d}\n", thrown);
  }
  return 0;
}
