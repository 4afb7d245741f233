// GPU check of the drop-in header layer: Optimization::LinearAlgebra::LOBPCG<std::vector<double>, DeviceMatrix>
// with block-operator descriptor functors (reference entry point LOBPCG.h:131-140) on the diagonal problem of the
// reference's tests/LOBPCG_unit_test.cpp (n = 1000, block 10, nev 5, tau 1e-8, generalized + preconditioned) and on a
// small 3-D Laplacian.  Input: binary file [u64 n][u64 nx][f64 adiag(n)][f64 bdiag(n)][f64 X0(n*nx)].
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "Optimization/b200/Device.h"

using namespace Optimization;
using b200::BlockOperator;
using b200::DeviceMatrix;
using Op = LinearAlgebra::SymmetricLinearOperator<DeviceMatrix>;

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  unsigned long long n = 0, nx = 0;
  if (fread(&n, 8, 1, f) != 1 || fread(&nx, 8, 1, f) != 1) return 2;
  std::vector<double> ad(n), bd(n), x0(n * nx), tabs(n);
  if (fread(ad.data(), 8, n, f) != n || fread(bd.data(), 8, n, f) != n || fread(x0.data(), 8, n * nx, f) != n * nx) return 2;
  fclose(f);
  for (size_t i = 0; i < n; ++i) tabs[i] = ad[i] < 0 ? -ad[i] : ad[i];

  b200::Context ctx(0);
  DeviceMatrix X0(ctx.get(), n, nx, x0.data());
  const Op A = BlockOperator::diagonal(ctx.get(), n, ad.data());
  const std::optional<Op> B(BlockOperator::diagonal(ctx.get(), n, bd.data()));
  const std::optional<Op> T(BlockOperator::diagonal(ctx.get(), n, tabs.data()));
  const std::optional<Op> none;
  for (int generalized = 0; generalized < 2; ++generalized) {
    size_t it = 0, nc = 0;
    auto res = LinearAlgebra::LOBPCG<std::vector<double>, DeviceMatrix>(A, generalized ? B : none, T, X0, 5, n, it, nc, 1e-8);
    printf("{\"case\": \"diag%s\", \"num_iters\": %zu, \"nc\": %zu, \"theta\": [", generalized ? "_generalized" : "", it, nc);
    for (size_t i = 0; i < res.first.size(); ++i) printf("%s%.17g", i ? ", " : "", res.first[i]);
    printf("], \"x_cols\": %zu}\n", res.second.cols());
  }
  // the functors are ordinary callables too: A(X) applies the operator
  {
    DeviceMatrix AX = A(X0);
    const std::vector<double> h = AX.to_host();
    double err = 0;
    for (size_t i = 0; i < n; ++i)
      for (size_t c = 0; c < nx; ++c) err = std::fmax(err, std::fabs(h[i * nx + c] - ad[i] * x0[i * nx + c]));
    printf("{\"case\": \"apply\", \"max_err\": %.3e}\n", err);
  }
  // opaque lambdas are not descriptor functors: the device-only layer must say so
  {
    int thrown = 0;
    const Op L = [](const DeviceMatrix &X) { return X; };
    size_t it = 0, nc = 0;
    try { LinearAlgebra::LOBPCG<std::vector<double>, DeviceMatrix>(L, none, none, X0, 5, 10, it, nc); } catch (const std::invalid_argument &) { ++thrown; }
    try { LinearAlgebra::LOBPCG<std::vector<double>, DeviceMatrix>(A, none, none, X0, nx + 1, 10, it, nc); } catch (const std::invalid_argument &) { ++thrown; }   // nev > nx (LOBPCG.h:148)
    printf("{\"case\": \"invalid_argument\", \"thrown\": %d}\n", thrown);
  }
  return 0;
}
