// GPU check of the least-squares entry points on device vectors: LinearAlgebra::LSQR<DeviceMatrix> (reference
// IterativeSolvers.h:552-855) and Riemannian::EuclideanTNLS<DeviceMatrix> (reference TNLS.h:265-729, over LSQR) with
// operators made of device level-1 kernels only (ob200_hadamard / ob200_axpby / ob200_div / ob200_dot).
// Input: [u64 n][f64 d(n)][f64 b(n)][f64 c(n)][f64 x0(n)][u64 lsqr_max_it][f64 lsqr_prm(4): lambda btol Atol cond_limit]
//        [u64 tnls_max_it][f64 tnls_tol(5): root grad rel step Delta]
// Output: JSON lines compared with tests/golden (unmodified reference headers on the same operators).
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "Optimization/b200/Device.h"
#include "Optimization/Riemannian/TNLS.h"

using namespace Optimization;
using b200::DeviceMatrix;

static DeviceMatrix had(const DeviceMatrix &a, const DeviceMatrix &b) {
  DeviceMatrix out = b.like();
  b200::check(b.context(), ob200_hadamard(b.context(), b.size(), a.data(), b.data(), out.data()));
  return out;
}
static void dump(const char *path, const DeviceMatrix &x) {
  const std::vector<double> h = x.to_host();
  FILE *o = fopen(path, "wb");
  fwrite(h.data(), 8, h.size(), o);
  fclose(o);
}

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  unsigned long long n = 0, lsqr_it = 0, tnls_it = 0;
  double lp[4], tp[5];
  if (fread(&n, 8, 1, f) != 1) return 2;
  std::vector<double> d(n), b(n), c(n), x0(n);
  if (fread(d.data(), 8, n, f) != n || fread(b.data(), 8, n, f) != n || fread(c.data(), 8, n, f) != n ||
      fread(x0.data(), 8, n, f) != n || fread(&lsqr_it, 8, 1, f) != 1 || fread(lp, 8, 4, f) != 4 ||
      fread(&tnls_it, 8, 1, f) != 1 || fread(tp, 8, 5, f) != 5)
    return 2;
  fclose(f);
  b200::Context ctx(0);
  const DeviceMatrix D(ctx.get(), n, 1, d.data()), B(ctx.get(), n, 1, b.data()), Cv(ctx.get(), n, 1, c.data());
  const std::vector<double> one(n, 1.0);
  const DeviceMatrix Ones(ctx.get(), n, 1, one.data());
  {   // ---- LSQR: min |diag(d) x - b| ----
    LinearAlgebra::LinearOperator<DeviceMatrix, DeviceMatrix> A = [&](const DeviceMatrix &x) { return had(D, x); };
    LinearAlgebra::InnerProduct<DeviceMatrix> ip = [](const DeviceMatrix &u, const DeviceMatrix &v) { return b200::dot(u, v); };
    double xnorm = -1;
    size_t it = 0;
    const unsigned long long l0 = ob200_kernel_launches(ctx.get());
    DeviceMatrix x = LinearAlgebra::LSQR<DeviceMatrix>(A, A, B, ip, xnorm, it, size_t(lsqr_it), lp[0], lp[1], lp[2], lp[3], DBL_MAX);
    printf("{\"case\": \"lsqr_diag\", \"num_iterations\": %zu, \"xnorm\": %.17g, \"launches\": %llu}\n", it, xnorm,
           ob200_kernel_launches(ctx.get()) - l0);
    dump(argv[2], x);
  }
  {   // ---- TNLS: F(x) = (d o x o x + x) - c,  DF = DF^T = diag(2 (d o x) + 1) ----
    Riemannian::Mapping<DeviceMatrix, DeviceMatrix> F = [&](const DeviceMatrix &x) { return (had(D, had(x, x)) + x) - Cv; };
    Riemannian::JacobianPairFunction<DeviceMatrix, DeviceMatrix, DeviceMatrix> JF = [&](const DeviceMatrix &x) {
      auto jd = std::make_shared<const DeviceMatrix>(2.0 * had(D, x) + Ones);
      Riemannian::Jacobian<DeviceMatrix, DeviceMatrix, DeviceMatrix> DF = [jd](const DeviceMatrix &, const DeviceMatrix &v) { return had(*jd, v); };
      Riemannian::JacobianAdjoint<DeviceMatrix, DeviceMatrix, DeviceMatrix> DFt = [jd](const DeviceMatrix &, const DeviceMatrix &w) { return had(*jd, w); };
      return std::make_pair(DF, DFt);
    };
    Riemannian::TNLSParams<double> prm;
    prm.max_iterations = size_t(tnls_it);
    prm.root_tolerance = tp[0]; prm.gradient_tolerance = tp[1]; prm.relative_decrease_tolerance = tp[2];
    prm.stepsize_tolerance = tp[3]; prm.Delta_tolerance = tp[4];
    const DeviceMatrix X0(ctx.get(), n, 1, x0.data());
    const std::optional<Riemannian::TNLSPreconditioner<DeviceMatrix, DeviceMatrix>> no_precon;
    auto res = Riemannian::EuclideanTNLS<DeviceMatrix>(F, JF, X0, no_precon, prm);
    printf("{\"case\": \"tnls_elem\", \"status_code\": %d, \"f\": %.17g, \"gradfx_norm\": %.17g, \"inner_iterations\": [",
           int(res.status), res.f, res.gradfx_norm);
    for (size_t i = 0; i < res.inner_iterations.size(); ++i) printf("%s%zu", i ? ", " : "", res.inner_iterations[i]);
    printf("], \"rho\": [");
    for (size_t i = 0; i < res.rho.size(); ++i) printf("%s%.17g", i ? ", " : "", res.rho[i]);
    printf("], \"trust_region_radius\": [");
    for (size_t i = 0; i < res.trust_region_radius.size(); ++i) printf("%s%.17g", i ? ", " : "", res.trust_region_radius[i]);
    printf("], \"objective_values\": [");
    for (size_t i = 0; i < res.objective_values.size(); ++i) printf("%s%.17g", i ? ", " : "", res.objective_values[i]);
    printf("]}\n");
    dump(argv[3], res.x);
  }
  return 0;
}
