// GPU check of TNT with a preconditioner through the drop-in header layer (reference TNT.h:247; adapter l.413-426):
//   (1) sphere model + pointwise Jacobi: b200::JacobiPreconditioner is recognised by TNT's inner views, the
//       preconditioned solves stay on the fused tCG path (few launches per outer iteration);
//   (2) Stiefel model + tangent-space preserving projected Jacobi P_Y(minv o V): generic STPCG loop over the device
//       level-1 kernels and ob200_stiefel_project.
// Input: binary file written by tests/test_headers.py
//   [u64 n][u64 k][f64 d(n)][f64 U(n*k)][f64 sigma(k)][f64 x0(n)][f64 minv(n)]
//   [u64 N][u64 p][u16 A(nblk*128*128)][f64 Y0(N*p)][f64 minv(N*p)]
// Output: JSON lines compared with tests/golden (generated from the unmodified reference headers).
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
#include "Optimization/b200/Device.h"

using namespace Optimization;
using b200::DeviceMatrix;

template <class Res>
static void report(const char *name, const Res &res, unsigned long long launches) {
  printf("{\"case\": \"%s\", \"status_code\": %d, \"f\": %.17g, \"gradfx_norm\": %.17g, \"launches\": %llu, \"inner_iterations\": [",
         name, int(res.status), res.f, res.gradfx_norm, launches);
  for (size_t i = 0; i < res.inner_iterations.size(); ++i) printf("%s%zu", i ? ", " : "", res.inner_iterations[i]);
  printf("], \"gain_ratios\": [");
  for (size_t i = 0; i < res.gain_ratios.size(); ++i) printf("%s%.17g", i ? ", " : "", res.gain_ratios[i]);
  printf("], \"trust_region_radius\": [");
  for (size_t i = 0; i < res.trust_region_radius.size(); ++i) printf("%s%.17g", i ? ", " : "", res.trust_region_radius[i]);
  printf("], \"objective_values\": [");
  for (size_t i = 0; i < res.objective_values.size(); ++i) printf("%s%.17g", i ? ", " : "", res.objective_values[i]);
  printf("], \"preconditioned_gradient_norms\": [");
  for (size_t i = 0; i < res.preconditioned_gradient_norms.size(); ++i)
    printf("%s%.17g", i ? ", " : "", res.preconditioned_gradient_norms[i]);
  printf("]}\n");
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
  b200::Context ctx(0);
  Riemannian::TNTParams<double> prm;   // defaults, like the golden runs
  {
    unsigned long long n = 0, k = 0;
    if (fread(&n, 8, 1, f) != 1 || fread(&k, 8, 1, f) != 1) return 2;
    std::vector<double> d(n), U(n * k), sigma(k), x0(n), minv(n);
    if (fread(d.data(), 8, n, f) != n || fread(U.data(), 8, n * k, f) != n * k || fread(sigma.data(), 8, k, f) != k ||
        fread(x0.data(), 8, n, f) != n || fread(minv.data(), 8, n, f) != n)
      return 2;
    b200::SphereRayleigh prob(ctx.get(), n, k, d.data(), U.data(), sigma.data());
    DeviceMatrix X(ctx.get(), n, 1, x0.data());
    const std::optional<Riemannian::LinearOperator<DeviceMatrix, DeviceMatrix>> precon =
        Riemannian::LinearOperator<DeviceMatrix, DeviceMatrix>(
            b200::JacobiPreconditioner{std::make_shared<const DeviceMatrix>(ctx.get(), n, 1, minv.data())});
    const unsigned long long l0 = ob200_kernel_launches(ctx.get());
    auto res = Riemannian::TNT<DeviceMatrix, DeviceMatrix, double>(prob.objective(), prob.quadratic_model(), prob.metric(),
                                                                  prob.retraction(), X, precon, prm);
    report("sphere_tnt_jacobi", res, ob200_kernel_launches(ctx.get()) - l0);
    dump(argv[2], res.x);
  }
  {
    unsigned long long n = 0, p = 0;
    if (fread(&n, 8, 1, f) != 1 || fread(&p, 8, 1, f) != 1) return 2;
    const size_t nblk = (n + 127) / 128;
    std::vector<uint16_t> A(nblk * 128 * 128);
    std::vector<double> Y0(n * p), minv(n * p);
    if (fread(A.data(), 2, A.size(), f) != A.size() || fread(Y0.data(), 8, Y0.size(), f) != Y0.size() ||
        fread(minv.data(), 8, minv.size(), f) != minv.size())
      return 2;
    b200::StiefelTraceMin prob(ctx.get(), n, p, A.data());
    DeviceMatrix Y(ctx.get(), n, p, Y0.data());
    const std::optional<Riemannian::LinearOperator<DeviceMatrix, DeviceMatrix>> precon =
        Riemannian::LinearOperator<DeviceMatrix, DeviceMatrix>(b200::ProjectedJacobiPreconditioner{
            ctx.get(), std::make_shared<const DeviceMatrix>(ctx.get(), n, p, minv.data())});
    const unsigned long long l0 = ob200_kernel_launches(ctx.get());
    auto res = Riemannian::TNT<DeviceMatrix, DeviceMatrix, double>(prob.objective(), prob.quadratic_model(), prob.metric(),
                                                                  prob.retraction(), Y, precon, prm);
    report("stiefel_tnt_pjacobi", res, ob200_kernel_launches(ctx.get()) - l0);
    dump(argv[3], res.x);
  }
  fclose(f);
  return 0;
}
