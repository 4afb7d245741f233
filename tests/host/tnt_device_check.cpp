// GPU check of the drop-in header layer: Stiefel trace-min TNT end to end through
// Optimization::Riemannian::TNT<DeviceMatrix, DeviceMatrix, double> (fused device tCG inside), plus a
// direct STPCG call on descriptor functors.  Input: a binary problem file written by tests/test_headers.py
//   [u64 n][u64 p][u16 A(nblk*128*128)][f64 Y0(n*p)][f64 g(n*p)]
// Output: JSON lines compared with tests/golden (generated from the unmodified reference headers).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "Optimization/b200/Device.h"

using namespace Optimization;
using b200::DeviceMatrix;

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  unsigned long long n = 0, p = 0;
  if (fread(&n, 8, 1, f) != 1 || fread(&p, 8, 1, f) != 1) return 2;
  const size_t nblk = (n + 127) / 128;
  std::vector<uint16_t> A(nblk * 128 * 128);
  std::vector<double> Y0(n * p), g(n * p);
  if (fread(A.data(), 2, A.size(), f) != A.size() || fread(Y0.data(), 8, Y0.size(), f) != Y0.size() ||
      fread(g.data(), 8, g.size(), f) != g.size())
    return 2;
  fclose(f);

  b200::Context ctx(0);
  b200::StiefelTraceMin prob(ctx.get(), n, p, A.data());
  DeviceMatrix Y(ctx.get(), n, p, Y0.data());

  // ---- TNT end to end (default TNTParams, like the golden run) ----
  Riemannian::TNTParams<double> prm;
  const std::optional<Riemannian::LinearOperator<DeviceMatrix, DeviceMatrix>> no_precon;
  auto res = Riemannian::TNT<DeviceMatrix, DeviceMatrix, double>(prob.objective(), prob.quadratic_model(),
                                                                prob.metric(), prob.retraction(), Y, no_precon, prm);
  printf("{\"case\": \"tnt\", \"status_code\": %d, \"f\": %.17g, \"gradfx_norm\": %.17g, \"last_path\": %d, \"inner_iterations\": [",
         int(res.status), res.f, res.gradfx_norm, ob200_last_path(ctx.get()));
  for (size_t i = 0; i < res.inner_iterations.size(); ++i) printf("%s%zu", i ? ", " : "", res.inner_iterations[i]);
  printf("], \"gain_ratios\": [");
  for (size_t i = 0; i < res.gain_ratios.size(); ++i) printf("%s%.17g", i ? ", " : "", res.gain_ratios[i]);
  printf("], \"trust_region_radius\": [");
  for (size_t i = 0; i < res.trust_region_radius.size(); ++i) printf("%s%.17g", i ? ", " : "", res.trust_region_radius[i]);
  printf("], \"objective_values\": [");
  for (size_t i = 0; i < res.objective_values.size(); ++i) printf("%s%.17g", i ? ", " : "", res.objective_values[i]);
  printf("]}\n");
  {
    const std::vector<double> x = res.x.to_host();
    FILE *o = fopen(argc > 2 ? argv[2] : "/tmp/tnt_x.bin", "wb");
    fwrite(x.data(), 8, x.size(), o);
    fclose(o);
  }

  // ---- direct STPCG on descriptor functors: must take the fused path ----
  {
    DeviceMatrix grad, G(ctx.get(), n, p, g.data());
    Riemannian::LinearOperator<DeviceMatrix, DeviceMatrix> Hess;
    prob.quadratic_model()(Y, grad, Hess);
    const b200::FusedHessian *fh = Hess.target<b200::FusedHessian>();
    LinearAlgebra::SymmetricLinearOperator<DeviceMatrix> H = b200::BoundHessian{fh->st};
    LinearAlgebra::InnerProduct<DeviceMatrix> ip = b200::FrobeniusProduct{};
    double mn = 0;
    size_t it = 0;
    const unsigned long long l0 = ob200_kernel_launches(ctx.get());
    DeviceMatrix s = LinearAlgebra::STPCG<DeviceMatrix, std::nullptr_t>(G, H, ip, mn, it, 1e6, 200, 1e-9, 0.0);
    const unsigned long long fused_launches = ob200_kernel_launches(ctx.get()) - l0;
    // same call through opaque lambdas: generic loop over level-1 kernels, same answer
    LinearAlgebra::SymmetricLinearOperator<DeviceMatrix> Hl = [&](const DeviceMatrix &v) { return b200::BoundHessian{fh->st}(v); };
    LinearAlgebra::InnerProduct<DeviceMatrix> ipl = [](const DeviceMatrix &a, const DeviceMatrix &b) { return b200::dot(a, b); };
    double mn2 = 0;
    size_t it2 = 0;
    const unsigned long long l1 = ob200_kernel_launches(ctx.get());
    DeviceMatrix s2 = LinearAlgebra::STPCG<DeviceMatrix, std::nullptr_t>(G, Hl, ipl, mn2, it2, 1e6, 200, 1e-9, 0.0);
    const unsigned long long generic_launches = ob200_kernel_launches(ctx.get()) - l1;
    DeviceMatrix d = s - s2;
    printf("{\"case\": \"stpcg\", \"num_iterations\": %zu, \"update_step_M_norm\": %.17g, \"generic_iterations\": %zu, "
           "\"generic_M_norm\": %.17g, \"rel_diff\": %.3e, \"fused_launches\": %llu, \"generic_launches\": %llu}\n",
           it, mn, it2, mn2, std::sqrt(b200::dot(d, d) / b200::dot(s, s)), fused_launches, generic_launches);
    const std::vector<double> x = s.to_host();
    FILE *o = fopen(argc > 3 ? argv[3] : "/tmp/stpcg_s.bin", "wb");
    fwrite(x.data(), 8, x.size(), o);
    fclose(o);
  }
  return 0;
}
