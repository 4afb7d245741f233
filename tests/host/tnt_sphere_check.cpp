// GPU check of the drop-in header layer, sphere model: Rayleigh-quotient TNT end to end through
// Optimization::Riemannian::TNT<DeviceMatrix, DeviceMatrix, double> (fused device tCG inside).
// Input: binary problem file written by tests/test_headers.py
//   [u64 n][u64 k][f64 d(n)][f64 U(n*k)][f64 sigma(k)][f64 x0(n)]
// Output: one JSON line compared with tests/golden (generated from the unmodified reference headers).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "Optimization/b200/Device.h"

using namespace Optimization;
using b200::DeviceMatrix;

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  unsigned long long n = 0, k = 0;
  if (fread(&n, 8, 1, f) != 1 || fread(&k, 8, 1, f) != 1) return 2;
  std::vector<double> d(n), U(n * k), sigma(k), x0(n);
  if (fread(d.data(), 8, n, f) != n || fread(U.data(), 8, n * k, f) != n * k || fread(sigma.data(), 8, k, f) != k ||
      fread(x0.data(), 8, n, f) != n)
    return 2;
  fclose(f);

  b200::Context ctx(0);
  b200::SphereRayleigh prob(ctx.get(), n, k, d.data(), U.data(), sigma.data());
  DeviceMatrix X(ctx.get(), n, 1, x0.data());
  Riemannian::TNTParams<double> prm;   // defaults, like the golden run
  const std::optional<Riemannian::LinearOperator<DeviceMatrix, DeviceMatrix>> no_precon;
  const unsigned long long l0 = ob200_kernel_launches(ctx.get());
  auto res = Riemannian::TNT<DeviceMatrix, DeviceMatrix, double>(prob.objective(), prob.quadratic_model(), prob.metric(),
                                                                prob.retraction(), X, no_precon, prm);
  printf("{\"case\": \"sphere_tnt\", \"status_code\": %d, \"f\": %.17g, \"gradfx_norm\": %.17g, \"launches\": %llu, \"inner_iterations\": [",
         int(res.status), res.f, res.gradfx_norm, ob200_kernel_launches(ctx.get()) - l0);
  for (size_t i = 0; i < res.inner_iterations.size(); ++i) printf("%s%zu", i ? ", " : "", res.inner_iterations[i]);
  printf("], \"gain_ratios\": [");
  for (size_t i = 0; i < res.gain_ratios.size(); ++i) printf("%s%.17g", i ? ", " : "", res.gain_ratios[i]);
  printf("], \"trust_region_radius\": [");
  for (size_t i = 0; i < res.trust_region_radius.size(); ++i) printf("%s%.17g", i ? ", " : "", res.trust_region_radius[i]);
  printf("], \"objective_values\": [");
  for (size_t i = 0; i < res.objective_values.size(); ++i) printf("%s%.17g", i ? ", " : "", res.objective_values[i]);
  printf("]}\n");
  const std::vector<double> x = res.x.to_host();
  FILE *o = fopen(argv[2], "wb");
  fwrite(x.data(), 8, x.size(), o);
  fclose(o);

  // ---- GradientDescent<DeviceMatrix, DeviceMatrix, double> on the same model (level-1 kernels + model / retraction
  //      entry points; reference GradientDescent.h:124-398), same parameters as the golden run ----
  if (argc > 3) {
    Riemannian::GradientDescentParams<double> gp;
    gp.max_iterations = 60;
    gp.gradient_tolerance = 1e-6;
    auto gd = Riemannian::GradientDescent<DeviceMatrix, DeviceMatrix, double>(prob.objective(), prob.gradient(),
                                                                             prob.metric(), prob.retraction(), X, gp);
    size_t ls = 0;
    for (size_t v : gd.linesearch_iterations) ls += v;
    printf("{\"case\": \"sphere_gd\", \"status_code\": %d, \"iterations\": %zu, \"linesearch_total\": %zu, \"f\": %.17g, "
           "\"gradfx_norm\": %.17g}\n", int(gd.status), gd.gradient_norms.size(), ls, gd.f, gd.gradfx_norm);
    const std::vector<double> xg = gd.x.to_host();
    FILE *og = fopen(argv[3], "wb");
    fwrite(xg.data(), 8, xg.size(), og);
    fclose(og);
  }
  return 0;
}
