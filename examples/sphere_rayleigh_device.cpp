// Example: the shape of the reference's examples/Riemannian_optimization_example.cpp (gradient descent, then the
// truncated-Newton trust-region method, on a sphere) with device-resident vectors: minimise the Rayleigh quotient
// f(x) = x^T A x over S^(n-1), A = diag(d) + U diag(sigma) U^T, through the drop-in headers.  The calls are the
// reference's (Riemannian::GradientDescent, Riemannian::TNT); only the vector type (b200::DeviceMatrix) and the
// functor set (b200::SphereRayleigh) differ, and the inner Steihaug-Toint solves run in one fused CUDA kernel each.
//
//   g++ -std=c++17 -O2 -Iinclude examples/sphere_rayleigh_device.cpp -Loptimization_b200 -loptimization_b200 \
//       -Wl,-rpath,$PWD/optimization_b200 -o build/sphere_rayleigh_device && build/sphere_rayleigh_device [n]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "Optimization/b200/Device.h"

using namespace Optimization;
using b200::DeviceMatrix;

int main(int argc, char **argv) {
  const size_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 1000000, k = 16;
  std::mt19937_64 gen(1);
  std::uniform_real_distribution<double> uni(0.0, 1.0);
  std::normal_distribution<double> normal(0.0, 1.0);
  std::vector<double> d(n), U(n * k), sigma(k), x0(n);
  for (auto &v : d) v = 1.0 + uni(gen);
  for (auto &v : U) v = normal(gen) / std::sqrt(double(n));
  for (auto &v : sigma) v = 2.0 * uni(gen) - 1.0;
  double nrm = 0;
  for (auto &v : x0) { v = normal(gen); nrm += v * v; }
  for (auto &v : x0) v /= std::sqrt(nrm);

  b200::Context ctx(0);
  b200::SphereRayleigh model(ctx.get(), n, k, d.data(), U.data(), sigma.data());
  DeviceMatrix X0(ctx.get(), n, 1, x0.data());

  // --- Riemannian gradient descent (reference GradientDescent.h) ---
  Riemannian::GradientDescentParams<double> gd_params;
  gd_params.max_iterations = 50;
  gd_params.verbose = true;
  auto gd = Riemannian::GradientDescent<DeviceMatrix, DeviceMatrix, double>(model.objective(), model.gradient(), model.metric(),
                                                                           model.retraction(), X0, gd_params);

  // --- truncated-Newton trust-region method (reference TNT.h), started from the gradient-descent iterate ---
  Riemannian::TNTParams<double> tnt_params;
  tnt_params.gradient_tolerance = 1e-8;
  tnt_params.verbose = true;
  const std::optional<Riemannian::LinearOperator<DeviceMatrix, DeviceMatrix>> no_precon;
  auto tnt = Riemannian::TNT<DeviceMatrix, DeviceMatrix, double>(model.objective(), model.quadratic_model(), model.metric(),
                                                                model.retraction(), gd.x, no_precon, tnt_params);
  std::printf("gradient descent: f = %.12g after %zu iterations; TNT: f = %.12g, |grad| = %.3e, %zu outer iterations, "
              "%llu kernels launched in total\n",
              gd.f, gd.objective_values.size(), tnt.f, tnt.gradfx_norm, tnt.inner_iterations.size(),
              (unsigned long long)ob200_kernel_launches(ctx.get()));
  return 0;
}
