// Host-side check (CPU, no GPU) of OUR Optimization/Riemannian/TNLS.h (EuclideanTNLS -> TNLS -> LSQR) on the
// curve-fitting problem of the reference's tests/TNLS_unit_test.cpp:  F(beta)_i = y_i - sin(beta_0 t_i + beta_1).
// Input: binary file written by tests/test_headers.py: [u64 ncases] then per case
//   [u64 m][f64 t(m)][f64 y(m)][u64 use_precon][u64 max_it][f64 root_tol][f64 grad_tol][f64 rel_tol][f64 step_tol][f64 Delta_tol]
// Output: one JSON line per case, compared with the golden runs of the unmodified reference header.
#include <cstdio>
#include <vector>
#include "Optimization/Riemannian/TNLS.h"

struct Vec {
  std::vector<double> d;
  Vec() = default;
  explicit Vec(size_t n) : d(n, 0.0) {}
  Vec &operator+=(const Vec &o) { for (size_t i = 0; i < d.size(); ++i) d[i] += o.d[i]; return *this; }
  Vec &operator-=(const Vec &o) { for (size_t i = 0; i < d.size(); ++i) d[i] -= o.d[i]; return *this; }
  Vec &operator*=(int a) { for (auto &x : d) x *= double(a); return *this; }
  Vec &operator/=(double a) { for (auto &x : d) x /= a; return *this; }
  double dot(const Vec &y) const {   // eight interleaved partial sums, pairwise tree (the oracle type's order)
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t n = d.size();
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
      for (int j = 0; j < 8; ++j) acc[j] += d[i + j] * y.d[i + j];
    for (int j = 0; i < n; ++i, ++j) acc[j] += d[i] * y.d[i];
    return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  }
};
static Vec operator*(double a, const Vec &v) { Vec o(v.d.size()); for (size_t i = 0; i < v.d.size(); ++i) o.d[i] = a * v.d[i]; return o; }
static Vec operator*(int a, const Vec &v) { return double(a) * v; }
static Vec operator/(const Vec &v, double a) { Vec o(v.d.size()); for (size_t i = 0; i < v.d.size(); ++i) o.d[i] = v.d[i] / a; return o; }
static Vec operator-(const Vec &v) { Vec o(v.d.size()); for (size_t i = 0; i < v.d.size(); ++i) o.d[i] = -v.d[i]; return o; }
static Vec operator+(const Vec &x, const Vec &y) { Vec o(x.d.size()); for (size_t i = 0; i < x.d.size(); ++i) o.d[i] = x.d[i] + y.d[i]; return o; }
static Vec operator-(const Vec &x, const Vec &y) { Vec o(x.d.size()); for (size_t i = 0; i < x.d.size(); ++i) o.d[i] = x.d[i] - y.d[i]; return o; }

using namespace Optimization;

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  unsigned long long ncases = 0;
  if (fread(&ncases, 8, 1, f) != 1) return 2;
  for (unsigned long long c = 0; c < ncases; ++c) {
    unsigned long long m = 0, use_precon = 0, max_it = 0;
    double tol[5];
    if (fread(&m, 8, 1, f) != 1) return 2;
    std::vector<double> t(m), y(m);
    if (fread(t.data(), 8, m, f) != m || fread(y.data(), 8, m, f) != m || fread(&use_precon, 8, 1, f) != 1 ||
        fread(&max_it, 8, 1, f) != 1 || fread(tol, 8, 5, f) != 5)
      return 2;
    std::vector<double> Jm(2 * m, 0.0), scale(2, 1.0);
    Riemannian::Mapping<Vec, Vec> F = [&](const Vec &b) { Vec o(m); for (size_t i = 0; i < m; ++i) o.d[i] = y[i] - std::sin(b.d[0] * t[i] + b.d[1]); return o; };
    Riemannian::JacobianPairFunction<Vec, Vec, Vec> JF = [&](const Vec &b) {
      for (size_t i = 0; i < m; ++i) { Jm[m + i] = -std::cos(b.d[0] * t[i] + b.d[1]); Jm[i] = Jm[m + i] * t[i]; }
      for (int k = 0; k < 2; ++k) { double s2 = 0; for (size_t i = 0; i < m; ++i) s2 += Jm[k * m + i] * Jm[k * m + i]; scale[k] = 1.0 / std::sqrt(s2); }
      Riemannian::Jacobian<Vec, Vec, Vec> DF = [&](const Vec &, const Vec &v) { Vec o(m); for (size_t i = 0; i < m; ++i) o.d[i] = Jm[i] * v.d[0] + Jm[m + i] * v.d[1]; return o; };
      Riemannian::JacobianAdjoint<Vec, Vec, Vec> DFt = [&](const Vec &, const Vec &w) { Vec o(2); for (int k = 0; k < 2; ++k) { double acc = 0; for (size_t i = 0; i < m; ++i) acc += Jm[k * m + i] * w.d[i]; o.d[k] = acc; } return o; };
      return std::make_pair(DF, DFt);
    };
    Riemannian::LinearOperator<Vec, Vec> M = [&](const Vec &, const Vec &v) { Vec o(2); o.d[0] = scale[0] * v.d[0]; o.d[1] = scale[1] * v.d[1]; return o; };
    std::optional<Riemannian::TNLSPreconditioner<Vec, Vec>> precon;
    if (use_precon) precon = std::make_pair(M, M);
    Riemannian::TNLSParams<double> prm;
    prm.max_iterations = size_t(max_it);
    prm.root_tolerance = tol[0]; prm.gradient_tolerance = tol[1]; prm.relative_decrease_tolerance = tol[2];
    prm.stepsize_tolerance = tol[3]; prm.Delta_tolerance = tol[4];
    Vec b0(2);
    b0.d = {1.0, 1.0};
    auto res = Riemannian::EuclideanTNLS<Vec>(F, JF, b0, precon, prm);
    printf("{\"case\": \"%llu\", \"status_code\": %d, \"f\": %.17g, \"gradfx_norm\": %.17g, \"x\": [%.17g, %.17g], \"inner_iterations\": [",
           c, int(res.status), res.f, res.gradfx_norm, res.x.d[0], res.x.d[1]);
    for (size_t i = 0; i < res.inner_iterations.size(); ++i) printf("%s%zu", i ? ", " : "", res.inner_iterations[i]);
    printf("], \"rho\": [");
    for (size_t i = 0; i < res.rho.size(); ++i) printf("%s%.17g", i ? ", " : "", res.rho[i]);
    printf("], \"trust_region_radius\": [");
    for (size_t i = 0; i < res.trust_region_radius.size(); ++i) printf("%s%.17g", i ? ", " : "", res.trust_region_radius[i]);
    printf("], \"objective_values\": [");
    for (size_t i = 0; i < res.objective_values.size(); ++i) printf("%s%.17g", i ? ", " : "", res.objective_values[i]);
    printf("]}\n");
  }
  fclose(f);
  return 0;
}
