// Host-side check (CPU, no GPU) of the projected / constraint-preconditioned path of OUR
// Optimization/LinearAlgebra/IterativeSolvers.h (P and At supplied, Multiplier != nullptr_t; reference
// IterativeSolvers.h:236-252, 388-404) and of the STPCGUserFunction hook (l.365-369), on the problem shape of the
// reference's tests/IterativeSolvers_unit_test.cpp:316-496.  Input: binary file [u64 n][u64 mc][h n][m n][A mc*n][g n]
// written by tests/test_headers.py; output JSON lines compared with the golden run of the unmodified reference header.
#include <cstdio>
#include <vector>
#include "Optimization/LinearAlgebra/IterativeSolvers.h"
#include "../../oracle/dense_lu.hpp"   // test helper shared with the oracle: identical KKT arithmetic on both sides

struct Vec {
  std::vector<double> d;
  Vec() = default;
  explicit Vec(size_t n) : d(n, 0.0) {}
  Vec &operator+=(const Vec &o) { for (size_t i = 0; i < d.size(); ++i) d[i] += o.d[i]; return *this; }
  Vec &operator-=(const Vec &o) { for (size_t i = 0; i < d.size(); ++i) d[i] -= o.d[i]; return *this; }
  Vec &operator*=(int a) { for (auto &x : d) x *= double(a); return *this; }
};
static Vec operator*(double a, const Vec &v) { Vec o(v.d.size()); for (size_t i = 0; i < v.d.size(); ++i) o.d[i] = a * v.d[i]; return o; }
static Vec operator*(int a, const Vec &v) { return double(a) * v; }
static Vec operator-(const Vec &v) { Vec o(v.d.size()); for (size_t i = 0; i < v.d.size(); ++i) o.d[i] = -v.d[i]; return o; }
static Vec operator+(const Vec &x, const Vec &y) { Vec o(x.d.size()); for (size_t i = 0; i < x.d.size(); ++i) o.d[i] = x.d[i] + y.d[i]; return o; }
// eight interleaved partial sums, pairwise tree (the summation order of the oracle's stand-in type)
static double dot(const Vec &x, const Vec &y) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const size_t n = x.d.size();
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; ++j) acc[j] += x.d[i + j] * y.d[i + j];
  for (int j = 0; i < n; ++i, ++j) acc[j] += x.d[i] * y.d[i];
  return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

using namespace Optimization::LinearAlgebra;

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  unsigned long long n = 0, mc = 0;
  if (fread(&n, 8, 1, f) != 1 || fread(&mc, 8, 1, f) != 1) return 2;
  std::vector<double> h(n), m(n), A(mc * n), g(n);
  if (fread(h.data(), 8, n, f) != n || fread(m.data(), 8, n, f) != n || fread(A.data(), 8, mc * n, f) != mc * n ||
      fread(g.data(), 8, n, f) != n)
    return 2;
  fclose(f);

  oracle::DenseLU lu;
  if (!lu.factor(oracle::kkt_matrix(m.data(), A.data(), n, mc), n + mc)) return 3;
  SymmetricLinearOperator<Vec> H = [&](const Vec &v) { Vec o(n); for (size_t i = 0; i < n; ++i) o.d[i] = h[i] * v.d[i]; return o; };
  InnerProduct<Vec> ip = [](const Vec &a, const Vec &b) { return dot(a, b); };
  STPCGPreconditioner<Vec, Vec> P = [&](const Vec &r) {
    std::vector<double> w(n + mc, 0.0);
    for (size_t i = 0; i < n; ++i) w[i] = r.d[i];
    const std::vector<double> z = lu.solve(w);
    Vec x(n), l(mc);
    for (size_t i = 0; i < n; ++i) x.d[i] = z[i];
    for (size_t c = 0; c < mc; ++c) l.d[c] = z[n + c];
    return std::make_pair(x, l);
  };
  LinearOperator<Vec, Vec> At = [&](const Vec &l) {
    Vec o(n);
    for (size_t i = 0; i < n; ++i) { double acc = 0; for (size_t c = 0; c < mc; ++c) acc += A[c * n + i] * l.d[c]; o.d[i] = acc; }
    return o;
  };
  const std::optional<STPCGPreconditioner<Vec, Vec>> Pop(P);
  const std::optional<LinearOperator<Vec, Vec>> Atop(At);
  Vec G(n);
  G.d = g;
  const double big = 1.7976931348623157e308;
  struct Case { const char *name; double Delta, kappa; } cases[] = {{"projected_exact", big, 1e-8}, {"projected_trunc", 1e-4, .1}};
  for (const Case &c : cases) {
    double mn = 0;
    size_t it = 0;
    Vec s = STPCG<Vec, Vec>(G, H, ip, mn, it, c.Delta, 250, c.kappa, .7, Pop, Atop);
    double As = 0;   // constraint residual ||A s||
    for (size_t r = 0; r < mc; ++r) { double acc = 0; for (size_t i = 0; i < n; ++i) acc += A[r * n + i] * s.d[i]; As += acc * acc; }
    double sMs = 0;
    for (size_t i = 0; i < n; ++i) sMs += m[i] * s.d[i] * s.d[i];
    printf("{\"case\": \"%s\", \"num_iterations\": %zu, \"update_step_M_norm\": %.17g, \"As_norm\": %.3e, \"s_M_norm\": %.17g, \"s\": [",
           c.name, it, mn, std::sqrt(As), std::sqrt(sMs));
    for (size_t i = 0; i < n; ++i) printf("%s%.17g", i ? ", " : "", s.d[i]);
    printf("]}\n");
  }
  // user hook: called once per iteration before the update, may stop the loop (reference l.365-369)
  {
    size_t calls = 0;
    STPCGUserFunction<Vec, Vec> hook = [&calls](size_t k, const Vec &, const SymmetricLinearOperator<Vec> &,
                                                const std::optional<STPCGPreconditioner<Vec, Vec>> &,
                                                const std::optional<LinearOperator<Vec, Vec>> &, const Vec &, const Vec &,
                                                const Vec &, const Vec &, double) { ++calls; return k == 2; };
    double mn = 0;
    size_t it = 0;
    const std::optional<STPCGUserFunction<Vec, Vec>> hop(hook);
    STPCG<Vec, Vec>(G, H, ip, mn, it, big, 250, 1e-8, .7, Pop, Atop, hop);
    printf("{\"case\": \"user_hook\", \"calls\": %zu, \"num_iterations\": %zu}\n", calls, it);
  }
  return 0;
}
