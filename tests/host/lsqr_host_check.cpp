// Host-side check (CPU, no GPU) of OUR LSQR (include/Optimization/LinearAlgebra/IterativeSolvers.h) on dense systems.
// Input: binary file written by tests/test_headers.py: [u64 ncases] then per case
//   [u64 m][u64 n][f64 A(m*n)][f64 b(m)][u64 max_it][f64 lambda][f64 btol][f64 Atol][f64 cond_limit][f64 Delta]
// Output: one JSON line per case, compared bit for bit with the golden runs of the unmodified reference header.
#include <cstdio>
#include <vector>
#include "Optimization/LinearAlgebra/IterativeSolvers.h"

struct Vec {
  std::vector<double> d;
  Vec() = default;
  explicit Vec(size_t n) : d(n, 0.0) {}
  Vec &operator+=(const Vec &o) { for (size_t i = 0; i < d.size(); ++i) d[i] += o.d[i]; return *this; }
  Vec &operator-=(const Vec &o) { for (size_t i = 0; i < d.size(); ++i) d[i] -= o.d[i]; return *this; }
  Vec &operator*=(int a) { for (auto &x : d) x *= double(a); return *this; }
  Vec &operator/=(double a) { for (auto &x : d) x /= a; return *this; }
};
static Vec operator*(double a, const Vec &v) { Vec o(v.d.size()); for (size_t i = 0; i < v.d.size(); ++i) o.d[i] = a * v.d[i]; return o; }
static Vec operator*(int a, const Vec &v) { return double(a) * v; }
static Vec operator-(const Vec &v) { Vec o(v.d.size()); for (size_t i = 0; i < v.d.size(); ++i) o.d[i] = -v.d[i]; return o; }
static Vec operator+(const Vec &x, const Vec &y) { Vec o(x.d.size()); for (size_t i = 0; i < x.d.size(); ++i) o.d[i] = x.d[i] + y.d[i]; return o; }
static Vec operator-(const Vec &x, const Vec &y) { Vec o(x.d.size()); for (size_t i = 0; i < x.d.size(); ++i) o.d[i] = x.d[i] - y.d[i]; return o; }
static double dot(const Vec &x, const Vec &y) {   // eight interleaved partial sums, pairwise tree (the oracle type's order)
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
  unsigned long long ncases = 0;
  if (fread(&ncases, 8, 1, f) != 1) return 2;
  for (unsigned long long c = 0; c < ncases; ++c) {
    unsigned long long m = 0, n = 0, max_it = 0;
    double prm[5];
    if (fread(&m, 8, 1, f) != 1 || fread(&n, 8, 1, f) != 1) return 2;
    std::vector<double> A(m * n), b(m);
    if (fread(A.data(), 8, m * n, f) != m * n || fread(b.data(), 8, m, f) != m || fread(&max_it, 8, 1, f) != 1 ||
        fread(prm, 8, 5, f) != 5)
      return 2;
    LinearOperator<Vec, Vec> Aop = [&](const Vec &x) { Vec o(m); for (size_t i = 0; i < m; ++i) { double acc = 0; for (size_t j = 0; j < n; ++j) acc += A[i * n + j] * x.d[j]; o.d[i] = acc; } return o; };
    LinearOperator<Vec, Vec> Atop = [&](const Vec &y) { Vec o(n); for (size_t j = 0; j < n; ++j) { double acc = 0; for (size_t i = 0; i < m; ++i) acc += A[i * n + j] * y.d[i]; o.d[j] = acc; } return o; };
    InnerProduct<Vec> ip = [](const Vec &a, const Vec &c2) { return dot(a, c2); };
    Vec B(m);
    B.d = b;
    double xnorm = -1;
    size_t it = 0;
    Vec x = LSQR<Vec>(Aop, Atop, B, ip, xnorm, it, size_t(max_it), prm[0], prm[1], prm[2], prm[3], prm[4]);
    x.d.resize(n, 0.0);
    printf("{\"case\": \"%llu\", \"num_iterations\": %zu, \"xnorm\": %.17g, \"x\": [", c, it, xnorm);
    for (size_t i = 0; i < n; ++i) printf("%s%.17g", i ? ", " : "", x.d[i]);
    printf("]}\n");
  }
  fclose(f);
  int thrown = 0;   // argument checks (reference IterativeSolvers.h:568-587)
  {
    LinearOperator<Vec, Vec> I = [](const Vec &x) { return x; };
    InnerProduct<Vec> ip = [](const Vec &a, const Vec &c2) { return dot(a, c2); };
    Vec B(2);
    double xn; size_t it;
    const double args5[5][5] = {{-1, 1e-6, 1e-6, 1e8, 1}, {0, -1, 1e-6, 1e8, 1}, {0, 1e-6, -1, 1e8, 1}, {0, 1e-6, 1e-6, 0, 1}, {0, 1e-6, 1e-6, 1e8, 0}};
    for (auto &a : args5) { try { LSQR<Vec>(I, I, B, ip, xn, it, 10, a[0], a[1], a[2], a[3], a[4]); } catch (const std::invalid_argument &) { ++thrown; } }
  }
  printf("{\"case\": \"lsqr_invalid_argument\", \"thrown\": %d}\n", thrown);
  return 0;
}
