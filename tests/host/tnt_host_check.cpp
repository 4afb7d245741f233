// Host-side check of the drop-in header layer (CPU, no GPU): the reference's own S^2 TNT test problem
// (tests/TNT_unit_test.cpp:63-187) and the closed-form STPCG known-answer tests
// (tests/IterativeSolvers_unit_test.cpp:138-251) through OUR Optimization/Riemannian/TNT.h and
// Optimization/LinearAlgebra/IterativeSolvers.h with a tiny host vector type (control flow only).
// Prints one JSON line per case; tests/test_headers.py compares with tests/golden/golden.json.
#include <array>
#include <cstdio>
#include "Optimization/Riemannian/TNT.h"

struct V3 {
  std::array<double, 3> d{0, 0, 0};
  V3 &operator+=(const V3 &o) { for (int i = 0; i < 3; ++i) d[i] += o.d[i]; return *this; }
  V3 &operator-=(const V3 &o) { for (int i = 0; i < 3; ++i) d[i] -= o.d[i]; return *this; }
  V3 &operator*=(int a) { for (int i = 0; i < 3; ++i) d[i] *= a; return *this; }
};
static V3 operator*(double a, const V3 &v) { V3 o; for (int i = 0; i < 3; ++i) o.d[i] = a * v.d[i]; return o; }
static V3 operator*(int a, const V3 &v) { return double(a) * v; }
static V3 operator-(const V3 &v) { V3 o; for (int i = 0; i < 3; ++i) o.d[i] = -v.d[i]; return o; }
static V3 operator+(const V3 &x, const V3 &y) { V3 o; for (int i = 0; i < 3; ++i) o.d[i] = x.d[i] + y.d[i]; return o; }
static double dot(const V3 &a, const V3 &b) { return a.d[0] * b.d[0] + a.d[1] * b.d[1] + a.d[2] * b.d[2]; }

using namespace Optimization;

static void print_vec(const char *k, const std::vector<double> &v) {
  printf("\"%s\": [", k);
  for (size_t i = 0; i < v.size(); ++i) printf("%s%.17g", i ? ", " : "", v[i]);
  printf("]");
}

static void s2_tnt(bool use_precon, bool tight) {
  auto project = [](const V3 &X, const V3 &W) { V3 o = W; const double c = dot(X, W); for (int i = 0; i < 3; ++i) o.d[i] -= c * X.d[i]; return o; };
  Objective<V3, double, V3> F = [](const V3 &X, V3 &P) { double s = 0; for (int i = 0; i < 3; ++i) s += (X.d[i] - P.d[i]) * (X.d[i] - P.d[i]); return s; };
  Riemannian::VectorField<V3, V3, V3> gradF = [project](const V3 &X, V3 &P) { V3 n; for (int i = 0; i < 3; ++i) n.d[i] = 2 * (X.d[i] - P.d[i]); return project(X, n); };
  Riemannian::LinearOperatorConstructor<V3, V3, V3> HC = [project, gradF](const V3 &, V3 &) {
    Riemannian::LinearOperator<V3, V3, V3> H = [project, gradF](const V3 &X, const V3 &Xdot, V3 &P) {
      V3 EH = 2.0 * Xdot;
      V3 out = project(X, EH);
      const double c = dot(X, gradF(X, P));     // the reference's quirk: Riemannian gradient here
      for (int i = 0; i < 3; ++i) out.d[i] -= c * Xdot.d[i];
      return out;
    };
    return H;
  };
  Riemannian::RiemannianMetric<V3, V3, double, V3> metric = [](const V3 &, const V3 &a, const V3 &b, V3 &) { return dot(a, b); };
  Riemannian::Retraction<V3, V3, V3> retract = [](const V3 &X, const V3 &T, V3 &) { V3 z = X + T; const double n = std::sqrt(dot(z, z)); for (int i = 0; i < 3; ++i) z.d[i] /= n; return z; };
  std::optional<Riemannian::LinearOperator<V3, V3, V3>> precon;
  if (use_precon) precon = [](const V3 &, const V3 &T, V3 &) { V3 o; o.d = {1.0 * T.d[0], 2.0 * T.d[1], 3.0 * T.d[2]}; return o; };
  Riemannian::TNTParams<double> prm;
  if (tight) { prm.gradient_tolerance = 1e-8; prm.preconditioned_gradient_tolerance = 1e-8; prm.relative_decrease_tolerance = 1e-12; prm.stepsize_tolerance = 1e-12; }
  V3 X0; X0.d = {-.5, -.5, -.707107};
  V3 P; P.d = {0, 0, 1};
  auto res = Riemannian::TNT<V3, V3, double, V3>(F, gradF, HC, metric, retract, X0, P, precon, prm);
  printf("{\"case\": \"s2_tnt%s%s\", \"status_code\": %d, \"f\": %.17g, \"gradfx_norm\": %.17g, ", use_precon ? "_precon" : "", tight ? "_tight" : "", int(res.status), res.f, res.gradfx_norm);
  printf("\"x\": [%.17g, %.17g, %.17g], \"inner_iterations\": [", res.x.d[0], res.x.d[1], res.x.d[2]);
  for (size_t i = 0; i < res.inner_iterations.size(); ++i) printf("%s%zu", i ? ", " : "", res.inner_iterations[i]);
  printf("], ");
  print_vec("gain_ratios", res.gain_ratios); printf(", ");
  print_vec("trust_region_radius", res.trust_region_radius); printf(", ");
  print_vec("objective_values", res.objective_values); printf(", ");
  print_vec("gradient_norms", res.gradient_norms); printf(", ");
  print_vec("update_step_M_norms", res.update_step_M_norms);
  printf("}\n");
}

static void kat(const char *name, const V3 &g, const V3 &h, const double *M, double Delta) {
  using namespace LinearAlgebra;
  SymmetricLinearOperator<V3> H = [h](const V3 &x) { V3 o; for (int i = 0; i < 3; ++i) o.d[i] = h.d[i] * x.d[i]; return o; };
  InnerProduct<V3> ip = [](const V3 &a, const V3 &b) { return dot(a, b); };
  std::optional<STPCGPreconditioner<V3, std::nullptr_t>> P;
  if (M) { std::array<double, 3> mi = {1 / M[0], 1 / M[1], 1 / M[2]}; P = [mi](const V3 &x) { V3 o; for (int i = 0; i < 3; ++i) o.d[i] = mi[i] * x.d[i]; return std::make_pair(o, nullptr); }; }
  double mn = 0; size_t it = 0;
  V3 s = STPCG<V3, std::nullptr_t>(g, H, ip, mn, it, Delta, 3, 1e-8, .999, P);
  printf("{\"case\": \"%s\", \"s\": [%.17g, %.17g, %.17g], \"update_step_M_norm\": %.17g, \"num_iterations\": %zu}\n", name, s.d[0], s.d[1], s.d[2], mn, it);
}

int main() {
  s2_tnt(false, false); s2_tnt(true, false); s2_tnt(false, true); s2_tnt(true, true);
  V3 g; g.d = {21, -.4, 19};
  V3 h; h.d = {1000, 100, 1};
  V3 hn = -h;
  const double M[3] = {100, 10, 1};
  const double big = 1.7976931348623157e308;
  kat("ExactSTPCG", g, h, nullptr, big);
  kat("ExactSTPCGwithNegativeCurvature", g, hn, nullptr, 1000);
  kat("ExactSTPCGwithPreconditioning", g, h, M, big);
  kat("ExactSTPCGwithNegativeCurvatureAndPreconditioning", g, hn, M, 1000);
  // invalid arguments throw std::invalid_argument (IterativeSolvers.h:183-205, TNT.h:260-318)
  int thrown = 0;
  try { kat("bad", g, h, nullptr, -1.0); } catch (const std::invalid_argument &) { ++thrown; }
  printf("{\"case\": \"invalid_argument\", \"thrown\": %d}\n", thrown);
  return 0;
}
