// Host-side check of the drop-in header layer (CPU, no GPU): OUR Optimization/Riemannian/GradientDescent.h on the
// reference's S^2 test problem (tests/GradientDescent_unit_test.cpp shape) -- compared bit for bit with the golden
// run of the unmodified reference header -- and the Euclidean conveniences (EuclideanGradientDescent, EuclideanTNT
// with the reference's signatures) on a small quadratic.
#include <array>
#include <cstdio>
#include "Optimization/Riemannian/GradientDescent.h"
#include "Optimization/Riemannian/TNT.h"

struct V3 {
  std::array<double, 3> d{0, 0, 0};
  V3 &operator+=(const V3 &o) { for (int i = 0; i < 3; ++i) d[i] += o.d[i]; return *this; }
  V3 &operator-=(const V3 &o) { for (int i = 0; i < 3; ++i) d[i] -= o.d[i]; return *this; }
  V3 &operator*=(int a) { for (int i = 0; i < 3; ++i) d[i] *= a; return *this; }
  double dot(const V3 &o) const { return d[0] * o.d[0] + d[1] * o.d[1] + d[2] * o.d[2]; }
};
static V3 operator*(double a, const V3 &v) { V3 o; for (int i = 0; i < 3; ++i) o.d[i] = a * v.d[i]; return o; }
static V3 operator*(int a, const V3 &v) { return double(a) * v; }
static V3 operator-(const V3 &v) { V3 o; for (int i = 0; i < 3; ++i) o.d[i] = -v.d[i]; return o; }
static V3 operator+(const V3 &x, const V3 &y) { V3 o; for (int i = 0; i < 3; ++i) o.d[i] = x.d[i] + y.d[i]; return o; }

using namespace Optimization;

int main() {
  // ---- S^2: f(X; P) = |X - P|^2, projection retraction ----
  {
    auto project = [](const V3 &X, const V3 &W) { V3 o = W; const double c = X.dot(W); for (int i = 0; i < 3; ++i) o.d[i] -= c * X.d[i]; return o; };
    Objective<V3, double, V3> F = [](const V3 &X, V3 &P) { double s = 0; for (int i = 0; i < 3; ++i) s += (X.d[i] - P.d[i]) * (X.d[i] - P.d[i]); return s; };
    Riemannian::VectorField<V3, V3, V3> gradF = [project](const V3 &X, V3 &P) { V3 n; for (int i = 0; i < 3; ++i) n.d[i] = 2 * (X.d[i] - P.d[i]); return project(X, n); };
    Riemannian::RiemannianMetric<V3, V3, double, V3> metric = [](const V3 &, const V3 &a, const V3 &b, V3 &) { return a.dot(b); };
    Riemannian::Retraction<V3, V3, V3> retract = [](const V3 &X, const V3 &T, V3 &) { V3 z = X + T; const double n = std::sqrt(z.dot(z)); for (int i = 0; i < 3; ++i) z.d[i] /= n; return z; };
    Riemannian::GradientDescentParams<double> prm;
    prm.max_iterations = 1000;
    prm.gradient_tolerance = 1e-6;
    V3 X0; X0.d = {-.5, -.5, -.707107};
    V3 P; P.d = {0, 0, 1};
    size_t calls = 0;
    Riemannian::GradientDescentUserFunction<V3, V3, double, V3> hook =
        [&calls](size_t, double, const V3 &, double, const V3 &, const V3 &, double, V3 &) { ++calls; };
    auto res = Riemannian::GradientDescent<V3, V3, double, V3>(F, gradF, metric, retract, X0, P, prm, hook);
    printf("{\"case\": \"s2_gd\", \"status_code\": %d, \"iterations\": %zu, \"f\": %.17g, \"gradfx_norm\": %.17g, "
           "\"x\": [%.17g, %.17g, %.17g], \"hook_calls\": %zu, \"accepted\": %zu}\n",
           int(res.status), res.gradient_norms.size(), res.f, res.gradfx_norm, res.x.d[0], res.x.d[1], res.x.d[2], calls,
           res.linesearch_iterations.size());
    int thrown = 0;
    for (int c = 0; c < 4; ++c) {
      Riemannian::GradientDescentParams<double> bad;
      if (c == 0) bad.alpha = 0; else if (c == 1) bad.beta = 1; else if (c == 2) bad.sigma = 0; else bad.gradient_tolerance = -1;
      try { Riemannian::GradientDescent<V3, V3, double, V3>(F, gradF, metric, retract, X0, P, bad); } catch (const std::invalid_argument &) { ++thrown; }
    }
    printf("{\"case\": \"gd_invalid_argument\", \"thrown\": %d}\n", thrown);
  }
  // ---- Euclidean conveniences on f(x) = 1/2 x^T D x - b^T x, D = diag(1, 2, 4): minimiser b ./ D ----
  {
    const std::array<double, 3> D = {1, 2, 4}, b = {1, -2, 3};
    Objective<V3, double> f = [D, b](const V3 &x) { double s = 0; for (int i = 0; i < 3; ++i) s += .5 * D[i] * x.d[i] * x.d[i] - b[i] * x.d[i]; return s; };
    Riemannian::EuclideanVectorField<V3> grad = [D, b](const V3 &x) { V3 g; for (int i = 0; i < 3; ++i) g.d[i] = D[i] * x.d[i] - b[i]; return g; };
    Riemannian::EuclideanLinearOperatorConstructor<V3> HC = [D](const V3 &) {
      Riemannian::EuclideanLinearOperator<V3> H = [D](const V3 &, const V3 &v) { V3 o; for (int i = 0; i < 3; ++i) o.d[i] = D[i] * v.d[i]; return o; };
      return H;
    };
    V3 x0;
    Riemannian::GradientDescentParams<double> gp;
    gp.max_iterations = 500;
    gp.gradient_tolerance = 1e-9;
    gp.relative_decrease_tolerance = 0;
    gp.stepsize_tolerance = 0;
    auto gd = Riemannian::EuclideanGradientDescent<V3, double>(f, grad, x0, gp);   // Args... sits mid-list: explicit template arguments, as with the reference
    Riemannian::TNTParams<double> tp;
    tp.gradient_tolerance = 1e-9;
    const std::optional<Riemannian::EuclideanLinearOperator<V3>> no_precon;
    auto tn = Riemannian::EuclideanTNT<V3, double>(f, grad, HC, x0, no_precon, tp);
    double egd = 0, etn = 0;
    for (int i = 0; i < 3; ++i) { egd = std::fmax(egd, std::fabs(gd.x.d[i] - b[i] / D[i])); etn = std::fmax(etn, std::fabs(tn.x.d[i] - b[i] / D[i])); }
    printf("{\"case\": \"euclidean\", \"gd_status\": %d, \"gd_err\": %.3e, \"tnt_status\": %d, \"tnt_err\": %.3e, \"tnt_outer\": %zu}\n",
           int(gd.status), egd, int(tn.status), etn, tn.inner_iterations.size());
  }
  return 0;
}
