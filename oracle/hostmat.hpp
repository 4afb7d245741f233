// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// HostMat: a minimal dense host vector/matrix type that satisfies exactly the
// requirements the reference's Krylov loops place on their `Vector` template
// argument (default-constructible, copy/move assignable, `int * v`, unary `-`,
// `Scalar * v`, `v + w`, `+=`, `-=`, `*= int`; see
// /root/reference/include/Optimization/LinearAlgebra/IterativeSolvers.h:211,256,
// 324,336,374,377,420).  Eigen is not installed in this image, so this type is
// the stand-in the "Eigen CPU reference" is instantiated with.  Like Eigen it
// evaluates `x + a*y`, `-v + b*p`, `r += a*Hp` in ONE fused pass with no
// temporaries (lightweight expression objects below), and its dot product uses
// eight independent accumulators (Eigen-style packet accumulation) so gcc can
// vectorise it without -ffast-math.  All loops are split statically over
// `g_threads` OpenMP threads with per-thread partials combined in thread
// order, so results are deterministic for a fixed thread count.
#pragma once
#include <cstddef>
#include <cstring>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

inline int &threads() {
  static int t = 1;
  return t;
}

struct HostMat;
struct Scaled { double a; const HostMat *v; };                 // a * v
struct Sum { const HostMat *x; double a; const HostMat *y; };  // x + a * y
struct Sum2 { double a; const HostMat *x; double b; const HostMat *y; }; // a*x+b*y

template <class F> inline void parallel_ranges(size_t n, F &&f) {
  int T = threads();
  if (T <= 1 || n < 65536) {
    f(0, size_t(0), n);
    return;
  }
#pragma omp parallel for schedule(static) num_threads(T)
  for (int t = 0; t < T; ++t) {
    size_t lo = n * size_t(t) / size_t(T), hi = n * size_t(t + 1) / size_t(T);
    f(t, lo, hi);
  }
}

struct HostMat {
  std::vector<double> d;
  HostMat() {}
  explicit HostMat(size_t n) : d(n, 0.0) {}
  HostMat(const double *src, size_t n) : d(src, src + n) {}
  HostMat(const Scaled &s);
  HostMat(const Sum &s);
  double dot(const HostMat &o) const;   // Euclidean helpers of the reference call V1.dot(V2)
  size_t size() const { return d.size(); }
  double *data() { return d.data(); }
  const double *data() const { return d.data(); }

  HostMat &operator=(const Scaled &s);
  HostMat &operator=(const Sum &s);
  HostMat &operator=(const Sum2 &s);
  HostMat &operator+=(const Scaled &s);
  HostMat &operator+=(const HostMat &o);
  HostMat &operator-=(const HostMat &o);
  HostMat &operator*=(int a);
  HostMat &operator/=(double a) {
    for (auto &v : d) v /= a;
    return *this;
  }
};

inline Scaled operator*(double a, const HostMat &v) { return Scaled{a, &v}; }
inline Scaled operator*(int a, const HostMat &v) { return Scaled{double(a), &v}; }
inline Scaled operator-(const HostMat &v) { return Scaled{-1.0, &v}; }
inline Sum operator+(const HostMat &x, const Scaled &s) { return Sum{&x, s.a, s.v}; }
inline Sum operator-(const HostMat &x, const Scaled &s) { return Sum{&x, -s.a, s.v}; }   // x - a y == x + (-a) y, bit for bit
inline Sum operator+(const HostMat &x, const HostMat &y) { return Sum{&x, 1.0, &y}; }    // 1.0 * y is exact
inline Sum2 operator+(const Scaled &s1, const Scaled &s2) {
  return Sum2{s1.a, s1.v, s2.a, s2.v};
}

inline HostMat::HostMat(const Scaled &s) : d(s.v->size()) { *this = s; }
inline HostMat::HostMat(const Sum &s) : d(s.x->size()) { *this = s; }
inline HostMat operator/(const HostMat &v, double a) {
  HostMat o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o.d[i] = v.d[i] / a;
  return o;
}

inline HostMat &HostMat::operator=(const Scaled &s) {
  if (d.size() != s.v->size()) d.resize(s.v->size());
  double *o = d.data();
  const double *v = s.v->data();
  const double a = s.a;
  if (a == -1.0)
    parallel_ranges(d.size(), [&](int, size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i) o[i] = -v[i];
    });
  else
    parallel_ranges(d.size(), [&](int, size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i) o[i] = a * v[i];
    });
  return *this;
}
inline HostMat &HostMat::operator=(const Sum &s) {
  if (d.size() != s.x->size()) d.resize(s.x->size());
  double *o = d.data();
  const double *x = s.x->data(), *y = s.y->data();
  const double a = s.a;
  parallel_ranges(d.size(), [&](int, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) o[i] = x[i] + a * y[i];
  });
  return *this;
}
inline HostMat &HostMat::operator=(const Sum2 &s) {
  if (d.size() != s.x->size()) d.resize(s.x->size());
  double *o = d.data();
  const double *x = s.x->data(), *y = s.y->data();
  const double a = s.a, b = s.b;
  if (a == -1.0)
    parallel_ranges(d.size(), [&](int, size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i) o[i] = -x[i] + b * y[i];
    });
  else
    parallel_ranges(d.size(), [&](int, size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i) o[i] = a * x[i] + b * y[i];
    });
  return *this;
}
inline HostMat &HostMat::operator+=(const Scaled &s) {
  double *o = d.data();
  const double *v = s.v->data();
  const double a = s.a;
  parallel_ranges(d.size(), [&](int, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) o[i] += a * v[i];
  });
  return *this;
}
inline HostMat &HostMat::operator+=(const HostMat &v) {
  double *o = d.data();
  const double *x = v.data();
  parallel_ranges(d.size(), [&](int, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) o[i] += x[i];
  });
  return *this;
}
inline HostMat &HostMat::operator-=(const HostMat &v) {
  double *o = d.data();
  const double *x = v.data();
  parallel_ranges(d.size(), [&](int, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) o[i] -= x[i];
  });
  return *this;
}
inline HostMat &HostMat::operator*=(int a) {
  double *o = d.data();
  const double aa = a;
  parallel_ranges(d.size(), [&](int, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) o[i] *= aa;
  });
  return *this;
}

// <x, y> with eight independent accumulators, combined as a fixed tree.
inline double dot_range(const double *x, const double *y, size_t lo, size_t hi) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t i = lo;
  for (; i + 8 <= hi; i += 8)
    for (int j = 0; j < 8; ++j) acc[j] += x[i + j] * y[i + j];
  for (int j = 0; i < hi; ++i, ++j) acc[j] += x[i] * y[i];
  return ((acc[0] + acc[1]) + (acc[2] + acc[3])) +
         ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

inline double dot(const double *x, const double *y, size_t n) {
  int T = threads();
  if (T <= 1 || n < 65536) return dot_range(x, y, 0, n);
  std::vector<double> part(size_t(T), 0.0);
  parallel_ranges(n, [&](int t, size_t lo, size_t hi) {
    part[size_t(t)] = dot_range(x, y, lo, hi);
  });
  double s = 0;
  for (int t = 0; t < T; ++t) s += part[size_t(t)];
  return s;
}
inline double dot(const HostMat &x, const HostMat &y) {
  return dot(x.data(), y.data(), x.size());
}
inline double HostMat::dot(const HostMat &o) const { return oracle::dot(*this, o); }

} // namespace oracle
