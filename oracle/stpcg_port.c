/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 *
 * Plain-C restatement of the reference's tCG hot path, used as the portable
 * CPU oracle (the other oracle, oracle/_ref, is the reference's own headers
 * compiled in the dev container).  Parity is PINNED: tests/test_oracle.py
 * checks this file against (a) the closed-form known-answer tests of
 * /root/reference/tests/IterativeSolvers_unit_test.cpp:138-251, (b) the golden
 * traces in tests/golden/ generated from oracle/_ref, and (c) oracle/_ref
 * itself, bit for bit, on seeded inputs (single thread).
 *
 * What is restated, with the reference lines each block follows:
 *   port_stpcg  <- Optimization::LinearAlgebra::STPCG,
 *                  include/Optimization/LinearAlgebra/IterativeSolvers.h:166-426
 *   port_tnt    <- Optimization::Riemannian::TNT,
 *                  include/Optimization/Riemannian/TNT.h:242-689
 * Vectors are flat double arrays; the operator / preconditioner / manifold are
 * C callbacks.  Arithmetic order matches oracle/hostmat.hpp (8-accumulator dot,
 * one fused pass per vector statement, no FMA contraction) so that the two
 * oracles agree exactly.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef void (*port_apply_fn)(void *ctx, const double *in, double *out);

static double dot8(const double *x, const double *y, size_t n) {
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int j = 0; j < 8; ++j) acc[j] += x[i + j] * y[i + j];
  for (int j = 0; i < n; ++i, ++j) acc[j] += x[i] * y[i];
  return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

double port_dot(const double *x, const double *y, uint64_t n) { return dot8(x, y, n); }

/* exit reasons (ours; the reference only distinguishes them by control flow) */
enum {
  PORT_EXIT_RESIDUAL = 0,   /* IterativeSolvers.h:290  target residual reached      */
  PORT_EXIT_MAXIT = 1,      /* IterativeSolvers.h:285  loop bound                   */
  PORT_EXIT_KERNEL = 2,     /* IterativeSolvers.h:305-337 p in ker(H), to boundary  */
  PORT_EXIT_BOUNDARY = 3,   /* IterativeSolvers.h:347-361 kappa<=0 or step too long */
  PORT_EXIT_BADARG = -1     /* IterativeSolvers.h:183-205 std::invalid_argument     */
};

/* Steihaug-Toint truncated preconditioned CG (unconstrained form: the At /
 * Multiplier path of IterativeSolvers.h:236-252,388-404 is never taken by TNT,
 * TNT.h:489-492). precon may be NULL (v = r, IterativeSolvers.h:229-231). */
int port_stpcg(uint64_t n, const double *g, port_apply_fn H, void *Hctx, port_apply_fn precon,
               void *Pctx, double Delta, uint64_t max_iterations, double kappa_fgr, double theta,
               double epsilon, double *s, double *update_step_M_norm, uint64_t *num_iterations) {
  /* argument checks, IterativeSolvers.h:183-205 */
  if (!(Delta > 0)) return PORT_EXIT_BADARG;
  if (kappa_fgr < 0 || kappa_fgr >= 1) return PORT_EXIT_BADARG;
  if (theta < 0 || theta > 1) return PORT_EXIT_BADARG;
  if (epsilon <= 0 || epsilon >= 1) return PORT_EXIT_BADARG;

  double *r = (double *)malloc(n * sizeof(double));
  double *v = (double *)malloc(n * sizeof(double));
  double *p = (double *)malloc(n * sizeof(double));
  double *Hp = (double *)malloc(n * sizeof(double));
  int reason = PORT_EXIT_MAXIT;

  for (size_t i = 0; i < n; ++i) s[i] = 0 * g[i];  /* l.211 */
  memcpy(r, g, n * sizeof(double));                /* l.214 */
  if (!precon) memcpy(v, r, n * sizeof(double));   /* l.231 */
  else precon(Pctx, r, v);                         /* l.234 */
  for (size_t i = 0; i < n; ++i) p[i] = -v[i];     /* l.256 */

  double sk_M_pk = 0, sk_M_2 = 0;                  /* l.259-263 */
  double pk_M_2 = dot8(r, v, n);                   /* l.266 */
  const double Delta_2 = Delta * Delta;            /* l.271 */
  const double r0_norm = sqrt(dot8(r, v, n));      /* l.275 */
  const double target = r0_norm * fmin(kappa_fgr, pow(r0_norm, theta)); /* l.278-279 */

  uint64_t k;
  for (k = 0; k < max_iterations; ++k) {           /* l.285 */
    if (sqrt(dot8(r, v, n)) <= target) {           /* l.290 */
      reason = PORT_EXIT_RESIDUAL;
      break;
    }
    H(Hctx, p, Hp);                                /* l.294 */
    const double kappa = dot8(p, Hp, n);           /* l.300 */
    if (sqrt(dot8(Hp, Hp, n)) / sqrt(dot8(p, p, n)) < epsilon) { /* l.305-307 */
      if (dot8(p, r, n) < 0) {                     /* l.320-326 */
        for (size_t i = 0; i < n; ++i) p[i] *= -1.0;
        sk_M_pk *= -1;
      }
      const double sigma =
          (-sk_M_pk + sqrt(sk_M_pk * sk_M_pk + pk_M_2 * (Delta_2 - sk_M_2))) / pk_M_2; /* l.330 */
      *update_step_M_norm = Delta;
      for (size_t i = 0; i < n; ++i) s[i] += sigma * p[i];
      *num_iterations = k;
      reason = PORT_EXIT_KERNEL;
      goto done;
    }
    const double alpha = dot8(r, v, n) / kappa;    /* l.341 */
    const double skp1_M_2 = sk_M_2 + 2 * alpha * sk_M_pk + alpha * alpha * pk_M_2; /* l.344 */
    if (kappa <= 0 || skp1_M_2 > Delta_2) {        /* l.347 */
      const double sigma =
          (-sk_M_pk + sqrt(sk_M_pk * sk_M_pk + pk_M_2 * (Delta_2 - sk_M_2))) / pk_M_2; /* l.355 */
      *update_step_M_norm = Delta;
      for (size_t i = 0; i < n; ++i) s[i] += sigma * p[i];
      *num_iterations = k;
      reason = PORT_EXIT_BOUNDARY;
      goto done;
    }
    for (size_t i = 0; i < n; ++i) s[i] = s[i] + alpha * p[i];   /* l.374 */
    for (size_t i = 0; i < n; ++i) r[i] += alpha * Hp[i];        /* l.377 */
    if (!precon) memcpy(v, r, n * sizeof(double));               /* l.383 */
    else precon(Pctx, r, v);                                     /* l.386 */
    const double rk_vk = dot8(r, v, n);                          /* l.408 */
    const double beta = rk_vk / (alpha * kappa);                 /* l.412 */
    sk_M_2 = skp1_M_2;                                           /* l.415 */
    sk_M_pk = beta * (sk_M_pk + alpha * pk_M_2);                 /* l.416 */
    pk_M_2 = rk_vk + beta * beta * pk_M_2;                       /* l.417 */
    for (size_t i = 0; i < n; ++i) p[i] = -v[i] + beta * p[i];   /* l.420 */
  }
  *num_iterations = k;
  *update_step_M_norm = sqrt(sk_M_2);                            /* l.424 */
done:
  free(r);
  free(v);
  free(p);
  free(Hp);
  return reason;
}

/* ---- built-in operators --------------------------------------------------- */
typedef struct { uint64_t n; const double *d; } port_diag;
static void diag_apply(void *c, const double *in, double *out) {
  const port_diag *D = (const port_diag *)c;
  for (size_t i = 0; i < D->n; ++i) out[i] = D->d[i] * in[i];
}

int port_stpcg_diag(uint64_t n, const double *g, const double *hdiag, const double *minv,
                    double Delta, uint64_t max_iterations, double kappa_fgr, double theta,
                    double epsilon, double *s, double *mnorm, uint64_t *iters) {
  port_diag Hd = {n, hdiag}, Pd = {n, minv};
  return port_stpcg(n, g, diag_apply, &Hd, minv ? diag_apply : NULL, &Pd, Delta, max_iterations,
                    kappa_fgr, theta, epsilon, s, mnorm, iters);
}

/* Rayleigh-quotient Hessian on the sphere (config C2; the model of the reference's
 * examples/Riemannian_optimization_example.cpp with a structured A):
 *   f(x) = x^T A x,  A = diag(d) + U diag(sigma) U^T  (U: n x k row-major),
 *   Hess f(x)[v] = 2 (A v - (x^T A v) x) - 2 (x^T A x) v.
 * Same statement order as oracle/ref_driver.cpp (SphereOp::apply, single-thread path, and
 * ref_sphere_stpcg), so both oracles round identically. */
typedef struct {
  uint64_t n, k;
  const double *d, *U, *sigma, *x;
  double xAx;
  double *Av; /* scratch, n */
} port_sphere;
static void sphere_A(const port_sphere *P, const double *v, double *out) {
  double t[64];
  for (size_t j = 0; j < P->k; ++j) t[j] = 0.0;
  for (size_t r = 0; r < P->n; ++r) {            /* t = U^T v, row order, fma */
    const double vr = v[r];
    const double *u = P->U + r * P->k;
    for (size_t j = 0; j < P->k; ++j) t[j] = fma(u[j], vr, t[j]);
  }
  for (size_t j = 0; j < P->k; ++j) t[j] = t[j] * P->sigma[j];
  for (size_t r = 0; r < P->n; ++r) {            /* out = d .* v + U t */
    double acc = P->d[r] * v[r];
    const double *u = P->U + r * P->k;
    for (size_t j = 0; j < P->k; ++j) acc = fma(u[j], t[j], acc);
    out[r] = acc;
  }
}
static void sphere_hess(void *c, const double *v, double *out) {
  port_sphere *P = (port_sphere *)c;
  sphere_A(P, v, P->Av);
  const double xAv = dot8(P->x, P->Av, P->n);
  for (size_t i = 0; i < P->n; ++i)
    out[i] = 2.0 * (P->Av[i] - xAv * P->x[i]) - 2.0 * P->xAx * v[i];
}
/* f = x^T A x, optional Ax (n), optional grad = 2 (A x - f x) */
double port_sphere_model(uint64_t n, uint64_t k, const double *d, const double *U, const double *sigma,
                         const double *x, double *Ax_out, double *grad_out) {
  if (k > 64) return NAN;
  double *Ax = (double *)malloc(n * sizeof(double));
  port_sphere P = {n, k, d, U, sigma, x, 0.0, NULL};
  sphere_A(&P, x, Ax);
  const double f = dot8(x, Ax, n);
  if (grad_out)
    for (size_t i = 0; i < n; ++i) grad_out[i] = 2.0 * (Ax[i] - f * x[i]);
  if (Ax_out) memcpy(Ax_out, Ax, n * sizeof(double));
  free(Ax);
  return f;
}
void port_sphere_hess(uint64_t n, uint64_t k, const double *d, const double *U, const double *sigma,
                      const double *x, const double *v, double *out) {
  port_sphere P = {n, k, d, U, sigma, x, 0.0, (double *)malloc(n * sizeof(double))};
  P.xAx = port_sphere_model(n, k, d, U, sigma, x, NULL, NULL);
  sphere_hess(&P, v, out);
  free(P.Av);
}
int port_stpcg_sphere(uint64_t n, uint64_t k, const double *d, const double *U, const double *sigma,
                      const double *x, const double *g, const double *minv, double Delta,
                      uint64_t max_iterations, double kappa_fgr, double theta, double epsilon, double *s,
                      double *mnorm, uint64_t *iters) {
  if (k > 64) return PORT_EXIT_BADARG;
  port_sphere P = {n, k, d, U, sigma, x, 0.0, (double *)malloc(n * sizeof(double))};
  P.xAx = port_sphere_model(n, k, d, U, sigma, x, NULL, NULL);
  port_diag Pd = {n, minv};
  const int rc = port_stpcg(n, g, sphere_hess, &P, minv ? diag_apply : NULL, &Pd, Delta, max_iterations,
                            kappa_fgr, theta, epsilon, s, mnorm, iters);
  free(P.Av);
  return rc;
}

/* Stiefel trace-min Hessian  Hess f(Y)[V] = P_Y(A V - V S),  S = sym(Y^T A Y),
 * A block-diagonal (double copy of the bf16 blocks).  Same loop order as
 * oracle/ref_driver.cpp so both oracles round identically. */
typedef struct {
  uint64_t n, p, nb, nblk;
  const double *A;   /* nblk*nb*nb */
  const double *Y;   /* n*p */
  const double *S;   /* p*p */
  double *W;         /* n*p scratch */
} port_stiefel;

static void blockdiag_apply(const port_stiefel *P, const double *V, double *out) {
  const size_t p = P->p, nb = P->nb;
  double acc[64];
  for (size_t b = 0; b < P->nblk; ++b) {
    const double *Ab = P->A + b * nb * nb;
    const size_t r0 = b * nb;
    const size_t rows = (P->n - r0 < nb) ? P->n - r0 : nb;
    for (size_t i = 0; i < rows; ++i) {
      for (size_t j = 0; j < p; ++j) acc[j] = 0.0;
      for (size_t k = 0; k < rows; ++k) {
        const double a = Ab[i * nb + k];
        const double *Vk = V + (r0 + k) * p;
        for (size_t j = 0; j < p; ++j) acc[j] = fma(a, Vk[j], acc[j]);
      }
      for (size_t j = 0; j < p; ++j) out[(r0 + i) * p + j] = acc[j];
    }
  }
}
static void gram_sym(const double *X, const double *Z, size_t n, size_t p, double *G) {
  memset(G, 0, p * p * sizeof(double));
  for (size_t r = 0; r < n; ++r)
    for (size_t i = 0; i < p; ++i) {
      const double xi = X[r * p + i];
      for (size_t j = 0; j < p; ++j) G[i * p + j] = fma(xi, Z[r * p + j], G[i * p + j]);
    }
  /* the driver sums per-thread partials starting from 0.0: with one thread
   * that is 0.0 + x == x, so no extra rounding */
  for (size_t i = 0; i < p; ++i)
    for (size_t j = i; j < p; ++j) {
      const double s = 0.5 * (G[i * p + j] + G[j * p + i]);
      G[i * p + j] = s;
      G[j * p + i] = s;
    }
}
static void sub_right_mul(const double *W, const double *X, const double *M, size_t n, size_t p,
                          double *out) {
  double acc[64];
  for (size_t r = 0; r < n; ++r) {
    for (size_t j = 0; j < p; ++j) acc[j] = 0.0;
    for (size_t k = 0; k < p; ++k) {
      const double xk = X[r * p + k];
      for (size_t j = 0; j < p; ++j) acc[j] = fma(xk, M[k * p + j], acc[j]);
    }
    for (size_t j = 0; j < p; ++j) out[r * p + j] = W[r * p + j] - acc[j];
  }
}
static void stiefel_apply(void *c, const double *V, double *out) {
  const port_stiefel *P = (const port_stiefel *)c;
  double G[64 * 64];
  blockdiag_apply(P, V, P->W);
  sub_right_mul(P->W, V, P->S, P->n, P->p, P->W);
  gram_sym(P->Y, P->W, P->n, P->p, G);
  sub_right_mul(P->W, P->Y, G, P->n, P->p, out);
}

/* S = sym(Y^T A Y) */
void port_stiefel_S(uint64_t n, uint64_t p, uint64_t nb, const double *A, const double *Y,
                    double *S) {
  port_stiefel P = {n, p, nb, (n + nb - 1) / nb, A, Y, NULL, NULL};
  double *AY = (double *)malloc(n * p * sizeof(double));
  blockdiag_apply(&P, Y, AY);
  gram_sym(Y, AY, n, p, S);
  free(AY);
}

int port_stpcg_stiefel(uint64_t n, uint64_t p, uint64_t nb, const double *A, const double *Y,
                       const double *g, const double *minv, double Delta,
                       uint64_t max_iterations, double kappa_fgr, double theta, double epsilon,
                       double *s, double *mnorm, uint64_t *iters) {
  if (p > 64) return PORT_EXIT_BADARG;
  double *S = (double *)malloc(p * p * sizeof(double));
  double *W = (double *)malloc(n * p * sizeof(double));
  port_stiefel_S(n, p, nb, A, Y, S);
  port_stiefel P = {n, p, nb, (n + nb - 1) / nb, A, Y, S, W};
  port_diag Pd = {n * p, minv};
  int rc = port_stpcg(n * p, g, stiefel_apply, &P, minv ? diag_apply : NULL, &Pd, Delta,
                      max_iterations, kappa_fgr, theta, epsilon, s, mnorm, iters);
  free(S);
  free(W);
  return rc;
}

/* ---- TNT outer loop (TNT.h:242-689) over C callbacks ----------------------- */
typedef struct {
  uint64_t nx;  /* scalars in a point  */
  uint64_t nt;  /* scalars in a tangent */
  double (*f)(void *ctx, const double *x);
  /* QM: writes grad and refreshes whatever state hess() reads (TNT.h:380,573) */
  void (*qm)(void *ctx, const double *x, double *grad);
  void (*hess)(void *ctx, const double *x, const double *v, double *out);
  double (*metric)(void *ctx, const double *x, const double *a, const double *b);
  void (*retract)(void *ctx, const double *x, const double *h, double *out);
  void (*precon)(void *ctx, const double *x, const double *v, double *out); /* nullable */
  void *ctx;
} port_manifold;

typedef struct {
  const port_manifold *M;
  const double *x;
} port_bound;
static void bound_hess(void *c, const double *v, double *out) {
  port_bound *B = (port_bound *)c;
  B->M->hess(B->M->ctx, B->x, v, out);      /* TNT.h:400-403 */
}
static void bound_precon(void *c, const double *v, double *out) {
  port_bound *B = (port_bound *)c;
  B->M->precon(B->M->ctx, B->x, v, out);    /* TNT.h:413-419 */
}

enum { ST_GRADIENT = 0, ST_PRECON_GRADIENT, ST_RELATIVE_DECREASE, ST_STEPSIZE, ST_TRUST_REGION,
       ST_ITERATION_LIMIT, ST_ELAPSED_TIME, ST_USER_FUNCTION };

/* prm layout = oracle/refapi.py:_prm.  Traces as TNTResult (TNT.h:168-194).
 * STPCG is called through `metric` exactly as TNT.h:406-410 binds it; because
 * every in-tree metric is the Frobenius product, port_stpcg's dot8 is used
 * directly (asserted by the parity tests against oracle/_ref). */
int port_tnt(const port_manifold *M, const double *x0, const double *prm, double *x_out,
             int *status, uint64_t *n_outer, uint64_t *n_trace, double *scalars, uint64_t cap,
             uint64_t *inner_iterations, double *radius, double *rho_tr, double *fvals,
             double *gradnorms, double *step_norms, double *step_M_norms) {
  const uint64_t max_iterations = (uint64_t)prm[0];
  const double gradient_tolerance = prm[1], relative_decrease_tolerance = prm[2],
               stepsize_tolerance = prm[3], precon_gradient_tolerance = prm[4],
               Delta_tolerance = prm[5], Delta0 = prm[6], eta1 = prm[7], eta2 = prm[8],
               alpha1 = prm[9], alpha2 = prm[10];
  const uint64_t max_TPCG = (uint64_t)prm[11];
  const double kappa_fgr = prm[12], theta = prm[13];
  /* argument checks TNT.h:260-318 */
  if (gradient_tolerance < 0 || precon_gradient_tolerance < 0 || relative_decrease_tolerance < 0 ||
      stepsize_tolerance < 0 || Delta_tolerance < 0 || Delta0 <= 0 || eta1 <= 0 || eta1 >= 1 ||
      eta1 > eta2 || eta2 >= 1 || alpha1 <= 0 || alpha1 >= 1 || alpha2 <= 1 || kappa_fgr <= 0 ||
      kappa_fgr >= 1 || theta < 0)
    return 1;
  const double sqrt_eps = sqrt(DBL_EPSILON);      /* l.323 */
  const uint64_t nx = M->nx, nt = M->nt;
  double *x = (double *)malloc(nx * sizeof(double)), *xp = (double *)malloc(nx * sizeof(double));
  double *grad = (double *)malloc(nt * sizeof(double)), *h = (double *)malloc(nt * sizeof(double));
  double *tmp = (double *)malloc(nt * sizeof(double));
  int st = ST_ITERATION_LIMIT;                     /* l.327 */
  uint64_t no = 0, ntr = 0;
  memcpy(x, x0, nx * sizeof(double));              /* l.375 */
  double fx = M->f(M->ctx, x);                     /* l.377 */
  M->qm(M->ctx, x, grad);                          /* l.380 */
  double gnorm = sqrt(M->metric(M->ctx, x, grad, grad)), pgnorm; /* l.382 */
  if (M->precon) {
    M->precon(M->ctx, x, grad, tmp);               /* l.385 */
    pgnorm = sqrt(M->metric(M->ctx, x, tmp, tmp));
  } else
    pgnorm = gnorm;                                /* l.391 */
  double Delta = Delta0;                           /* l.429 */
  port_bound B = {M, x};
  for (uint64_t it = 0; it < max_iterations; ++it) {   /* l.446 */
    if (ntr < cap) { radius[ntr] = Delta; fvals[ntr] = fx; gradnorms[ntr] = gnorm; }
    ++ntr;                                         /* l.455-459 */
    if (gnorm < gradient_tolerance) { st = ST_GRADIENT; break; }          /* l.474 */
    if (pgnorm < precon_gradient_tolerance) { st = ST_PRECON_GRADIENT; break; } /* l.478 */
    double hM = 0;
    uint64_t inner = 0;
    int rc = port_stpcg(nt, grad, bound_hess, &B, M->precon ? bound_precon : NULL, &B, Delta,
                        max_TPCG, kappa_fgr, theta, 1e-8, h, &hM, &inner);   /* l.489-492 */
    if (rc == PORT_EXIT_BADARG) { free(x); free(xp); free(grad); free(h); free(tmp); return 1; }
    const double hnorm = sqrt(M->metric(M->ctx, x, h, h));                  /* l.493 */
    M->retract(M->ctx, x, h, xp);                                          /* l.505 */
    const double fxp = M->f(M->ctx, xp);                                   /* l.508 */
    M->hess(M->ctx, x, h, tmp);
    const double dm = -M->metric(M->ctx, x, grad, h) - .5 * M->metric(M->ctx, x, h, tmp); /* l.511 */
    const double df = fx - fxp;                                            /* l.515 */
    const double rel = df / (sqrt_eps + fabs(fx));                         /* l.518 */
    const double rho = df / dm;                                            /* l.521 */
    const int accepted = (!isnan(rho) && rho > eta1);                      /* l.532 */
    if (no < cap) { inner_iterations[no] = inner; step_norms[no] = hnorm; step_M_norms[no] = hM; rho_tr[no] = rho; }
    ++no;                                                                  /* l.538-541 */
    if (accepted) {                                                        /* l.555 */
      double *t = x; x = xp; xp = t; B.x = x;
      fx = fxp;
      if (rel < relative_decrease_tolerance) { st = ST_RELATIVE_DECREASE; break; } /* l.561 */
      if (hnorm < stepsize_tolerance) { st = ST_STEPSIZE; break; }         /* l.567 */
      M->qm(M->ctx, x, grad);                                              /* l.573 */
      gnorm = sqrt(M->metric(M->ctx, x, grad, grad));
      if (M->precon) {
        M->precon(M->ctx, x, grad, tmp);
        pgnorm = sqrt(M->metric(M->ctx, x, tmp, tmp));
      } else
        pgnorm = gnorm;
    }
    if (!isnan(rho) && rho >= eta2)                                        /* l.590 */
      Delta = fmax(alpha2 * hM, Delta);
    else if (isnan(rho) || rho < eta1) {                                   /* l.594 */
      Delta = alpha1 * hM;
      if (Delta < Delta_tolerance) { st = ST_TRUST_REGION; break; }        /* l.599 */
    }
  }
  if (ntr < cap) { radius[ntr] = Delta; fvals[ntr] = fx; gradnorms[ntr] = gnorm; }
  ++ntr;                                                                   /* l.617-621 */
  memcpy(x_out, x, nx * sizeof(double));
  *status = st; *n_outer = no; *n_trace = ntr;
  scalars[0] = fx; scalars[1] = gnorm; scalars[2] = pgnorm; scalars[3] = 0.0;
  free(x); free(xp); free(grad); free(h); free(tmp);
  return 0;
}

/* ---- the reference's S^2 test problem (tests/TNT_unit_test.cpp:63-122) ---- */
typedef struct { double P[3]; } s2_ctx;
static void s2_project(const double *X, const double *W, double *out) {
  const double c = dot8(X, W, 3);
  for (int i = 0; i < 3; ++i) out[i] = W[i] - c * X[i];
}
static double s2_f(void *c, const double *X) {
  const s2_ctx *S = (const s2_ctx *)c;
  double s = 0;
  for (int i = 0; i < 3; ++i) s += (X[i] - S->P[i]) * (X[i] - S->P[i]);
  return s;
}
static void s2_grad(void *c, const double *X, double *g) {
  const s2_ctx *S = (const s2_ctx *)c;
  double nabla[3];
  for (int i = 0; i < 3; ++i) nabla[i] = 2 * (X[i] - S->P[i]);
  s2_project(X, nabla, g);
}
static void s2_hess(void *c, const double *X, const double *V, double *out) {
  double EH[3], g[3];
  for (int i = 0; i < 3; ++i) EH[i] = 2.0 * V[i];
  s2_project(X, EH, out);
  s2_grad(c, X, g);                        /* Riemannian gradient: the reference's quirk, l.96 */
  const double cc = dot8(X, g, 3);
  for (int i = 0; i < 3; ++i) out[i] -= cc * V[i];
}
static double s2_metric(void *c, const double *X, const double *a, const double *b) {
  (void)c; (void)X;
  return dot8(a, b, 3);
}
static void s2_retract(void *c, const double *X, const double *V, double *out) {
  (void)c;
  for (int i = 0; i < 3; ++i) out[i] = X[i] + V[i];
  const double nrm = sqrt(dot8(out, out, 3));
  for (int i = 0; i < 3; ++i) out[i] /= nrm;
}
static void s2_precon(void *c, const double *X, const double *V, double *out) {
  (void)c; (void)X;
  out[0] = 1.0 * V[0]; out[1] = 2.0 * V[1]; out[2] = 3.0 * V[2];
}

int port_s2_tnt(const double *x0, const double *Ppt, int use_precon, const double *prm,
                double *x_out, int *status, uint64_t *n_outer, uint64_t *n_trace, double *scalars,
                uint64_t cap, uint64_t *inner_iterations, double *radius, double *rho,
                double *fvals, double *gradnorms, double *step_norms, double *step_M_norms) {
  s2_ctx S;
  memcpy(S.P, Ppt, sizeof S.P);
  port_manifold M = {3, 3, s2_f, s2_grad, s2_hess, s2_metric, s2_retract,
                     use_precon ? s2_precon : NULL, &S};
  return port_tnt(&M, x0, prm, x_out, status, n_outer, n_trace, scalars, cap, inner_iterations,
                  radius, rho, fvals, gradnorms, step_norms, step_M_norms);
}

/* ---- sparse Hessian families (configs C5 / C4; operator definitions in sparse_ops.h) ------------------- */
#include "sparse_ops.h"

typedef struct {
  uint64_t N; int r;
  const uint64_t *rowptr; const uint32_t *colidx; const double *blocks, *lambda, *X;
} port_csr3;
static void csr3_apply_cb(void *c, const double *v, double *out) {
  const port_csr3 *P = (const port_csr3 *)c;
  csr3_hess_apply(P->N, P->r, P->rowptr, P->colidx, P->blocks, P->lambda, P->X, v, out);
}
/* f = tr(X^T Q X); lambda: N x 9; grad: 3N x r (nullable) */
double port_csr3_model(uint64_t N, uint64_t r, const uint64_t *rowptr, const uint32_t *colidx, const double *blocks,
                       const double *X, double *lambda, double *grad) {
  return csr3_model(N, (int)r, rowptr, colidx, blocks, X, lambda, grad);
}
void port_csr3_hess(uint64_t N, uint64_t r, const uint64_t *rowptr, const uint32_t *colidx, const double *blocks,
                    const double *lambda, const double *X, const double *v, double *out) {
  csr3_hess_apply(N, (int)r, rowptr, colidx, blocks, lambda, X, v, out);
}
int port_stpcg_csr3(uint64_t N, uint64_t r, const uint64_t *rowptr, const uint32_t *colidx, const double *blocks,
                    const double *lambda, const double *X, const double *g, const double *minv, double Delta,
                    uint64_t max_iterations, double kappa_fgr, double theta, double epsilon, double *s,
                    double *mnorm, uint64_t *iters) {
  if (r < 3 || r > 8) return PORT_EXIT_BADARG;
  port_csr3 P = {N, (int)r, rowptr, colidx, blocks, lambda, X};
  port_diag Pd = {3 * N * r, minv};
  return port_stpcg(3 * N * r, g, csr3_apply_cb, &P, minv ? diag_apply : NULL, &Pd, Delta, max_iterations,
                    kappa_fgr, theta, epsilon, s, mnorm, iters);
}

typedef struct { uint32_t gx, gy, gz; int p; } port_stencil;
static void stencil_apply_cb(void *c, const double *v, double *out) {
  const port_stencil *P = (const port_stencil *)c;
  stencil7_apply(P->gx, P->gy, P->gz, P->p, v, out);
}
void port_stencil7_apply(uint32_t gx, uint32_t gy, uint32_t gz, uint64_t p, const double *v, double *out) {
  stencil7_apply(gx, gy, gz, (int)p, v, out);
}
int port_stpcg_stencil7(uint32_t gx, uint32_t gy, uint32_t gz, uint64_t p, const double *g, const double *minv,
                        double Delta, uint64_t max_iterations, double kappa_fgr, double theta, double epsilon,
                        double *s, double *mnorm, uint64_t *iters) {
  port_stencil P = {gx, gy, gz, (int)p};
  const uint64_t n = (uint64_t)gx * gy * gz * p;
  port_diag Pd = {n, minv};
  return port_stpcg(n, g, stencil_apply_cb, &P, minv ? diag_apply : NULL, &Pd, Delta, max_iterations, kappa_fgr,
                    theta, epsilon, s, mnorm, iters);
}
