/* TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
 *
 * Operator restatements shared by the two CPU oracles (oracle/stpcg_port.c and oracle/ref_driver.cpp) for the
 * sparse Hessian families of BASELINE configs C5 and C4.  The reference contains no such
 * operators: like the Stiefel / sphere models they are user functors (`Riemannian::LinearOperator`,
 * /root/reference/include/Optimization/Riemannian/Concepts.h:49-51) that the reference's STPCG merely calls
 * (IterativeSolvers.h:294).  They are defined here, once, in plain C; the arithmetic order is the one the CUDA
 * kernels use (csrc/tcg_sparse.cu), so results agree to rounding.
 *
 * C5 -- rotation synchronisation on SO(3)^N relaxed to St(3, r)^N (SE-Sync's rank-r relaxation, [external
 * formulation]):  X in R^{3N x r}, row-major, pose i = rows 3i .. 3i+2 with X_i X_i^T = I_3;
 *   f(X)     = tr(X^T Q X),  Q = connection Laplacian, symmetric, 3 x 3 blocks in block-CSR
 *   G        = 2 Q X                                   (Euclidean gradient)
 *   Lambda_i = sym(G_i X_i^T)                          (3 x 3, per pose)
 *   grad     = G - Lambda X                            (blockwise)
 *   Hess[V]  = Proj_X(2 Q V - Lambda V),  Proj_X(Z)_i = Z_i - sym(Z_i X_i^T) X_i
 * C4 operator as a tCG Hessian: 7-point Laplacian with Dirichlet boundary on a gx x gy x gz grid (x fastest),
 * applied to every column of an n x p matrix (n = gx gy gz).
 */
#ifndef ORACLE_SPARSE_OPS_H
#define ORACLE_SPARSE_OPS_H
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* sum over the columns of a pose with the association of the kernel's xor-shuffle tree over LPP = 4 or 8 lanes */
static inline double csr3_colsum(const double *t /* r values */, int r) {
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = 0; c < r; ++c) v[c] = t[c];
  if (r <= 4) return (v[0] + v[1]) + (v[2] + v[3]);
  return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}

/* Z_i = 2 sum_j Q_ij V_j  for one pose (3 x r), CSR order, one fma per term */
static inline void csr3_row_2qv(uint64_t i, int r, const uint64_t *rowptr, const uint32_t *colidx, const double *blocks,
                                const double *V, double *Z /* 3 x r */) {
  for (int c = 0; c < r; ++c) {
    double z0 = 0.0, z1 = 0.0, z2 = 0.0;
    for (uint64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      const double *B = blocks + 9 * e;
      const double *Vj = V + (size_t)3 * colidx[e] * r + c;
      const double v0 = Vj[0], v1 = Vj[r], v2 = Vj[2 * r];
      z0 = fma(B[0], v0, z0); z0 = fma(B[1], v1, z0); z0 = fma(B[2], v2, z0);
      z1 = fma(B[3], v0, z1); z1 = fma(B[4], v1, z1); z1 = fma(B[5], v2, z1);
      z2 = fma(B[6], v0, z2); z2 = fma(B[7], v1, z2); z2 = fma(B[8], v2, z2);
    }
    Z[c] = 2.0 * z0; Z[r + c] = 2.0 * z1; Z[2 * r + c] = 2.0 * z2;
  }
}

/* out_i = Proj_X(W)_i = W_i - sym(W_i X_i^T) X_i */
static inline void csr3_project(int r, const double *Xi, const double *W, double *out) {
  double M[9], t[8];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      for (int c = 0; c < r; ++c) t[c] = W[a * r + c] * Xi[b * r + c];
      M[3 * a + b] = csr3_colsum(t, r);
    }
  double S[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) S[3 * a + b] = 0.5 * (M[3 * a + b] + M[3 * b + a]);
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < r; ++c) {
      double h = W[a * r + c];
      h = fma(-S[3 * a + 0], Xi[c], h);
      h = fma(-S[3 * a + 1], Xi[r + c], h);
      h = fma(-S[3 * a + 2], Xi[2 * r + c], h);
      out[a * r + c] = h;
    }
}

/* out = Hess f(X)[V] */
static inline void csr3_hess_apply(uint64_t N, int r, const uint64_t *rowptr, const uint32_t *colidx,
                                   const double *blocks, const double *lambda, const double *X, const double *V,
                                   double *out) {
  for (uint64_t i = 0; i < N; ++i) {
    double Z[24], W[24];
    csr3_row_2qv(i, r, rowptr, colidx, blocks, V, Z);
    const double *L = lambda + 9 * i, *Vi = V + (size_t)3 * i * r;
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < r; ++c) {
        double w = Z[a * r + c];
        w = fma(-L[3 * a + 0], Vi[c], w);
        w = fma(-L[3 * a + 1], Vi[r + c], w);
        w = fma(-L[3 * a + 2], Vi[2 * r + c], w);
        W[a * r + c] = w;
      }
    csr3_project(r, X + (size_t)3 * i * r, W, out + (size_t)3 * i * r);
  }
}

/* model at X: Lambda (N x 9), grad (3N x r, nullable); returns f = tr(X^T Q X) = 1/2 <X, G> (plain ordered sum) */
static inline double csr3_model(uint64_t N, int r, const uint64_t *rowptr, const uint32_t *colidx, const double *blocks,
                                const double *X, double *lambda, double *grad) {
  double f = 0.0;
  for (uint64_t i = 0; i < N; ++i) {
    double G[24], t[8];
    csr3_row_2qv(i, r, rowptr, colidx, blocks, X, G);
    const double *Xi = X + (size_t)3 * i * r;
    double M[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        for (int c = 0; c < r; ++c) t[c] = G[a * r + c] * Xi[b * r + c];
        M[3 * a + b] = csr3_colsum(t, r);
      }
    f += 0.5 * ((M[0] + M[4]) + M[8]);
    double *L = lambda + 9 * i;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) L[3 * a + b] = 0.5 * (M[3 * a + b] + M[3 * b + a]);
    if (grad) {
      double *gi = grad + (size_t)3 * i * r;
      for (int a = 0; a < 3; ++a)
        for (int c = 0; c < r; ++c) {
          double h = G[a * r + c];
          h = fma(-L[3 * a + 0], Xi[c], h);
          h = fma(-L[3 * a + 1], Xi[r + c], h);
          h = fma(-L[3 * a + 2], Xi[2 * r + c], h);
          gi[a * r + c] = h;
        }
    }
  }
  return f;
}

/* out = A V, A = 7-point Dirichlet Laplacian on gx x gy x gz (x fastest), V: n x p row-major.
 * Per element: 6 v - west - east - south - north - down - up, subtractions in that order (missing neighbours skipped). */
static inline void stencil7_apply(uint32_t gx, uint32_t gy, uint32_t gz, int p, const double *V, double *out) {
  const size_t sx = (size_t)p, sy = (size_t)gx * p, sz = (size_t)gx * gy * p;
  for (uint32_t z = 0; z < gz; ++z)
    for (uint32_t y = 0; y < gy; ++y)
      for (uint32_t x = 0; x < gx; ++x) {
        const size_t o = ((size_t)z * gy + y) * gx * p + (size_t)x * p;
        for (int c = 0; c < p; ++c) {
          double h = 6.0 * V[o + c];
          if (x > 0) h -= V[o + c - sx];
          if (x + 1 < gx) h -= V[o + c + sx];
          if (y > 0) h -= V[o + c - sy];
          if (y + 1 < gy) h -= V[o + c + sy];
          if (z > 0) h -= V[o + c - sz];
          if (z + 1 < gz) h -= V[o + c + sz];
          out[o + c] = h;
        }
      }
}
#endif
