/* optimization_b200.h -- C ABI of liboptimization_b200.so
 *
 * Drop-in boundary for the ONE hot path of david-m-rosen/Optimization that this
 * library accelerates: the Steihaug-Toint truncated preconditioned CG inner
 * loop (reference include/Optimization/LinearAlgebra/IterativeSolvers.h:166-426)
 * as driven by the Riemannian truncated-Newton trust-region method (reference
 * include/Optimization/Riemannian/TNT.h:242-689).
 *
 * The reference is a header-only template library with no ABI of its own: its
 * "plugin" surface is the std::function functor set of
 * include/Optimization/Riemannian/Concepts.h:44-112 and
 * include/Optimization/LinearAlgebra/Concepts.h:16-26.  The C++ host templates
 * shipped in include/Optimization/ (same include paths and signatures as the
 * reference) call the entry points below whenever the tangent type is the
 * device matrix handle `Optimization::b200::DeviceMatrix`; each entry point
 * cites the reference statement(s) it replaces.
 *
 * Conventions: plain pointers and sizes only; every function returns an int
 * status (OB200_OK == 0) and never throws; `*_dev` pointers are device memory on
 * the context's GPU, everything else is host memory; all work is issued on the
 * context's stream and functions that return scalars synchronise that stream.
 * Matrices are row-major n x p doubles (leading dimension p).
 */
#ifndef OPTIMIZATION_B200_H
#define OPTIMIZATION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes -------------------------------------------------------- */
#define OB200_OK 0
#define OB200_INVALID_ARGUMENT 1 /* host templates map this to std::invalid_argument
                                    (reference IterativeSolvers.h:183-205, TNT.h:260-318) */
#define OB200_CUDA_ERROR 2
#define OB200_UNSUPPORTED 3
#define OB200_NUMERIC_RANGE 4 /* fixed-point Gram bound exceeded / non-finite data */
#define OB200_ABORTED 5       /* device watchdog fired (grid barrier timeout) */

/* ---- tCG exit reasons (control-flow exits of IterativeSolvers.h:285-426) -- */
#define OB200_EXIT_RESIDUAL 0       /* l.290: ||r||_P <= target                      */
#define OB200_EXIT_MAX_ITERATIONS 1 /* l.285: loop bound reached                     */
#define OB200_EXIT_KERNEL 2         /* l.305-337: p in ker(H), stepped to the boundary */
#define OB200_EXIT_BOUNDARY 3       /* l.347-361: kappa <= 0 or step leaves the region */

typedef struct ob200_context ob200_context;

/* Create a context on `device` (CUDA ordinal).  Fails loudly (OB200_CUDA_ERROR) if
 * there is no usable GPU: there is no CPU fallback anywhere in this library.
 *
 * Stream contract.  Every kernel and copy of the library is issued on ONE stream:
 *   stream == NULL            the context creates its own cudaStreamNonBlocking stream.  It has NO implicit
 *                             ordering with the legacy default stream: a caller that produces inputs or consumes
 *                             outputs on another stream must order the two itself (ob200_synchronize, events).
 *   stream == cudaStreamLegacy ((void *)0x1) or cudaStreamPerThread ((void *)0x2)
 *                             the CUDA default stream of that flavour (what a caller whose own work runs on
 *                             "stream 0" must pass: a literal 0 cannot be told apart from NULL).
 *   any other cudaStream_t    work is issued on the caller's stream, in order with the caller's own work.
 * Entries that return scalars to the host (ob200_dot, ob200_stpcg, the model calls ...) synchronise that stream
 * before returning; entries that only enqueue work (ob200_axpby, ob200_hadamard, ob200_block_apply, ob200_hvp)
 * do not: their outputs are ordered after them on the same stream. */
int ob200_create(int device, void *stream, ob200_context **ctx);
int ob200_destroy(ob200_context *ctx);
const char *ob200_last_error(const ob200_context *ctx);
int ob200_version(void);
int ob200_sm_count(const ob200_context *ctx);
int ob200_synchronize(ob200_context *ctx);
/* number of kernels this context has launched (bench.py's gpu_launches) */
uint64_t ob200_kernel_launches(const ob200_context *ctx);

/* Options.  "tcgen05" (default 1): run the block contraction A*p of
 * OB200_OP_STIEFEL_BLOCKDIAG on the 5th-generation tensor cores through the exact bf16
 * digit-plane scheme whenever every block of A is 22-bit block-fixed-point (tcg_stiefel_v5_kernel);
 * 2 selects the previous generation of that kernel (tcg_stiefel_tc_kernel, kept for A/B measurements);
 * 0 forces the fp64 tensor-core (mma.sync) kernel.  ob200_last_path: which of them the last ob200_stpcg
 * ran (1, 2 or 0). */
int ob200_set_option(ob200_context *ctx, const char *name, int value);
int ob200_last_path(const ob200_context *ctx);

/* Profiling aid: when enabled, the persistent tCG kernels accumulate, per CTA, the
 * nanoseconds spent in {phase A, A reduction+barrier, phase B, B reduction+barrier};
 * a call with non-null outputs returns the max / min over CTAs and resets them. */
int ob200_debug_phase_times(ob200_context *ctx, int enable, uint64_t *out4_max, uint64_t *out4_min);

/* Validation aid: out = A V for the block-diagonal bf16 A (n x 32 V), either on the
 * fp64 tensor cores (use_tcgen05 = 0, mma.sync DMMA) or through the exact bf16
 * digit-plane scheme on tcgen05 (use_tcgen05 = 1: one TMEM lane per thread; 2: the 16-lane
 * fragment read-back the persistent kernel uses). */
int ob200_debug_block_apply(ob200_context *ctx, uint64_t n, const uint16_t *A_bf16_dev, const double *V_dev,
                            double *out_dev, int use_tcgen05);

/* Validation aid (host only, no GPU needed): the eigen-decomposition S = Q diag(lambda) Q^T (32 x 32, row-major, eigenvector
 * k in column k of Q) that ob200_stpcg uses to run the Stiefel solve in the eigenbasis of S = sym(Y^T A Y). */
int ob200_debug_sym_eig32(const double *S_host, double *Q_host, double *lambda_host);

/* ---- Hessian operator descriptors -----------------------------------------
 * Replaces the user's `Riemannian::LinearOperator` Hessian functor
 * (reference Riemannian/Concepts.h:49-51, bound to x at TNT.h:400-403 and
 * called at IterativeSolvers.h:294 and TNT.h:512) for the recognised operator
 * families, so the Hessian-vector product can be fused into the CG step. */
#define OB200_OP_DIAG 1             /* H v = d .* v                                   */
#define OB200_OP_STIEFEL_BLOCKDIAG 2 /* H V = P_Y(A V - V S), P_Y(Z) = Z - Y sym(Y^T Z),
                                        A block-diagonal dense (bf16 storage), p == 32  */
#define OB200_OP_SPHERE_LOWRANK 3   /* H v = 2 P_x(A v) - 2 (x^T A x) v, P_x(z) = z - x (x^T z),
                                        A = diag(d) + U diag(sigma) U^T, k <= 16, p == 1 */
#define OB200_OP_BLOCK_CSR3 4       /* rotation synchronisation on SO(3)^N relaxed to St(3,r)^N (SE-Sync's Hessian shape,
                                        BASELINE config C5): X (= Y_dev) is 3N x r row-major, pose i = rows 3i..3i+2 with
                                        X_i X_i^T = I_3;  H V = Proj_X(2 Q V - Lambda V), Proj_X(Z)_i = Z_i - sym(Z_i X_i^T) X_i,
                                        Q symmetric with 3 x 3 blocks in block-CSR, Lambda_i = sym((2 Q X)_i X_i^T)
                                        (ob200_csr3_model fills it); n == 3 N, 3 <= p == r <= 8 */
#define OB200_OP_STENCIL7 5         /* H V = 7-point Dirichlet Laplacian on a gx x gy x gz grid (x fastest) applied to each
                                        of the p columns of V; n == gx gy gz (BASELINE config C4's operator as a Hessian) */
#define OB200_OP_HOST_CALLBACK 6    /* any other Hessian: `apply` is called once per CG iteration (reference
                                        IterativeSolvers.h:294 with an opaque functor).  UNFUSED fallback: ob200_stpcg then
                                        runs the Steihaug-Toint loop on the host over the device level-1 kernels (exact
                                        reductions), statement for statement IterativeSolvers.h:211-424 */

/* out = H v (or v = P r for a preconditioner callback) on n*p doubles in DEVICE memory.  The library synchronises the
 * context's stream before the call; work the callback launches must be complete, or ordered on the context's stream,
 * when it returns.  Return 0 on success (anything else aborts the solve with OB200_ABORTED). */
typedef int (*ob200_apply_fn)(void *user, const double *in_dev, double *out_dev);

typedef struct {
  int kind;
  uint64_t n; /* rows */
  uint64_t p; /* columns (1 for plain vectors) */
  /* OB200_OP_DIAG: d (n*p, device).  OB200_OP_SPHERE_LOWRANK: d (n, device). */
  const double *diag_dev;
  /* OB200_OP_STIEFEL_BLOCKDIAG */
  const uint16_t *A_bf16_dev; /* ceil(n/128) blocks of 128x128 bf16, row-major, zero padded */
  const double *Y_dev;        /* n x p base point, orthonormal columns */
  const double *S_host;       /* p x p, sym(Y^T A Y) (host memory; see ob200_stiefel_model) */
  double op_norm_bound;       /* >= ||A||_2 + ||S||_2 (ob200_stiefel_model returns one) */
  /* OB200_OP_SPHERE_LOWRANK */
  const double *x_dev;     /* n, unit vector (base point) */
  const double *U_dev;     /* U TRANSPOSED: k rows of ldu doubles (row j = column j of U), 16-byte aligned */
  const double *sigma_host; /* k */
  uint64_t k;
  double xAx;              /* x^T A x   (ob200_sphere_model returns it) */
  const double *Ax_dev;    /* n, A x    (ob200_sphere_model fills it)   */
  uint64_t ldu;            /* row stride of U_dev in doubles: even and >= n; 0 means n (n must then be even) */
  /* OB200_OP_BLOCK_CSR3 (X in Y_dev) */
  const uint64_t *csr_rowptr_dev; /* N + 1 */
  const uint32_t *csr_colidx_dev; /* nnz block columns (pose indices), ascending within a row */
  const double *csr_blocks_dev;   /* nnz x 9, row-major 3 x 3 */
  const double *csr_lambda_dev;   /* N x 9 */
  uint64_t csr_nnz;               /* number of stored 3 x 3 blocks (for the byte counts) */
  /* OB200_OP_STENCIL7 */
  uint32_t gx, gy, gz;
  /* OB200_OP_BLOCK_CSR3, row-sharded over the ranks of ob200_comm_connect (each rank owns a contiguous range of poses,
   * a multiple of 256; n = 3 x local poses).  Column indices are LOCAL: [0, local poses) own poses, then `csr_n_halo`
   * halo poses (rows of other ranks that local rows refer to); X (Y_dev) holds own poses followed by the halo poses.
   * Before each operator apply every rank pushes the rows of p its peers need straight into their halo buffers
   * (ob200_halo_create / ob200_halo_connect) over NVLink from inside the kernel. */
  uint64_t csr_n_halo;                 /* halo poses of this rank (0: not sharded) */
  const uint32_t *halo_send_idx_dev;   /* own pose indices to push, grouped by destination rank */
  uint64_t halo_send_ptr[9];           /* entries [ptr[q], ptr[q+1]) of halo_send_idx_dev go to rank q */
  uint64_t halo_dst_off[8];            /* first halo slot (in poses) of my rows inside rank q's halo buffer */
  /* OB200_OP_HOST_CALLBACK */
  ob200_apply_fn apply;
  void *apply_user;
} ob200_operator;

/* Replaces the optional preconditioner functor (reference
 * IterativeSolvers.h:83-85, adapter TNT.h:413-426). */
#define OB200_PRECON_NONE 0
#define OB200_PRECON_JACOBI 1 /* v = minv .* r */
#define OB200_PRECON_HOST_CALLBACK 2 /* v = apply(user, r): any other preconditioner (selects the unfused loop) */
#define OB200_PRECON_STIEFEL_PROJECTED_JACOBI 3 /* v = P_Y(minv .* r), P_Y(Z) = Z - Y sym(Y^T Z), Y = the operator's Y_dev:
                                                  * the tangent-space preserving form of the Jacobi scaling for
                                                  * OB200_OP_STIEFEL_BLOCKDIAG (what TNT's adapter, TNT.h:413-426, hands
                                                  * to STPCG for precon(Y, V) = P_Y(minv o V)); runs the unfused loop
                                                  * with the one-launch device HVP */
typedef struct {
  int kind;
  const double *minv_dev; /* n*p */
  ob200_apply_fn apply;   /* OB200_PRECON_HOST_CALLBACK */
  void *apply_user;
} ob200_precon;

typedef struct {
  double Delta;            /* trust-region radius, > 0          (IterativeSolvers.h:171,183) */
  uint64_t max_iterations; /* default 1000                      (l.172) */
  double kappa_fgr;        /* in [0,1), default .1              (l.172,191) */
  double theta;            /* in [0,1], default .5              (l.172,196) */
  double epsilon;          /* in (0,1), default 1e-8            (l.179,201) */
} ob200_stpcg_params;

typedef struct {
  double update_step_M_norm; /* IterativeSolvers.h:334,359,424 */
  uint64_t num_iterations;   /* IterativeSolvers.h:285 */
  int exit_reason;           /* OB200_EXIT_* */
  double r0_norm;            /* l.275 */
  double final_rv;           /* last <r, v> */
  uint64_t kernel_launches;  /* kernels launched by this call */
  float solve_kernel_ms;     /* device time of the persistent fused tCG kernel alone
                                (CUDA events on the context's stream around its launch) */
} ob200_stpcg_result;

/* Steihaug-Toint truncated preconditioned CG, whole solve on the device.
 * Mirrors  Vector STPCG(g, H, inner_product, update_step_M_norm, num_iterations,
 * Delta, max_iterations, kappa_fgr, theta, P, At, user_function, epsilon)
 * (reference IterativeSolvers.h:166-179) with the Frobenius inner product
 * (every in-tree metric: Riemannian/Concepts.h:181-185), no constraint operator
 * `At` and no per-iteration host hook -- exactly how TNT calls it (TNT.h:489-492).
 * g_dev, s_dev: n*p doubles on the device.  s_dev receives the step. */
int ob200_stpcg(ob200_context *ctx, const ob200_operator *H, const ob200_precon *P,
                const double *g_dev, const ob200_stpcg_params *params, double *s_dev,
                ob200_stpcg_result *result);

/* Same call with HOST buffers for g and s (pinned or pageable): the copies are
 * part of the call.  This is the end-to-end entry bench.py times as `e2e`. */
int ob200_stpcg_host(ob200_context *ctx, const ob200_operator *H, const ob200_precon *P,
                     const double *g_host, const ob200_stpcg_params *params, double *s_host,
                     ob200_stpcg_result *result);

/* Stand-alone Hessian-vector product out = H(v) (reference call sites
 * IterativeSolvers.h:294, TNT.h:512). */
int ob200_hvp(ob200_context *ctx, const ob200_operator *H, const double *v_dev, double *out_dev);

/* Algorithmic HBM bytes of ONE fused tCG step / one stand-alone HVP for this
 * operator (the roofline numerator; definition in DESIGN.md section 4):
 *   step = 10 N e + B_op (+ 2 N e with Jacobi),  hvp = 2 N e + B_op,  e = 8;
 *   B_op: DIAG N e; STIEFEL_BLOCKDIAG 2 N e (Y twice) + 2 bytes per stored entry of A;
 *   SPHERE_LOWRANK (2k + 4) N e (U twice, d, x, A x, p re-read by the second pass);
 *   BLOCK_CSR3 N e (p re-read by the operator pass) + nnz (72 + 4) + N_poses (8 + 72) + N e (X); STENCIL7 N e. */
uint64_t ob200_stpcg_step_bytes(const ob200_operator *H, const ob200_precon *P);
uint64_t ob200_hvp_bytes(const ob200_operator *H);

/* ---- level-1 primitives with the exact, order-independent reduction --------
 * Replace `metric(x, a, b)` (TNT.h:382,387,493,511-512,575,579) and the
 * generic vector expressions of the Krylov loops when the user's Hessian is an
 * arbitrary functor over device matrices. */
int ob200_dot(ob200_context *ctx, uint64_t n, const double *a_dev, const double *b_dev,
              double *result);
/* up to 4 inner products <a_i, b_i> in one pass */
int ob200_dots(ob200_context *ctx, uint64_t n, int count, const double *const *a_dev,
               const double *const *b_dev, double *results);
/* out = alpha * x + beta * y   (out may alias x or y) */
int ob200_axpby(ob200_context *ctx, uint64_t n, double alpha, const double *x_dev, double beta,
                const double *y_dev, double *out_dev);
/* out = x / a, one IEEE division per element (out may alias x): the `v /= Scalar` of the reference's LSQR / TNLS
 * loops (IterativeSolvers.h:707-799, TNLS.h) on device vectors */
int ob200_div(ob200_context *ctx, uint64_t n, const double *x_dev, double a, double *out_dev);
/* out = d .* x */
int ob200_hadamard(ob200_context *ctx, uint64_t n, const double *d_dev, const double *x_dev,
                   double *out_dev);

/* ---- Stiefel trace-minimisation model on the device -------------------------
 * f(Y) = 1/2 tr(Y^T A Y).  These replace the user's Objective / QuadraticModel /
 * Retraction functors (reference Base/Concepts.h:37-38, Riemannian/Concepts.h:
 * 63-67,110-112; call sites TNT.h:377,380,505,508,573) for this model so the
 * outer loop never copies n x p matrices to the host. */
/* S = sym(Y^T A Y) (host, p*p), f = 1/2 tr(S), optional grad = A Y - Y S (device),
 * op_norm_bound = ||A||_inf + ||S||_F. */
int ob200_stiefel_model(ob200_context *ctx, uint64_t n, uint64_t p, const uint16_t *A_bf16_dev,
                        const double *Y_dev, double *S_host, double *f, double *grad_dev,
                        double *op_norm_bound);
/* Cholesky-QR retraction  out = qf(Y + V) */
int ob200_stiefel_retract(ob200_context *ctx, uint64_t n, uint64_t p, const double *Y_dev,
                          const double *V_dev, double *out_dev);
/* Tangent-space projection  out = Z - Y sym(Y^T Z)  at a point Y with orthonormal columns (p == 32).  Building block of
 * tangent-space preserving preconditioners for the Stiefel model: the reference hands the user's `precon` functor
 * (TNT.h:247) to STPCG through the adapter of TNT.h:413-426, and an elementwise scaling alone leaves T_Y St(n,p). */
int ob200_stiefel_project(ob200_context *ctx, uint64_t n, uint64_t p, const double *Y_dev,
                          const double *Z_dev, double *out_dev);

/* ---- rotation synchronisation f(X) = tr(X^T Q X) on St(3,r)^N (BASELINE config C5) -----------------------------
 * Replace the Objective / QuadraticModel / Retraction functors of that model (reference call sites TNT.h:377,380,505,
 * 508,573).  ob200_csr3_model: lambda_dev <- Lambda (N x 9, the `csr_lambda_dev` field of the Hessian descriptor),
 * *f <- f(X), optional grad_dev <- (2 Q X) - Lambda X.  ob200_csr3_retract: out_i = rows of X_i + V_i re-orthonormalised
 * (Gram-Schmidt in row order = Q factor of the QR decomposition of (X_i + V_i)^T). */
int ob200_csr3_model(ob200_context *ctx, uint64_t N, uint64_t r, const uint64_t *rowptr_dev, const uint32_t *colidx_dev,
                     const double *blocks_dev, const double *X_dev, double *lambda_dev, double *f, double *grad_dev);
int ob200_csr3_retract(ob200_context *ctx, uint64_t N, uint64_t r, const double *X_dev, const double *V_dev,
                       double *out_dev);

/* ---- Rayleigh quotient on the sphere, f(x) = x^T A x, A = diag(d) + U diag(sigma) U^T ----
 * Replace the Objective / QuadraticModel / Retraction functors of the sphere model
 * (reference call sites TNT.h:377,380,505,508,573; the model of BASELINE configs C1 / C2,
 * shaped after examples/Riemannian_optimization_example.cpp).
 * Ax_dev <- A x (n doubles, also the `Ax_dev` field of the Hessian descriptor), *xAx <- f(x),
 * optional grad_dev <- 2 (A x - (x^T A x) x). */
int ob200_sphere_model(ob200_context *ctx, uint64_t n, uint64_t k, const double *d_dev, const double *Ut_dev,
                       uint64_t ldu, const double *sigma_host, const double *x_dev, double *Ax_dev, double *xAx,
                       double *grad_dev);
/* projection retraction  out = (x + v) / ||x + v||   (out may alias v) */
int ob200_sphere_retract(ob200_context *ctx, uint64_t n, const double *x_dev, const double *v_dev, double *out_dev);

/* ---- LOBPCG: smallest eigenpairs of A x = lambda B x  (reference LinearAlgebra/LOBPCG.h:131-337) -------------
 * Replaces  LOBPCG<Vector, Matrix>(A, B, T, X0, nev, max_iters, num_iters, nc, tau)  for block operators the library
 * knows (the `SymmetricLinearOperator<Matrix>` functors A, B, T of LOBPCG.h:131-140 become descriptors), with the
 * reference's iteration: basis S = [X, T R, P] with soft locking of the nc converged columns (l.207-221),
 * Rayleigh-Ritz on (S^T A S, S^T B S) with diagonal equilibration (l.53-62; the dense generalised eigensolve runs in
 * cuSOLVER, LIBRARY code: it stands for Eigen's GeneralizedSelfAdjointEigenSolver), X = S C, P = S_{W,P} C_{W,P},
 * R = A X - B X Theta, convergence test |r_i| <= tau (|A| + |B| |theta_i|) |x_i| on the first nev columns (l.254-269).
 * Block vectors are row-major m x k; X_dev: m x nx (in: X0, out: eigenvector estimates, first nev columns meaningful).
 * Omega_dev: m x nx probe block for the operator norm estimates of l.172-181 (NULL: X0 is used). */
#define OB200_BLK_DIAG 1     /* (A X)[r, :] = diag_dev[r] * X[r, :]                                      */
#define OB200_BLK_STENCIL7 2 /* 7-point Laplacian, Dirichlet boundary, grid gx x gy x gz (x fastest), m = gx gy gz */
#define OB200_BLK_SCALAR 3   /* A X = alpha X  (e.g. the Jacobi preconditioner 1/6 of the Laplacian)     */
typedef struct {
  int kind;
  const double *diag_dev; /* OB200_BLK_DIAG: m doubles */
  double alpha;           /* OB200_BLK_SCALAR */
  uint32_t gx, gy, gz;    /* OB200_BLK_STENCIL7 */
} ob200_block_operator;
int ob200_lobpcg(ob200_context *ctx, const ob200_block_operator *A, const ob200_block_operator *B /* NULL: identity */,
                 const ob200_block_operator *T /* NULL: none */, uint64_t m, uint64_t nx, double *X_dev, uint64_t nev,
                 uint64_t max_iters, double tau, const double *Omega_dev, double *theta_host /* nev */,
                 uint64_t *num_iters, uint64_t *num_converged);
/* out = Op(in) for a block operator (m x k row-major blocks with leading dimensions) */
int ob200_block_apply(ob200_context *ctx, const ob200_block_operator *Op, uint64_t m, uint64_t k, const double *in_dev,
                      uint64_t ldi, double *out_dev, uint64_t ldo);

/* ---- device memory helpers (so non-CUDA hosts can stage data) -------------- */
int ob200_malloc(ob200_context *ctx, size_t bytes, void **ptr_dev);
int ob200_free(ob200_context *ctx, void *ptr_dev);
int ob200_memcpy_h2d(ob200_context *ctx, void *dst_dev, const void *src_host, size_t bytes);
int ob200_memcpy_d2h(ob200_context *ctx, void *dst_host, const void *src_dev, size_t bytes);
int ob200_malloc_host(ob200_context *ctx, size_t bytes, void **ptr_host); /* pinned */
int ob200_free_host(ob200_context *ctx, void *ptr_host);

/* ---- multi-GPU (row-block sharding, one process per GPU) --------------------
 * Each rank holds a contiguous row block (a multiple of 128 rows) of every n x p
 * matrix and passes its LOCAL row count / pointers to the calls above.  All
 * reductions are exchanged as exact integer accumulators through NVLink peer
 * memory from inside the kernels (no host round trip, no NCCL call on the data
 * path), so results are bit-identical for any world size.  Set-up: every rank
 * exports a 64-byte CUDA IPC handle, the host exchanges the handles (e.g.
 * torch.distributed all_gather), every rank connects.  All ranks must then issue
 * the same sequence of library calls. */
#define OB200_COMM_HANDLE_BYTES 64
/* Halo buffer of a row-sharded sparse operator (peers store the rows of p this rank needs into it): allocate `bytes`
 * on this rank and export its 64-byte IPC handle; connect to the peers' buffers (handles in rank order, after
 * ob200_comm_connect). */
int ob200_halo_create(ob200_context *ctx, uint64_t bytes, void *handle_out /* 64 bytes */);
int ob200_halo_connect(ob200_context *ctx, const void *handles /* world_size x 64 bytes, rank order */);
int ob200_comm_export(ob200_context *ctx, void *handle_out /* 64 bytes */);
int ob200_comm_connect(ob200_context *ctx, int rank, int world_size,
                       const void *handles /* world_size x 64 bytes, rank order */);
int ob200_comm_rank(const ob200_context *ctx);
int ob200_comm_world(const ob200_context *ctx);

#ifdef __cplusplus
}
#endif
#endif /* OPTIMIZATION_B200_H */
