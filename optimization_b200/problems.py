"""Synthetic problem generators.

The reference ships no data sets and no Stiefel / Rayleigh problems; these
generators define the synthetic inputs every test and benchmark uses.  All
randomness comes from a counter-based splitmix64 stream (index -> value), so
any slice of any array can be regenerated independently and identically on
every rank of a sharded run.  Pure numpy: the SAME arrays are handed to the CPU
oracle and to the CUDA path.
"""
from __future__ import annotations

import dataclasses

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """Counter-based generator: value = mix(seed * golden + idx)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + idx.astype(np.uint64)
             + np.uint64(0x632BE59BD9B4E019))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(seed: int, start: int, count: int) -> np.ndarray:
    """count doubles in [0,1), stream positions start .. start+count-1."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    return (splitmix64(seed, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def gaussish(seed: int, start: int, count: int) -> np.ndarray:
    """Zero-mean unit-variance bell-shaped samples (sum of 4 uniforms; no libm)."""
    acc = np.zeros(count)
    for k in range(4):
        acc += uniform01(seed * 4 + k + 1000003, start, count)
    return (acc - 2.0) * np.sqrt(3.0)


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float -> bf16 bit pattern; asserts the values are exactly representable."""
    f = np.ascontiguousarray(x, dtype=np.float32)
    bits = f.view(np.uint32)
    assert np.all((bits & np.uint32(0xFFFF)) == 0), "value not on the bf16 grid"
    return (bits >> np.uint32(16)).astype(np.uint16)


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """Round doubles to the nearest bf16 value (ties to even), returned as float64."""
    bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    bits = (bits + np.uint64(0x7FFF) + ((bits >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    return bits.astype(np.uint32).view(np.float32).astype(np.float64)


def from_bf16_bits(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << np.uint32(16)).view(np.float32).astype(np.float64)


@dataclasses.dataclass
class StiefelProblem:
    """Trace minimisation f(Y) = 1/2 tr(Y^T A Y) on St(n, p), A block-diagonal.

    A: nblk blocks of nb x nb, symmetric, every entry exactly bf16-representable
    (so bf16 storage is lossless and the CPU oracle holds the same matrix in
    double).  Rows of the last block beyond n are zero.
    """
    n: int
    p: int
    nb: int
    A_bf16: np.ndarray      # (nblk, nb, nb) uint16
    Y0: np.ndarray          # (n, p) float64, orthonormal columns
    g: np.ndarray           # (n, p) float64, tangent at Y0 (stand-alone tCG rhs)

    @property
    def nblk(self) -> int:
        return (self.n + self.nb - 1) // self.nb

    def A_dense_blocks(self) -> np.ndarray:
        return from_bf16_bits(self.A_bf16)


def stiefel_rows(n: int, nb: int, row0: int, row1: int, seed: int = 21,
                 diag_width: float = 28.0, low: np.ndarray | None = None, generic_bf16: bool = False) -> np.ndarray:
    """A blocks covering rows [row0,row1) (block aligned) as float64 (nblk_local, nb, nb).

    Block = diag(c) + R, c in [3, 3+diag_width) on a 1/4 grid, R symmetric with
    entries m * 2^-11, |m| <= 31 (all exactly bf16).  Rows/columns listed in
    `low` get diagonal 1 and zero off-diagonals, so span{e_i : i in low} is the
    exact minimising subspace of tr(Y^T A Y)."""
    assert row0 % nb == 0
    b0, b1 = row0 // nb, (row1 + nb - 1) // nb
    out = np.zeros((b1 - b0, nb, nb))
    iu = np.triu_indices(nb, 1)
    for b in range(b0, b1):
        base = b * nb * nb
        u = uniform01(seed, base, nb * nb).reshape(nb, nb)
        m = np.floor(u * 63.0) - 31.0            # integers in [-31, 31]
        R = np.zeros((nb, nb))
        if generic_bf16:
            # generic bf16 data: bell-shaped samples of scale 2^-7 rounded to bf16 (full 8-bit significands, exponents
            # spread over far more than 22 bits within a block: such an A is NOT block-fixed-point)
            R[iu] = round_to_bf16(gaussish(seed + 7, base, nb * nb).reshape(nb, nb) * 2.0 ** -7)[iu]
        else:
            R[iu] = (m * 2.0 ** -11)[iu]
        R = R + R.T
        assert 0.0 < diag_width <= 28.0
        dg = 3.0 + np.floor(u.diagonal() * diag_width * 4.0) / 4.0   # [3, 3+width) step 1/4
        blk = R + np.diag(dg)
        rows = min(nb, n - b * nb)
        if low is not None:
            for i in low[(low >= b * nb) & (low < (b + 1) * nb)] - b * nb:
                blk[i, :] = 0.0
                blk[:, i] = 0.0
                blk[i, i] = 1.0
        blk[rows:, :] = 0.0
        blk[:, rows:] = 0.0
        out[b - b0] = blk
    return out


def low_rows(n: int, p: int) -> np.ndarray:
    """The p row indices whose diagonal entry is lowered to 1 (the minimising subspace)."""
    return (np.arange(p) * (n // p) + min(5, n // p - 1)).astype(np.int64)


def make_stiefel(n: int, p: int = 32, nb: int = 128, seed: int = 21,
                 y_noise: float = 0.3, diag_width: float = 28.0) -> StiefelProblem:
    """Y0 = qf(E_L + (y_noise / sqrt(n)) * N): a point at relative distance
    ~y_noise from the minimising subspace; g = grad f(Y0) (computed by callers
    that need it) or the random tangent below."""
    L = low_rows(n, p)
    A = stiefel_rows(n, nb, 0, n, seed, diag_width, L)
    E = np.zeros((n, p))
    E[L, np.arange(p)] = 1.0
    Z = E + (y_noise / np.sqrt(n)) * gaussish(seed + 1, 0, n * p).reshape(n, p)
    Y0, Rq = np.linalg.qr(Z)
    Y0 = Y0 * np.sign(np.diag(Rq))[None, :]
    Y0 = np.ascontiguousarray(Y0)
    N = gaussish(seed + 2, 0, n * p).reshape(n, p)
    G = Y0.T @ N
    g = N - Y0 @ (0.5 * (G + G.T))
    return StiefelProblem(n, p, nb, to_bf16_bits(A), Y0, np.ascontiguousarray(g))


def make_stiefel_critical(n: int, p: int = 32, nb: int = 128, seed: int = 21,
                          diag_width: float = 28.0, generic_bf16: bool = False) -> StiefelProblem:
    """Stand-alone tCG workload (config C3 throughput runs): Y0 = E_L Q is an exact
    minimiser (E_L spans the eigenvalue-1 invariant subspace of A, Q a p x p
    rotation), so Hess f(Y0) = P(A V - V) is positive definite on the horizontal
    space (spectrum within [~1.7, ~31]) and CG runs a natural 50-80 iterations to
    1e-12; g is a horizontal tangent (Y0^T g = 0), like a Riemannian gradient.
    Away from critical points the trace-min Hessian on St(n,p) is indefinite (the
    cost is invariant under Y -> Y Q), and tCG leaves through the trust-region
    boundary after a handful of iterations: `make_stiefel` covers that regime."""
    L = low_rows(n, p)
    A = stiefel_rows(n, nb, 0, n, seed, diag_width, L, generic_bf16)
    Qm, Rq = np.linalg.qr(gaussish(seed + 5, 0, p * p).reshape(p, p))
    Qm = Qm * np.sign(np.diag(Rq))[None, :]
    Y0 = np.zeros((n, p))
    Y0[L, :] = Qm
    N = gaussish(seed + 2, 0, n * p).reshape(n, p)
    g = N - Y0 @ (Y0.T @ N)
    g = g - Y0 @ (Y0.T @ g)
    return StiefelProblem(n, p, nb, to_bf16_bits(A), np.ascontiguousarray(Y0), np.ascontiguousarray(g))


def stiefel_hess_numpy(prob: StiefelProblem, Y: np.ndarray, V: np.ndarray) -> np.ndarray:
    """Dense numpy Hess f(Y)[V] = P_Y(A V - V sym(Y^T A Y)); small n only."""
    A = prob.A_dense_blocks()
    n, p, nb = prob.n, prob.p, prob.nb

    def applyA(X):
        out = np.zeros_like(X)
        for b in range(prob.nblk):
            r0, r1 = b * nb, min(n, (b + 1) * nb)
            out[r0:r1] = A[b, : r1 - r0, : r1 - r0] @ X[r0:r1]
        return out
    AY = applyA(Y)
    S = Y.T @ AY
    S = 0.5 * (S + S.T)
    W = applyA(V) - V @ S
    G = Y.T @ W
    return W - Y @ (0.5 * (G + G.T))


@dataclasses.dataclass
class SphereProblem:
    """Rayleigh quotient f(x) = x^T A x on S^{n-1}, A = diag(d) + U diag(sigma) U^T."""
    n: int
    k: int
    d: np.ndarray
    U: np.ndarray
    sigma: np.ndarray
    x0: np.ndarray
    g: np.ndarray


def make_sphere(n: int, k: int = 16, seed: int = 11) -> SphereProblem:
    d = 1.0 + uniform01(seed, 0, n)
    U = gaussish(seed + 1, 0, n * k).reshape(n, k) / np.sqrt(n) if k else np.zeros((n, 0))
    sigma = 2.0 * uniform01(seed + 2, 0, k) - 1.0
    x0 = gaussish(seed + 3, 0, n)
    x0 /= np.linalg.norm(x0)
    z = gaussish(seed + 4, 0, n)
    g = z - x0 * float(x0 @ z)
    return SphereProblem(n, k, d, np.ascontiguousarray(U), sigma, x0, g)


def make_sphere_critical(n: int, k: int = 16, seed: int = 11, x_noise: float = 0.02) -> SphereProblem:
    """Same family with a known minimiser: d_0 = 0.25, d_i in [1.5, 2.5] otherwise, row 0 of U zero and
    |sigma| <= 0.5, so e_0 is the eigenvector of the smallest eigenvalue (0.25) and the Riemannian
    Hessian 2 (A - lambda I) is positive definite on the tangent space (spectrum within about
    [1.4, 5.6]): tCG runs its natural course to the residual target instead of leaving through the
    boundary at once.  x0 = normalised e_0 + noise of norm about x_noise; g = a tangent vector at x0."""
    d = 1.5 + uniform01(seed, 0, n)
    d[0] = 0.25
    U = gaussish(seed + 1, 0, n * k).reshape(n, k) / np.sqrt(n) if k else np.zeros((n, 0))
    if k:
        U[0, :] = 0.0
    sigma = uniform01(seed + 2, 0, k) - 0.5
    x0 = (x_noise / np.sqrt(n)) * gaussish(seed + 3, 0, n)
    x0[0] = 1.0
    x0 /= np.linalg.norm(x0)
    z = gaussish(seed + 4, 0, n)
    g = z - x0 * float(x0 @ z)
    return SphereProblem(n, k, d, np.ascontiguousarray(U), sigma, x0, g)


def _torch_uniform01(seed: int, start: int, count: int, device):
    """uniform01 on the device, bit-identical to the numpy version (int64 arithmetic wraps like uint64;
    logical right shifts are emulated with a mask)."""
    import torch

    def c(v):  # uint64 constant as a wrapped python int for int64 tensors
        v &= 0xFFFFFFFFFFFFFFFF
        return v - (1 << 64) if v >= (1 << 63) else v

    def shr(z, sh):
        return (z >> sh) & ((1 << (64 - sh)) - 1)

    idx = torch.arange(start, start + count, dtype=torch.int64, device=device)
    z = idx + c(seed * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019)
    z = (z ^ shr(z, 30)) * c(0xBF58476D1CE4E5B9)
    z = (z ^ shr(z, 27)) * c(0x94D049BB133111EB)
    z = z ^ shr(z, 31)
    return shr(z, 11).to(torch.float64) * (1.0 / (1 << 53))


def _torch_gaussish(seed: int, start: int, count: int, device):
    import torch
    acc = torch.zeros(count, dtype=torch.float64, device=device)
    for k in range(4):
        acc += _torch_uniform01(seed * 4 + k + 1000003, start, count, device)
    return (acc - 2.0) * float(np.sqrt(3.0))


def make_sphere_critical_device(n: int, k: int = 16, seed: int = 11, x_noise: float = 0.02, device="cuda:0"):
    """make_sphere_critical generated directly in device memory (full-size config C2: U alone is
    2 GB at n = 2^24).  Returns torch tensors (d, Ut [k x n], sigma (numpy), x0, g); d and U are
    bit-identical to the host generator, x0 / g agree up to the rounding of the normalisation."""
    import torch
    d = 1.5 + _torch_uniform01(seed, 0, n, device)
    d[0] = 0.25
    Ut = torch.empty((k, n), dtype=torch.float64, device=device)
    rows = max(1, (1 << 22) // max(k, 1))
    for r0 in range(0, n, rows):                  # U[r, j] = sample r * k + j
        r1 = min(n, r0 + rows)
        blk = _torch_gaussish(seed + 1, r0 * k, (r1 - r0) * k, device).view(r1 - r0, k)
        Ut[:, r0:r1] = (blk / float(np.sqrt(n))).t()
    if k:
        Ut[:, 0] = 0.0
    sigma = uniform01(seed + 2, 0, k) - 0.5
    x0 = float(x_noise / np.sqrt(n)) * _torch_gaussish(seed + 3, 0, n, device)
    x0[0] = 1.0
    x0 /= torch.linalg.norm(x0)
    z = _torch_gaussish(seed + 4, 0, n, device)
    g = z - x0 * torch.dot(x0, z)
    return d, Ut, sigma, x0, g


def make_projected(n: int = 50, mc: int = 3, seed: int = 61):
    """Equality-constrained quadratic model for the constraint-preconditioned STPCG tests (the shape of the
    reference's tests/IterativeSolvers_unit_test.cpp:316-496): H = diag(h) SPD, M = diag(m) SPD, A: mc x n, g."""
    h = 1000.0 + 2000.0 * uniform01(seed, 0, n)
    m = 1000.0 + 2000.0 * uniform01(seed + 1, 0, n)
    A = 1000.0 * (2.0 * uniform01(seed + 2, 0, mc * n) - 1.0).reshape(mc, n)
    g = 2.0 * uniform01(seed + 3, 0, n) - 1.0
    return h, m, np.ascontiguousarray(A), g


LSQR_A43 = np.array([[10., 5., 10.], [2., 9., 8.], [10., 2., 10.], [10., 5., 7.]])   # the 4 x 3 system of the
LSQR_B4 = np.array([1., 9., 10., 2.])                                                # reference's LSQR tests


def lsqr_cases():
    """LSQR problems: the five cases of the reference's tests/IterativeSolvers_unit_test.cpp:517-700 and a larger
    seeded 60 x 40 system (plain, Tikhonov, trust-region).  name -> (A, b, kwargs of LSQR)."""
    A0 = np.zeros((3, 2))
    A0[1:, :] = np.eye(2)
    xls = np.linalg.lstsq(LSQR_A43, LSQR_B4, rcond=None)[0]
    m, n = 60, 40
    Abig = (2.0 * uniform01(71, 0, m * n) - 1.0).reshape(m, n)
    bbig = 2.0 * uniform01(72, 0, m) - 1.0
    big = float(np.sqrt(np.finfo(np.float64).max))
    return {
        "lsqr_trivial": (A0, np.array([1., 0., 0.]), dict()),
        "lsqr_consistent": (LSQR_A43, LSQR_A43 @ np.array([1., 2., 3.]), dict(max_iterations=1000, lam=0.0, btol=1e-6)),
        "lsqr_inconsistent": (LSQR_A43, LSQR_B4, dict(max_iterations=1000, lam=0.0, btol=0.0, Atol=1e-6)),
        "lsqr_trust_region": (LSQR_A43, LSQR_B4, dict(max_iterations=1000, lam=0.0, btol=0.0, Atol=0.0, cond_limit=1e12,
                                                      Delta=float(np.linalg.norm(xls)) / 2)),
        "lsqr_tikhonov": (LSQR_A43, LSQR_B4, dict(max_iterations=1000, lam=1.0, btol=0.0, Atol=1e-6)),
        "lsqr_big": (Abig, bbig, dict(max_iterations=500, lam=0.0, btol=1e-10, Atol=1e-10, cond_limit=1e12, Delta=big)),
        "lsqr_big_tikhonov_tr": (Abig, bbig, dict(max_iterations=500, lam=0.5, btol=1e-10, Atol=1e-10, cond_limit=1e12,
                                                  Delta=0.3)),
    }


def tnls_sine_cases(m: int = 100):
    """Curve-fitting problem of the reference's tests/TNLS_unit_test.cpp: y = sin(omega t + phi) on t = linspace(-pi, pi, m),
    omega = pi/2, phi = pi/4, start (1, 1); noiseless root finding, noisy least squares, noisy + preconditioner.
    name -> (t, y, kwargs)."""
    t = np.linspace(-np.pi, np.pi, m)
    y = np.sin(0.5 * np.pi * t + 0.25 * np.pi)
    z = 0.1 * (2.0 * uniform01(81, 0, m) - 1.0)
    fit = dict(rel_tol=0.0, grad_tol=1e-6, step_tol=0.0, Delta_tol=1e-10, root_tol=1e-6)
    return {
        "tnls_root": (t, y, dict(rel_tol=0.0, grad_tol=0.0, step_tol=0.0, Delta_tol=0.0, root_tol=1e-6)),
        "tnls_fit": (t, y + z, dict(fit)),
        "tnls_fit_precon": (t, y + z, dict(fit, use_precon=True)),
    }


def device_lsq_case(n: int = 5000):
    """Inputs of the device LSQR / TNLS checks (tests/host/lsq_device_check.cpp): a diagonal operator d in [0.5, 2],
    right-hand sides and a start, all from the counter-based generator.  LSQR: min |diag(d) x - b|;
    TNLS: F(x) = (d o x o x + x) - c  with c = F-consistent data of a known root x* in [0.2, 1.2]."""
    d = 0.5 + 1.5 * uniform01(91, 0, n)
    b = 2.0 * uniform01(92, 0, n) - 1.0
    xstar = 0.2 + uniform01(93, 0, n)
    c = (d * (xstar * xstar) + xstar)
    x0 = np.full(n, 0.5)
    return dict(d=d, b=b, c=c, x0=x0, xstar=xstar,
                lsqr=dict(max_iterations=200, lam=0.0, btol=1e-12, Atol=1e-12, cond_limit=1e12),
                tnls=dict(max_iterations=50, root_tol=1e-9, grad_tol=0.0, rel_tol=0.0, step_tol=0.0, Delta_tol=0.0))


def stiefel_row_scaling(n: int, p: int = 32) -> np.ndarray:
    """Elementwise scaling minv (n x p, constant along rows) of the projected Jacobi preconditioner used by the
    TNT + preconditioner checks: 1 / (1 + ((37 r) mod 64) / 128)."""
    w = 1.0 / (1.0 + ((37 * np.arange(n)) % 64) / 128.0)
    return np.ascontiguousarray(np.repeat(w[:, None], p, axis=1))


def laplacian3d_apply(X: np.ndarray, gx: int, gy: int, gz: int) -> np.ndarray:
    """7-point Laplacian with Dirichlet boundary on a gx x gy x gz grid (x fastest), applied to the rows of the
    block vector X (m x k, m = gx gy gz): (A X)[i] = 6 X[i] - sum of the existing neighbours (config C4 operator)."""
    k = X.shape[1]
    V = X.reshape(gz, gy, gx, k)
    out = 6.0 * V
    out[:, :, 1:] -= V[:, :, :-1]
    out[:, :, :-1] -= V[:, :, 1:]
    out[:, 1:] -= V[:, :-1]
    out[:, :-1] -= V[:, 1:]
    out[1:] -= V[:-1]
    out[:-1] -= V[1:]
    return out.reshape(X.shape)


@dataclasses.dataclass
class DiagProblem:
    """Diagonal SPD Hessian with optional Jacobi preconditioner (the shape of the
    reference's own STPCG tests, tests/IterativeSolvers_unit_test.cpp:82-131)."""
    n: int
    g: np.ndarray
    h: np.ndarray
    minv: np.ndarray


def make_diag(n: int, seed: int = 5, lo: float = 1000.0, hi: float = 3000.0) -> DiagProblem:
    g = 2.0 * uniform01(seed, 0, n) - 1.0
    h = lo + (hi - lo) * uniform01(seed + 1, 0, n)
    m = lo + (hi - lo) * uniform01(seed + 2, 0, n)
    return DiagProblem(n, g, h, 1.0 / m)


# ---- BASELINE config C5: rotation synchronisation on SO(3)^N (SE-Sync's relaxation) ----------------------------
@dataclasses.dataclass
class PoseGraphProblem:
    """min tr(X^T Q X) over X in St(3, r)^N (row-major 3N x r, pose i = rows 3i..3i+2, X_i X_i^T = I_3);
    Q = connection Laplacian of the rotation graph in block-CSR with 3 x 3 blocks (symmetric)."""
    N: int
    r: int
    rowptr: np.ndarray      # (N+1,) uint64
    colidx: np.ndarray      # (nnz,) uint32, sorted within a row
    blocks: np.ndarray      # (nnz, 9) float64, row-major 3 x 3
    X0: np.ndarray          # (3N, r)
    g: np.ndarray           # (3N, r) tangent at X0 (right-hand side of stand-alone tCG runs)

    @property
    def nnz(self) -> int:
        return int(self.colidx.size)


def _rot_from_axis_angle(w: np.ndarray) -> np.ndarray:
    """Rodrigues: (m, 3) rotation vectors -> (m, 3, 3) rotation matrices."""
    th = np.linalg.norm(w, axis=1)
    k = w / np.maximum(th, 1e-300)[:, None]
    K = np.zeros((w.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * K + (1.0 - c) * (K @ K)


def _tangent_project_poses(X: np.ndarray, Z: np.ndarray, r: int) -> np.ndarray:
    """Proj_X(Z)_i = Z_i - sym(Z_i X_i^T) X_i for every pose."""
    N = X.shape[0] // 3
    Xb, Zb = X.reshape(N, 3, r), Z.reshape(N, 3, r)
    M = Zb @ Xb.transpose(0, 2, 1)
    S = 0.5 * (M + M.transpose(0, 2, 1))
    return (Zb - S @ Xb).reshape(3 * N, r)


def make_posegraph(dims=(10, 10, 10), r: int = 4, seed: int = 41, sigma: float = 0.05, x_noise: float = 0.1,
                   closures: int = 1) -> PoseGraphProblem:
    """Synthetic pose graph: poses on a gx x gy x gz grid (x fastest) with odometry edges to the
    +x, +y, +z neighbours and `closures` random loop closures per pose; relative rotations measured with isotropic
    noise of `sigma` radians, unit weights.  sigma = 0 gives consistent measurements: the ground truth is then an exact
    minimiser (f = 0, Lambda = 0) and the Hessian is positive semidefinite -- the stand-alone tCG throughput workload
    (x_noise = 0 to start there).  X0 = ground truth perturbed by rotations of x_noise radians, embedded in r columns."""
    gx, gy, gz = dims
    N = gx * gy * gz
    idx = np.arange(N, dtype=np.int64)
    x, y, z = idx % gx, (idx // gx) % gy, idx // (gx * gy)
    ei, ej = [], []
    for cond, step in ((x + 1 < gx, 1), (y + 1 < gy, gx), (z + 1 < gz, gx * gy)):
        ei.append(idx[cond]); ej.append(idx[cond] + step)
    for c in range(closures):
        tgt = (splitmix64(seed + 3 + c, idx.astype(np.uint64)) % np.uint64(N)).astype(np.int64)
        keep = tgt != idx
        ei.append(idx[keep]); ej.append(tgt[keep])
    ei, ej = np.concatenate(ei), np.concatenate(ej)
    m = ei.size
    Rgt = _rot_from_axis_angle(2.0 * gaussish(seed, 0, 3 * N).reshape(N, 3))                 # ground-truth rotations
    noise = _rot_from_axis_angle(sigma * gaussish(seed + 1, 0, 3 * m).reshape(m, 3)) if sigma > 0 else np.eye(3)[None]
    Rij = Rgt[ei].transpose(0, 2, 1) @ Rgt[ej] @ noise                                      # measured R_i^T R_j
    # COO assembly: Q_ii += I, Q_jj += I, Q_ij -= R_ij, Q_ji -= R_ij^T
    rows = np.concatenate([ei, ej, ei, ej])
    cols = np.concatenate([ei, ej, ej, ei])
    eye = np.broadcast_to(np.eye(3), (m, 3, 3))
    vals = np.concatenate([eye, eye, -Rij, -Rij.transpose(0, 2, 1)]).reshape(-1, 9)
    key = rows * N + cols
    order = np.argsort(key, kind="stable")
    key, vals = key[order], vals[order]
    first = np.concatenate([[True], key[1:] != key[:-1]])
    starts = np.flatnonzero(first)
    blocks = np.add.reduceat(vals, starts, axis=0)
    ukey = key[starts]
    urow, ucol = ukey // N, ukey % N
    rowptr = np.zeros(N + 1, dtype=np.uint64)
    np.add.at(rowptr, urow + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.uint64)
    # X0: rows of (R_gt * Exp(x_noise))^T in the first three columns, small fill in the others, rows re-orthonormalised
    Rp = Rgt @ _rot_from_axis_angle(x_noise * gaussish(seed + 2, 0, 3 * N).reshape(N, 3)) if x_noise > 0 else Rgt
    Xb = np.zeros((N, 3, r))
    Xb[:, :, :3] = Rp.transpose(0, 2, 1)
    if r > 3 and x_noise > 0:
        Xb[:, :, 3:] = 0.1 * x_noise * gaussish(seed + 4, 0, 3 * N * (r - 3)).reshape(N, 3, r - 3)
    for a in range(3):                                                                       # Gram-Schmidt on the 3 rows
        for b in range(a):
            Xb[:, a] -= np.sum(Xb[:, a] * Xb[:, b], axis=1)[:, None] * Xb[:, b]
        Xb[:, a] /= np.linalg.norm(Xb[:, a], axis=1)[:, None]
    X0 = np.ascontiguousarray(Xb.reshape(3 * N, r))
    g = _tangent_project_poses(X0, gaussish(seed + 5, 0, 3 * N * r).reshape(3 * N, r), r)
    return PoseGraphProblem(N, r, rowptr, ucol.astype(np.uint32), np.ascontiguousarray(blocks), X0,
                            np.ascontiguousarray(g))


def posegraph_hess_numpy(prob: PoseGraphProblem, X: np.ndarray, V: np.ndarray):
    """Dense-free numpy Hess f(X)[V] = Proj_X(2 Q V - Lambda V), Lambda_i = sym(G_i X_i^T), G = 2 Q X (small N only).
    Returns (HV, Lambda (N, 3, 3), f, grad)."""
    import scipy.sparse as sp
    N, r = prob.N, prob.r
    Q = sp.bsr_matrix((prob.blocks.reshape(-1, 3, 3), prob.colidx.astype(np.int64), prob.rowptr.astype(np.int64)),
                      shape=(3 * N, 3 * N)).tocsr()
    G = 2.0 * (Q @ X)
    M = G.reshape(N, 3, r) @ X.reshape(N, 3, r).transpose(0, 2, 1)
    L = 0.5 * (M + M.transpose(0, 2, 1))
    f = 0.5 * float(np.sum(X * G))
    grad = G - (L @ X.reshape(N, 3, r)).reshape(3 * N, r)
    W = 2.0 * (Q @ V) - (L @ V.reshape(N, 3, r)).reshape(3 * N, r)
    return _tangent_project_poses(X, W, r), L, f, grad
