"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy + scipy.linalg.eigh) of the reference's LOBPCG
(include/Optimization/LinearAlgebra/LOBPCG.h).  The reference header needs Eigen, which is not in this
image, so this path has no compiled reference oracle: parity is pinned by the known-answer
tests of the reference's own tests/LOBPCG_unit_test.cpp (tests/test_oracle_lobpcg.py) and by analytic spectra.

Each statement cites the reference line it follows.  Block vectors are m x k numpy arrays; operators are callables
on such arrays.  `Omega` (the random probe block of l.172-181) is passed in explicitly so that the CUDA path and
this restatement use the same norm estimates; the reference draws it from std::default_random_engine.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg


def rayleigh_ritz(A: np.ndarray, B: np.ndarray):
    """LOBPCG.h:53-62: generalized symmetric eigenproblem A c = theta B c after diagonal equilibration of B;
    eigenvalues ascending, C^T B C = I."""
    D = 1.0 / np.sqrt(np.diag(B))                                       # l.56
    As = (D[:, None] * A) * D[None, :]
    Bs = (D[:, None] * B) * D[None, :]
    theta, V = scipy.linalg.eigh(0.5 * (As + As.T), 0.5 * (Bs + Bs.T))  # l.58-59 (self-adjoint solver: symmetric part)
    return theta, D[:, None] * V                                        # l.61


def lobpcg(A, B, T, X0: np.ndarray, nev: int, max_iters: int, tau: float = 1e-6, Omega: np.ndarray | None = None,
           user_function=None):
    """LOBPCG.h:131-337.  Returns (Theta[:nev], X[:, :nev], num_iters, nc)."""
    m, nx = X0.shape                                                     # l.142-143
    if nev > nx:
        raise ValueError("Block size nx must be greater than or equal to the number nev of desired eigenpairs")  # l.149
    if nx > m:
        raise ValueError("Block size nx must be less than or equal to the dimension m of the problem")           # l.153
    X = X0.copy()                                                        # l.158
    if Omega is None:
        Omega = np.random.default_rng(0).standard_normal((m, nx))        # l.172-176
    A2normest = np.linalg.norm(A(Omega)) / np.linalg.norm(Omega)        # l.180 (Frobenius norms)
    B2normest = np.linalg.norm(B(Omega)) / np.linalg.norm(Omega) if B is not None else 1.0   # l.181
    AX = A(X)                                                            # l.186
    BX = B(X) if B is not None else X                                    # l.187
    Theta, C = rayleigh_ritz(X.T @ AX, X.T @ BX)                         # l.190-191
    AX = AX @ C                                                          # l.194
    BX = BX @ C                                                          # l.195
    R = AX - BX * Theta[None, :]                                         # l.198
    nc = 0                                                               # l.201
    P = None
    num_iters = 1
    S = np.zeros((m, 3 * nx))                                            # l.166
    while num_iters < max_iters:                                         # l.204
        W = T(R) if T is not None else R                                 # l.207
        S[:, :nx] = X                                                    # l.210
        S[:, nx:2 * nx - nc] = W[:, nc:]                                 # l.213
        if num_iters > 1:
            S[:, 2 * nx - nc:3 * nx - 2 * nc] = P[:, nc:]               # l.217
            ns = 3 * nx - 2 * nc                                         # l.218
        else:
            ns = 2 * nx - nc                                             # l.220
        Sa = S[:, :ns]
        AS = A(Sa)                                                       # l.225
        BS = B(Sa) if B is not None else Sa                              # l.226
        Theta, C = rayleigh_ritz(Sa.T @ AS, Sa.T @ BS)                   # l.229-233
        X = Sa @ C[:, :nx]                                               # l.239
        AX = A(X)                                                        # l.242
        BX = B(X) if B is not None else X                                # l.243
        R = AX - BX * Theta[None, :nx]                                   # l.246
        P = Sa[:, nx:] @ C[nx:ns, :nx]                                   # l.249
        r = np.linalg.norm(R, axis=0)                                    # l.254
        tol = tau * (A2normest + B2normest * np.abs(Theta[:nx])) * np.linalg.norm(X, axis=0)   # l.257-261
        conv = r[:nev] <= tol[:nev]                                      # l.263-264
        nc = 0
        while nc < nev and conv[nc]:                                     # l.267-269
            nc += 1
        if user_function is not None and user_function(num_iters, nev, Theta[:nx], X, r, nc):   # l.272-274
            break
        if nc == nev:                                                    # l.277
            break
        num_iters += 1
    return Theta[:nev].copy(), X[:, :nev].copy(), num_iters, nc          # l.332-336
