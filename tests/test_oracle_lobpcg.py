"""CPU: pins the LOBPCG restatement (oracle/lobpcg_port.py) with the known-answer tests of the reference's own
tests/LOBPCG_unit_test.cpp (no compiled reference exists for this path: LOBPCG.h needs Eigen) and with the analytic
spectrum of the 3-D Dirichlet Laplacian (the operator of BASELINE config C4)."""
import numpy as np

from oracle import lobpcg_port as L
from optimization_b200 import problems as P

N, M, NEV, TAU = 1000, 10, 5, 1e-8
ADIAG = np.linspace(-.5 * N, .5 * N, N)
BDIAG = np.linspace(1.0, N, N)


def _A(X): return ADIAG[:, None] * X
def _B(X): return BDIAG[:, None] * X
def _T(X): return np.abs(ADIAG)[:, None] * X        # the reference's test preconditioner (LOBPCG_unit_test.cpp:58-61)


def _x0(m, nx, seed=91):
    return (2.0 * P.uniform01(seed, 0, m * nx) - 1.0).reshape(m, nx)       # Matrix::Random is uniform on [-1, 1]


def test_rayleigh_ritz_properties():
    # LOBPCG_unit_test.cpp:66-91
    rng = np.random.default_rng(3)
    AL, BL = rng.uniform(-1, 1, (7, 7)), rng.uniform(-1, 1, (7, 7))
    A, B = -AL @ AL.T, BL @ BL.T
    theta, C = L.rayleigh_ritz(A, B)
    assert np.linalg.norm(C.T @ A @ C - np.diag(theta)) < 1e-8
    assert np.linalg.norm(C.T @ B @ C - np.eye(7)) < 1e-8
    assert np.all(np.diff(theta) >= 0)


def test_small_eigenvalue_problem():
    # :94-120, with the literal X0 of the reference test
    lam = np.array([1., 2., 3., 4.])
    X0 = np.array([[0.8147, 0.6324], [0.9058, 0.0975], [0.1270, 0.2785], [0.9134, 0.5469]])
    theta, X, it, nc = L.lobpcg(lambda X: lam[:, None] * X, None, None, X0, 2, N, TAU)
    assert nc == 2 and np.linalg.norm(theta - lam[:2]) < 1e-3


def test_eigenvalue_problem_unpreconditioned():
    theta, X, it, nc = L.lobpcg(_A, None, None, _x0(N, M), NEV, 10 * N, TAU)                      # :123-140
    assert nc == NEV and np.linalg.norm(theta - ADIAG[:NEV]) < 1e-4


def test_preconditioned_eigenvalue_problem():
    theta, X, it, nc = L.lobpcg(_A, None, _T, _x0(N, M), NEV, N, TAU)                             # :144-161
    assert nc == NEV and np.linalg.norm(theta - ADIAG[:NEV]) < 1e-4


def test_generalized_eigenvalue_problems():
    lam = np.sort(ADIAG / BDIAG)
    for T in (_T, None):                                                                       # :164-208
        theta, X, it, nc = L.lobpcg(_A, _B, T, _x0(N, M), NEV, N, TAU)
        assert nc == NEV and np.linalg.norm(theta - lam[:NEV]) < 1e-4


def test_laplacian_spectrum_with_jacobi():
    # config C4 shape at small size: 7-point Dirichlet Laplacian on a g^3 grid, T = diag^-1 = 1/6
    g = 10
    lam1 = 2.0 - 2.0 * np.cos(np.arange(1, g + 1) * np.pi / (g + 1))
    exact = np.sort((lam1[:, None, None] + lam1[None, :, None] + lam1[None, None, :]).ravel())
    A = lambda X: P.laplacian3d_apply(X, g, g, g)
    theta, X, it, nc = L.lobpcg(A, None, lambda R: R / 6.0, _x0(g ** 3, 8, seed=31), 4, 300, 1e-8)
    assert nc == 4 and np.allclose(theta, exact[:4], rtol=1e-6)


def test_numpy_restatement_is_pinned_to_the_reference_header():
    """oracle/lobpcg_port.py (numpy) against the reference's own LOBPCG.h compiled with the Eigen stand-in
    (oracle/_ref/libref_lobpcg.so): same probe block, iteration counts within one, eigenvalues to 1e-9."""
    import numpy as np
    import pytest
    from optimization_b200 import problems as P
    from oracle import lobpcg_port as L, refapi
    try:
        R = refapi.RefLobpcg()
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built here")
    n, nx, nev = 1000, 10, 5
    ad, bd = np.linspace(-.5 * n, .5 * n, n), np.linspace(1.0, n, n)
    X0 = (2.0 * P.uniform01(91, 0, n * nx) - 1.0).reshape(n, nx)
    Om = R.omega(n, nx)
    for gen, pre in ((False, False), (False, True), (True, True), (True, False)):
        th, X, it, nc = R.lobpcg(("diag", ad), ("diag", bd) if gen else None, ("diag", np.abs(ad)) if pre else None, X0, nev,
                                 10 * n, 1e-8)
        th2, _, it2, nc2 = L.lobpcg(lambda V: ad[:, None] * V, (lambda V: bd[:, None] * V) if gen else None,
                                    (lambda V: np.abs(ad)[:, None] * V) if pre else None, X0, nev, 10 * n, 1e-8, Omega=Om)
        assert nc == nc2 == nev and abs(it - it2) <= 1 and np.allclose(th, th2, rtol=1e-9, atol=1e-9)
        exact = np.sort(ad / bd)[:nev] if gen else ad[:nev]
        assert np.linalg.norm(th - exact) < 1e-4          # the bar of the reference's tests/LOBPCG_unit_test.cpp
