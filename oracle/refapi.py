"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for oracle/_ref/libref_oracle.so
(the unmodified reference headers instantiated with a stand-in host type, see
oracle/ref_driver.cpp) and oracle/liboracle_port.so (plain-C restatement)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libref_oracle.so")
PORT_SO = os.path.join(_HERE, "liboracle_port.so")

_dp = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)
_u16p = C.POINTER(C.c_uint16)

TNT_STATUS = ["Gradient", "PreconditionedGradient", "RelativeDecrease", "Stepsize",
              "TrustRegion", "IterationLimit", "ElapsedTime", "UserFunction"]


def _d(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def default_tnt_params(**kw):
    """TNTParams defaults (reference TNT.h:76-130, Concepts.h:42-60,116-131)."""
    p = dict(max_iterations=100, gradient_tolerance=1e-6, relative_decrease_tolerance=1e-6,
             stepsize_tolerance=1e-6, preconditioned_gradient_tolerance=1e-6,
             Delta_tolerance=1e-6, Delta0=1.0, eta1=0.05, eta2=0.9, alpha1=0.25, alpha2=2.5,
             max_TPCG_iterations=1000, kappa_fgr=0.1, theta=0.5,
             max_computation_time=1.7976931348623157e308)
    for k, v in kw.items():
        assert k in p, k
        p[k] = v
    return p


def _prm(p):
    keys = ["max_iterations", "gradient_tolerance", "relative_decrease_tolerance",
            "stepsize_tolerance", "preconditioned_gradient_tolerance", "Delta_tolerance",
            "Delta0", "eta1", "eta2", "alpha1", "alpha2", "max_TPCG_iterations", "kappa_fgr",
            "theta", "max_computation_time"]
    return np.array([float(p[k]) for k in keys], dtype=np.float64)


class _TraceBufs:
    def __init__(self, cap):
        self.cap = cap
        self.status = C.c_int(-1)
        self.n_outer = C.c_uint64(0)
        self.n_trace = C.c_uint64(0)
        self.scalars = np.zeros(4)
        self.inner = np.zeros(cap, dtype=np.uint64)
        self.radius = np.zeros(cap)
        self.rho = np.zeros(cap)
        self.fvals = np.zeros(cap)
        self.gradnorms = np.zeros(cap)
        self.step_norms = np.zeros(cap)
        self.step_M_norms = np.zeros(cap)

    def args(self):
        return (C.byref(self.status), C.byref(self.n_outer), C.byref(self.n_trace),
                _d(self.scalars), C.c_uint64(self.cap), self.inner.ctypes.data_as(_u64p),
                _d(self.radius), _d(self.rho), _d(self.fvals), _d(self.gradnorms),
                _d(self.step_norms), _d(self.step_M_norms))

    def result(self, x):
        no, nt = int(self.n_outer.value), int(self.n_trace.value)
        return dict(x=x, status=TNT_STATUS[self.status.value], status_code=self.status.value,
                    f=float(self.scalars[0]), gradfx_norm=float(self.scalars[1]),
                    preconditioned_grad_f_x_norm=float(self.scalars[2]),
                    elapsed_time=float(self.scalars[3]),
                    inner_iterations=[int(v) for v in self.inner[:no]],
                    gain_ratios=self.rho[:no].tolist(),
                    update_step_norms=self.step_norms[:no].tolist(),
                    update_step_M_norms=self.step_M_norms[:no].tolist(),
                    trust_region_radius=self.radius[:nt].tolist(),
                    objective_values=self.fvals[:nt].tolist(),
                    gradient_norms=self.gradnorms[:nt].tolist())


class RefOracle:
    """The reference's own STPCG / TNT / GradientDescent, compiled from
    /root/reference/include (oracle/Makefile target `ref`)."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` in the dev container)")
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_stiefel_create.restype = C.c_void_p
        L.ref_stiefel_create.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _u16p]
        L.ref_stiefel_destroy.argtypes = [C.c_void_p]
        L.ref_stiefel_model.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.ref_stiefel_hess.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.ref_stiefel_retract.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.ref_stiefel_stpcg.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_uint64,
                                        C.c_double, C.c_double, C.c_double, _dp, _dp, _u64p]
        L.ref_stiefel_stpcg_precon.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_int, C.c_double, C.c_uint64,
                                               C.c_double, C.c_double, C.c_double, _dp, _dp, _u64p]
        L.ref_stpcg_diag.argtypes = [C.c_uint64, _dp, _dp, _dp, C.c_double, C.c_uint64,
                                     C.c_double, C.c_double, C.c_double, _dp, _dp, _u64p]
        L.ref_sphere_stpcg.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, _dp,
                                       C.c_double, C.c_uint64, C.c_double, C.c_double,
                                       C.c_double, _dp, _dp, _u64p]

    # -- threads -----------------------------------------------------------
    def set_threads(self, t):
        self.lib.ref_set_threads(int(t))

    def max_threads(self):
        return int(self.lib.ref_max_threads())

    # -- STPCG, diagonal -----------------------------------------------------
    def stpcg_diag(self, g, h, minv=None, Delta=1.0, max_iterations=1000, kappa_fgr=0.1,
                   theta=0.5, epsilon=1e-8):
        n = g.size
        s = np.zeros(n)
        mn = C.c_double(0)
        it = C.c_uint64(0)
        rc = self.lib.ref_stpcg_diag(n, _d(g), _d(h), _d(minv), Delta, max_iterations,
                                     kappa_fgr, theta, epsilon, _d(s), C.byref(mn), C.byref(it))
        if rc:
            raise ValueError("std::invalid_argument from reference STPCG")
        return s, float(mn.value), int(it.value)

    # -- Stiefel -------------------------------------------------------------
    def stiefel(self, prob):
        return RefStiefel(self, prob)

    def sphere_stpcg(self, prob, x, g, Delta=1.0, max_iterations=1000, kappa_fgr=0.1, theta=0.5,
                     epsilon=1e-8):
        s = np.zeros(prob.n)
        mn = C.c_double(0)
        it = C.c_uint64(0)
        rc = self.lib.ref_sphere_stpcg(prob.n, prob.k, _d(prob.d), _d(prob.U), _d(prob.sigma),
                                       _d(x), _d(g), Delta, max_iterations, kappa_fgr, theta,
                                       epsilon, _d(s), C.byref(mn), C.byref(it))
        if rc:
            raise ValueError("std::invalid_argument from reference STPCG")
        return s, float(mn.value), int(it.value)

    def csr3_stpcg(self, prob, X, lam, g, minv=None, Delta=1.0, max_iterations=1000, kappa_fgr=0.1, theta=0.5,
                   epsilon=1e-8):
        """The reference's STPCG on the rotation-synchronisation Hessian (operator restated in oracle/sparse_ops.h)."""
        L = self.lib
        L.ref_csr3_stpcg.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_double,
                                     C.c_uint64, C.c_double, C.c_double, C.c_double, _dp, _dp, _u64p]
        s = np.zeros_like(g)
        mn, it = C.c_double(0), C.c_uint64(0)
        rc = L.ref_csr3_stpcg(prob.N, prob.r, prob.rowptr.ctypes.data, prob.colidx.ctypes.data, _d(prob.blocks), _d(lam),
                              _d(X), _d(g), _d(minv), Delta, max_iterations, kappa_fgr, theta, epsilon, _d(s),
                              C.byref(mn), C.byref(it))
        if rc:
            raise ValueError("std::invalid_argument from reference STPCG")
        return s, float(mn.value), int(it.value)

    def stencil7_stpcg(self, dims, p, g, minv=None, Delta=1.0, max_iterations=1000, kappa_fgr=0.1, theta=0.5,
                       epsilon=1e-8):
        L = self.lib
        L.ref_stencil7_stpcg.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, _dp, _dp, C.c_double, C.c_uint64,
                                         C.c_double, C.c_double, C.c_double, _dp, _dp, _u64p]
        s = np.zeros_like(g)
        mn, it = C.c_double(0), C.c_uint64(0)
        rc = L.ref_stencil7_stpcg(dims[0], dims[1], dims[2], p, _d(g), _d(minv), Delta, max_iterations, kappa_fgr, theta,
                                  epsilon, _d(s), C.byref(mn), C.byref(it))
        if rc:
            raise ValueError("std::invalid_argument from reference STPCG")
        return s, float(mn.value), int(it.value)

    def sphere_tnt(self, prob, x0, params=None, cap=2048, minv=None):
        """minv: pointwise Jacobi preconditioner precon(x, v) = minv o v (None: no preconditioner)."""
        p = params or default_tnt_params()
        tb = _TraceBufs(cap)
        x = np.zeros(prob.n)
        mv = None if minv is None else np.ascontiguousarray(minv, dtype=np.float64)
        rc = self.lib.ref_sphere_tnt_precon(C.c_uint64(prob.n), C.c_uint64(prob.k), _d(prob.d),
                                            _d(prob.U), _d(prob.sigma), _d(x0), _d(mv), _d(_prm(p)), _d(x),
                                            *tb.args())
        if rc:
            raise ValueError("std::invalid_argument from reference TNT")
        return tb.result(x)

    def s2_tnt(self, x0, P, use_precon=False, params=None, cap=2048):
        p = params or default_tnt_params()
        tb = _TraceBufs(cap)
        x = np.zeros(3)
        rc = self.lib.ref_s2_tnt(_d(np.ascontiguousarray(x0, dtype=np.float64)),
                                 _d(np.ascontiguousarray(P, dtype=np.float64)),
                                 C.c_int(1 if use_precon else 0), _d(_prm(p)), _d(x), *tb.args())
        if rc:
            raise ValueError("std::invalid_argument from reference TNT")
        return tb.result(x)

    def s2_gd(self, x0, P, max_iterations=1000, gradient_tolerance=1e-6):
        x = np.zeros(3)
        st = C.c_int(-1)
        it = C.c_uint64(0)
        f = C.c_double(0)
        gn = C.c_double(0)
        rc = self.lib.ref_s2_gd(_d(np.ascontiguousarray(x0, dtype=np.float64)),
                                _d(np.ascontiguousarray(P, dtype=np.float64)),
                                C.c_uint64(max_iterations), C.c_double(gradient_tolerance),
                                _d(x), C.byref(st), C.byref(it), C.byref(f), C.byref(gn))
        if rc:
            raise ValueError("std::invalid_argument from reference GradientDescent")
        return dict(x=x, status_code=st.value, iterations=int(it.value), f=f.value,
                    gradfx_norm=gn.value)


    def stpcg_projected(self, h, m, A, g, Delta, max_iterations, kappa_fgr, theta):
        """Reference STPCG with constraint preconditioning (P = KKT solve, At = A^T); A: mc x n."""
        mc, n = A.shape
        s = np.zeros(n)
        mn, it = C.c_double(0), C.c_uint64(0)
        self.lib.ref_stpcg_projected.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, C.c_double, C.c_uint64,
                                                 C.c_double, C.c_double, _dp, C.POINTER(C.c_double), _u64p]
        rc = self.lib.ref_stpcg_projected(n, mc, _d(h), _d(m), _d(np.ascontiguousarray(A)), _d(g), Delta,
                                          max_iterations, kappa_fgr, theta, _d(s), C.byref(mn), C.byref(it))
        if rc:
            raise ValueError("reference projected STPCG failed")
        return s, float(mn.value), int(it.value)

    def lsqr(self, A, b, max_iterations=1000, lam=0.0, btol=1e-6, Atol=1e-6, cond_limit=1e8, Delta=None):
        """Reference LSQR on a dense matrix."""
        m, n = A.shape
        if Delta is None:
            Delta = float(np.sqrt(np.finfo(np.float64).max))
        x = np.zeros(n)
        xn, it = C.c_double(0), C.c_uint64(0)
        self.lib.ref_lsqr.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, C.c_uint64, C.c_double, C.c_double, C.c_double,
                                      C.c_double, C.c_double, _dp, C.POINTER(C.c_double), _u64p]
        rc = self.lib.ref_lsqr(m, n, _d(np.ascontiguousarray(A, dtype=np.float64)), _d(b), max_iterations, lam, btol,
                               Atol, cond_limit, Delta, _d(x), C.byref(xn), C.byref(it))
        if rc:
            raise ValueError("std::invalid_argument from reference LSQR")
        return x, float(xn.value), int(it.value)

    def tnls_sine(self, t, y, beta0, use_precon=False, max_iterations=100, root_tol=1e-6, grad_tol=1e-6, rel_tol=1e-6,
                  step_tol=1e-6, Delta_tol=1e-6, cap=256):
        """Reference EuclideanTNLS on the sine-fit problem of tests/TNLS_unit_test.cpp."""
        beta = np.zeros(2)
        st, n = C.c_int(-1), C.c_uint64(0)
        inner = np.zeros(cap, dtype=np.uint64)
        rho, rad, fv = np.zeros(cap), np.zeros(cap), np.zeros(cap)
        f, gn = C.c_double(0), C.c_double(0)
        dd = C.c_double
        self.lib.ref_tnls_sine.argtypes = [C.c_uint64, _dp, _dp, _dp, C.c_int, C.c_uint64, dd, dd, dd, dd, dd, _dp,
                                           C.POINTER(C.c_int), _u64p, C.c_uint64, _u64p, _dp, _dp, _dp,
                                           C.POINTER(dd), C.POINTER(dd)]
        rc = self.lib.ref_tnls_sine(t.size, _d(t), _d(y), _d(np.ascontiguousarray(beta0, dtype=np.float64)),
                                    1 if use_precon else 0, max_iterations, root_tol, grad_tol, rel_tol, step_tol,
                                    Delta_tol, _d(beta), C.byref(st), C.byref(n), cap, inner.ctypes.data_as(_u64p),
                                    _d(rho), _d(rad), _d(fv), C.byref(f), C.byref(gn))
        if rc:
            raise ValueError("std::invalid_argument from reference TNLS")
        k = int(n.value)
        nrec = min(cap, k + 1)
        return dict(x=beta.tolist(), status_code=st.value, inner_iterations=inner[:k].astype(int).tolist(),
                    rho=rho[:k].tolist(), trust_region_radius=rad[:nrec].tolist(), objective_values=fv[:nrec].tolist(),
                    f=f.value, gradfx_norm=gn.value)

    def lsqr_diag(self, d, b, max_iterations=1000, lam=0.0, btol=1e-6, Atol=1e-6, cond_limit=1e8, Delta=None):
        """Reference LSQR on A = diag(d) (the device check's operator shape)."""
        n = d.size
        x = np.zeros(n)
        xn, it = C.c_double(0), C.c_uint64(0)
        dd = C.c_double
        self.lib.ref_lsqr_diag.argtypes = [C.c_uint64, _dp, _dp, C.c_uint64, dd, dd, dd, dd, dd, _dp, C.POINTER(dd), _u64p]
        rc = self.lib.ref_lsqr_diag(n, _d(d), _d(b), max_iterations, lam, btol, Atol, cond_limit,
                                    np.finfo(np.float64).max if Delta is None else Delta, _d(x), C.byref(xn), C.byref(it))
        if rc:
            raise ValueError("std::invalid_argument from reference LSQR")
        return x, float(xn.value), int(it.value)

    def tnls_elem(self, d, b, x0, max_iterations=100, root_tol=1e-6, grad_tol=1e-6, rel_tol=1e-6, step_tol=1e-6,
                  Delta_tol=1e-6, cap=256):
        """Reference EuclideanTNLS on F(x) = (d o x o x + x) - b (separable; every operation a device level-1 kernel)."""
        n = d.size
        x = np.zeros(n)
        st, no = C.c_int(-1), C.c_uint64(0)
        inner = np.zeros(cap, dtype=np.uint64)
        rho, rad, fv = np.zeros(cap), np.zeros(cap), np.zeros(cap)
        f, gn = C.c_double(0), C.c_double(0)
        dd = C.c_double
        self.lib.ref_tnls_elem.argtypes = [C.c_uint64, _dp, _dp, _dp, C.c_uint64, dd, dd, dd, dd, dd, _dp,
                                           C.POINTER(C.c_int), _u64p, C.c_uint64, _u64p, _dp, _dp, _dp,
                                           C.POINTER(dd), C.POINTER(dd)]
        rc = self.lib.ref_tnls_elem(n, _d(d), _d(b), _d(np.ascontiguousarray(x0, dtype=np.float64)), max_iterations,
                                    root_tol, grad_tol, rel_tol, step_tol, Delta_tol, _d(x), C.byref(st), C.byref(no), cap,
                                    inner.ctypes.data_as(_u64p), _d(rho), _d(rad), _d(fv), C.byref(f), C.byref(gn))
        if rc:
            raise ValueError("std::invalid_argument from reference TNLS")
        k = int(no.value)
        nrec = min(cap, k + 1)
        return dict(x=x, status_code=st.value, inner_iterations=inner[:k].astype(int).tolist(), rho=rho[:k].tolist(),
                    trust_region_radius=rad[:nrec].tolist(), objective_values=fv[:nrec].tolist(), f=f.value,
                    gradfx_norm=gn.value)

    def sphere_gd(self, prob, x0, max_iterations=100, gradient_tolerance=1e-6):
        x = np.zeros(prob.n)
        st, it, ls = C.c_int(-1), C.c_uint64(0), C.c_uint64(0)
        f, gn = C.c_double(0), C.c_double(0)
        self.lib.ref_sphere_gd.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, C.c_uint64, C.c_double,
                                           _dp, C.POINTER(C.c_int), _u64p, _u64p, C.POINTER(C.c_double),
                                           C.POINTER(C.c_double)]
        rc = self.lib.ref_sphere_gd(prob.n, prob.k, _d(prob.d), _d(prob.U), _d(prob.sigma), _d(x0),
                                    max_iterations, gradient_tolerance, _d(x), C.byref(st), C.byref(it),
                                    C.byref(ls), C.byref(f), C.byref(gn))
        if rc:
            raise ValueError("std::invalid_argument from reference GradientDescent")
        return dict(x=x, status_code=st.value, iterations=int(it.value), linesearch_total=int(ls.value),
                    f=f.value, gradfx_norm=gn.value)


class RefStiefel:
    def __init__(self, ora, prob):
        self.ora, self.prob = ora, prob
        A = np.ascontiguousarray(prob.A_bf16, dtype=np.uint16)
        self.h = ora.lib.ref_stiefel_create(prob.n, prob.p, prob.nb, A.ctypes.data_as(_u16p))

    def __del__(self):
        try:
            self.ora.lib.ref_stiefel_destroy(self.h)
        except Exception:
            pass

    def model(self, Y):
        p = self.prob.p
        S = np.zeros((p, p))
        f = C.c_double(0)
        grad = np.zeros_like(Y)
        self.ora.lib.ref_stiefel_model(self.h, _d(Y), _d(S), C.byref(f), _d(grad))
        return S, float(f.value), grad

    def hess(self, Y, S, V):
        out = np.zeros_like(V)
        self.ora.lib.ref_stiefel_hess(self.h, _d(Y), _d(S), _d(V), _d(out))
        return out

    def retract(self, Y, V):
        out = np.zeros_like(V)
        self.ora.lib.ref_stiefel_retract(self.h, _d(Y), _d(V), _d(out))
        return out

    def stpcg(self, Y, g, minv=None, Delta=1.0, max_iterations=1000, kappa_fgr=0.1, theta=0.5,
              epsilon=1e-8, projected=False):
        """projected: the preconditioner is P_Y(minv o r) (tangent-space preserving) instead of minv o r."""
        s = np.zeros_like(g)
        mn = C.c_double(0)
        it = C.c_uint64(0)
        rc = self.ora.lib.ref_stiefel_stpcg_precon(self.h, _d(Y), _d(g), _d(minv), int(bool(projected)), Delta,
                                                   max_iterations, kappa_fgr, theta, epsilon, _d(s), C.byref(mn),
                                                   C.byref(it))
        if rc:
            raise ValueError("std::invalid_argument from reference STPCG")
        return s, float(mn.value), int(it.value)

    def tnt(self, Y0, params=None, cap=2048, minv=None):
        """minv: elementwise scaling of the projected Jacobi preconditioner precon(Y, V) = P_Y(minv o V) (None: no
        preconditioner)."""
        p = params or default_tnt_params()
        tb = _TraceBufs(cap)
        Y = np.zeros_like(Y0)
        mv = None if minv is None else np.ascontiguousarray(minv, dtype=np.float64)
        rc = self.ora.lib.ref_stiefel_tnt_precon(C.c_void_p(self.h), _d(Y0), _d(mv),
                                                 _d(_prm(p)), _d(Y), *tb.args())
        if rc:
            raise ValueError("std::invalid_argument from reference TNT")
        return tb.result(Y)


class PortOracle:
    """oracle/liboracle_port.so: the plain-C restatement (oracle/stpcg_port.c)."""

    EXIT = {0: "residual", 1: "max_iterations", 2: "kernel", 3: "boundary", -1: "bad_argument"}

    def __init__(self, path=PORT_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle port`)")
        self.lib = C.CDLL(path)
        L = self.lib
        L.port_dot.restype = C.c_double
        L.port_dot.argtypes = [_dp, _dp, C.c_uint64]
        L.port_stpcg_diag.argtypes = [C.c_uint64, _dp, _dp, _dp, C.c_double, C.c_uint64,
                                      C.c_double, C.c_double, C.c_double, _dp, _dp, _u64p]
        L.port_stpcg_stiefel.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp,
                                         C.c_double, C.c_uint64, C.c_double, C.c_double,
                                         C.c_double, _dp, _dp, _u64p]
        L.port_stiefel_S.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, _dp, _dp, _dp]
        L.port_stpcg_sphere.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, _dp, _dp,
                                        C.c_double, C.c_uint64, C.c_double, C.c_double,
                                        C.c_double, _dp, _dp, _u64p]
        L.port_sphere_model.restype = C.c_double
        L.port_sphere_model.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, _dp, _dp]
        L.port_sphere_hess.argtypes = [C.c_uint64, C.c_uint64, _dp, _dp, _dp, _dp, _dp, _dp]

    def dot(self, x, y):
        return float(self.lib.port_dot(_d(x), _d(y), x.size))

    def stpcg_diag(self, g, h, minv=None, Delta=1.0, max_iterations=1000, kappa_fgr=0.1,
                   theta=0.5, epsilon=1e-8):
        s = np.zeros(g.size)
        mn = C.c_double(0)
        it = C.c_uint64(0)
        rc = self.lib.port_stpcg_diag(g.size, _d(g), _d(h), _d(minv), Delta, max_iterations,
                                      kappa_fgr, theta, epsilon, _d(s), C.byref(mn), C.byref(it))
        if rc < 0:
            raise ValueError("invalid argument")
        return s, float(mn.value), int(it.value), self.EXIT[rc]

    def stpcg_sphere(self, prob, x, g, minv=None, Delta=1.0, max_iterations=1000, kappa_fgr=0.1,
                     theta=0.5, epsilon=1e-8):
        s = np.zeros(prob.n)
        mn = C.c_double(0)
        it = C.c_uint64(0)
        rc = self.lib.port_stpcg_sphere(prob.n, prob.k, _d(prob.d), _d(prob.U), _d(prob.sigma), _d(x),
                                        _d(g), _d(minv), Delta, max_iterations, kappa_fgr, theta,
                                        epsilon, _d(s), C.byref(mn), C.byref(it))
        if rc < 0:
            raise ValueError("invalid argument")
        return s, float(mn.value), int(it.value), self.EXIT[rc]

    def sphere_model(self, prob, x):
        Ax = np.zeros(prob.n)
        grad = np.zeros(prob.n)
        f = self.lib.port_sphere_model(prob.n, prob.k, _d(prob.d), _d(prob.U), _d(prob.sigma), _d(x),
                                       _d(Ax), _d(grad))
        return float(f), Ax, grad

    def sphere_hess(self, prob, x, v):
        out = np.zeros(prob.n)
        self.lib.port_sphere_hess(prob.n, prob.k, _d(prob.d), _d(prob.U), _d(prob.sigma), _d(x), _d(v),
                                  _d(out))
        return out

    def stpcg_stiefel(self, prob, Y, g, minv=None, Delta=1.0, max_iterations=1000, kappa_fgr=0.1,
                      theta=0.5, epsilon=1e-8):
        from optimization_b200.problems import from_bf16_bits
        A = np.ascontiguousarray(from_bf16_bits(prob.A_bf16))
        s = np.zeros_like(g)
        mn = C.c_double(0)
        it = C.c_uint64(0)
        rc = self.lib.port_stpcg_stiefel(prob.n, prob.p, prob.nb, _d(A), _d(Y), _d(g), _d(minv),
                                         Delta, max_iterations, kappa_fgr, theta, epsilon, _d(s),
                                         C.byref(mn), C.byref(it))
        if rc < 0:
            raise ValueError("invalid argument")
        return s, float(mn.value), int(it.value), self.EXIT[rc]

    # -- sparse Hessian families (configs C5 / C4) ------------------------------------------------------
    def csr3_model(self, prob, X):
        L = self.lib
        L.port_csr3_model.restype = C.c_double
        L.port_csr3_model.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, _dp, _dp, _dp, _dp]
        lam = np.zeros((prob.N, 9))
        grad = np.zeros_like(X)
        f = L.port_csr3_model(prob.N, prob.r, prob.rowptr.ctypes.data, prob.colidx.ctypes.data, _d(prob.blocks), _d(X),
                              _d(lam), _d(grad))
        return lam, float(f), grad

    def csr3_hess(self, prob, X, lam, v):
        L = self.lib
        L.port_csr3_hess.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, _dp, _dp, _dp, _dp, _dp]
        out = np.zeros_like(v)
        L.port_csr3_hess(prob.N, prob.r, prob.rowptr.ctypes.data, prob.colidx.ctypes.data, _d(prob.blocks), _d(lam),
                         _d(X), _d(v), _d(out))
        return out

    def stpcg_csr3(self, prob, X, lam, g, minv=None, Delta=1.0, max_iterations=1000, kappa_fgr=0.1, theta=0.5,
                   epsilon=1e-8):
        L = self.lib
        L.port_stpcg_csr3.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, _dp, _dp, _dp, _dp, _dp, C.c_double,
                                      C.c_uint64, C.c_double, C.c_double, C.c_double, _dp, _dp, _u64p]
        s = np.zeros_like(g)
        mn, it = C.c_double(0), C.c_uint64(0)
        rc = L.port_stpcg_csr3(prob.N, prob.r, prob.rowptr.ctypes.data, prob.colidx.ctypes.data, _d(prob.blocks), _d(lam),
                               _d(X), _d(g), _d(minv), Delta, max_iterations, kappa_fgr, theta, epsilon, _d(s),
                               C.byref(mn), C.byref(it))
        if rc < 0:
            raise ValueError("invalid argument")
        return s, float(mn.value), int(it.value), self.EXIT[rc]

    def stencil7_apply(self, dims, p, v):
        L = self.lib
        L.port_stencil7_apply.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, _dp, _dp]
        out = np.zeros_like(v)
        L.port_stencil7_apply(dims[0], dims[1], dims[2], p, _d(v), _d(out))
        return out

    def stpcg_stencil7(self, dims, p, g, minv=None, Delta=1.0, max_iterations=1000, kappa_fgr=0.1, theta=0.5,
                       epsilon=1e-8):
        L = self.lib
        L.port_stpcg_stencil7.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, _dp, _dp, C.c_double, C.c_uint64,
                                          C.c_double, C.c_double, C.c_double, _dp, _dp, _u64p]
        s = np.zeros_like(g)
        mn, it = C.c_double(0), C.c_uint64(0)
        rc = L.port_stpcg_stencil7(dims[0], dims[1], dims[2], p, _d(g), _d(minv), Delta, max_iterations, kappa_fgr, theta,
                                   epsilon, _d(s), C.byref(mn), C.byref(it))
        if rc < 0:
            raise ValueError("invalid argument")
        return s, float(mn.value), int(it.value), self.EXIT[rc]

    def s2_tnt(self, x0, P, use_precon=False, params=None, cap=2048):
        p = params or default_tnt_params()
        tb = _TraceBufs(cap)
        x = np.zeros(3)
        rc = self.lib.port_s2_tnt(_d(np.ascontiguousarray(x0, dtype=np.float64)),
                                  _d(np.ascontiguousarray(P, dtype=np.float64)),
                                  C.c_int(1 if use_precon else 0), _d(_prm(p)), _d(x), *tb.args())
        if rc:
            raise ValueError("invalid argument")
        return tb.result(x)


class RefLobpcg:
    """oracle/_ref/libref_lobpcg.so: the reference's own LOBPCG.h compiled against oracle/eigen_shim (see
    oracle/ref_lobpcg_driver.cpp).  Operators are descriptors: None | ("diag", array) | ("scalar", alpha) |
    ("laplacian", (gx, gy, gz))."""

    def __init__(self, path=os.path.join(_HERE, "_ref", "libref_lobpcg.so")):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` in the dev container)")
        self.lib = C.CDLL(path)
        self.lib.ref_lobpcg.argtypes = [C.c_int, _dp, C.c_double, C.POINTER(C.c_uint32), C.c_int, _dp, C.c_double, C.c_int,
                                        _dp, C.c_double, C.c_uint64, C.c_uint64, _dp, C.c_uint64, C.c_uint64, C.c_double,
                                        _dp, _dp, _u64p, _u64p, _dp]
        self.lib.ref_lobpcg_omega.argtypes = [C.c_uint64, C.c_uint64, _dp]

    @staticmethod
    def _op(op):
        if op is None:
            return 0, None, 0.0, None, None
        kind, val = op
        if kind == "diag":
            a = np.ascontiguousarray(val, dtype=np.float64)
            return 1, _d(a), 0.0, None, a
        if kind == "scalar":
            return 2, None, float(val), None, None
        g = (C.c_uint32 * 3)(*val)
        return 3, None, 0.0, g, g

    def omega(self, m, nx):
        out = np.zeros((m, nx))
        self.lib.ref_lobpcg_omega(m, nx, _d(out))
        return out

    def lobpcg(self, A, B, T, X0, nev, max_iters, tau=1e-6, trace=False):
        m, nx = X0.shape
        ka, da, aa, ga, keep_a = self._op(A)
        kb, db, ab, _, keep_b = self._op(B)
        kt, dt, at, _, keep_t = self._op(T)
        theta, X = np.zeros(nev), np.zeros((m, nev))
        it, nc = C.c_uint64(0), C.c_uint64(0)
        tr = np.zeros((max_iters, nx)) if trace else None
        rc = self.lib.ref_lobpcg(ka, da, aa, ga, kb, db, ab, kt, dt, at, m, nx, _d(np.ascontiguousarray(X0)), nev,
                                 max_iters, tau, _d(theta), _d(X), C.byref(it), C.byref(nc), _d(tr))
        if rc == 1:
            raise ValueError("std::invalid_argument from reference LOBPCG")
        if rc:
            raise RuntimeError("reference LOBPCG failed")
        out = (theta, X, int(it.value), int(nc.value))
        return out + (tr[:int(it.value) + 1],) if trace else out
