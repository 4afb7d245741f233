"""Generates tests/golden/*.json|npz from oracle/_ref (the UNMODIFIED reference
headers compiled against oracle::HostMat, see oracle/ref_driver.cpp).  Run in the
dev container only (needs /root/reference to rebuild oracle/_ref):

    make -C oracle && python tests/golden/make_golden.py

Single thread, so the floating-point summation order is fixed."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refapi import RefOracle, default_tnt_params  # noqa: E402
from optimization_b200 import problems as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DBL_MAX = 1.7976931348623157e308


def main():
    R = RefOracle()
    R.set_threads(1)
    out = {}
    # --- reference's S^2 problem (tests/TNT_unit_test.cpp:63-187) -------------
    x0, Ppt = [-.5, -.5, -.707107], [0., 0., 1.]
    for name, pre in (("s2_tnt", False), ("s2_tnt_precon", True)):
        for tag, prm in (("", default_tnt_params()),
                         ("_tight", default_tnt_params(gradient_tolerance=1e-8,
                                                       preconditioned_gradient_tolerance=1e-8,
                                                       relative_decrease_tolerance=1e-12,
                                                       stepsize_tolerance=1e-12))):
            r = R.s2_tnt(x0, Ppt, pre, prm)
            r["x"] = r["x"].tolist()
            r["params"] = prm
            out[name + tag] = r
    # --- STPCG known-answer tests (tests/IterativeSolvers_unit_test.cpp:138-251)
    g = np.array([21., -.4, 19.])
    H = np.array([1000., 100., 1.])
    M = np.array([100., 10., 1.])
    kats = {
        "ExactSTPCG": dict(g=g, h=H, minv=None, Delta=DBL_MAX, max_iterations=3, kappa_fgr=1e-8, theta=.999),
        "ExactSTPCGwithNegativeCurvature": dict(g=g, h=-H, minv=None, Delta=1000., max_iterations=3, kappa_fgr=1e-8, theta=.999),
        "ExactSTPCGwithPreconditioning": dict(g=g, h=H, minv=1. / M, Delta=DBL_MAX, max_iterations=3, kappa_fgr=1e-8, theta=.999),
        "ExactSTPCGwithNegativeCurvatureAndPreconditioning": dict(g=g, h=-H, minv=1. / M, Delta=1000., max_iterations=3, kappa_fgr=1e-8, theta=.999),
    }
    for k, a in kats.items():
        s, mn, it = R.stpcg_diag(**a)
        out[k] = dict(s=s.tolist(), update_step_M_norm=mn, num_iterations=it,
                      args={kk: (vv.tolist() if isinstance(vv, np.ndarray) else vv) for kk, vv in a.items()})
    # --- truncated cases on seeded data (shape of :254-310) ---------------------
    dp = P.make_diag(1000, seed=5)
    arrays = {}
    for name, minv, kw in (("diag1000_trunc", None, dict(Delta=1000., max_iterations=1000, kappa_fgr=.1, theta=.7)),
                           ("diag1000_precon_trunc", dp.minv, dict(Delta=1000., max_iterations=1000, kappa_fgr=.1, theta=.7)),
                           ("diag1000_tight", None, dict(Delta=1e6, max_iterations=1000, kappa_fgr=1e-9, theta=0.0)),
                           ("diag1000_boundary", None, dict(Delta=1e-4, max_iterations=1000, kappa_fgr=1e-9, theta=0.0))):
        s, mn, it = R.stpcg_diag(dp.g, dp.h, minv, **kw)
        out[name] = dict(update_step_M_norm=mn, num_iterations=it, args=kw, problem="make_diag(1000, seed=5)")
        arrays[name + "_s"] = s
    # --- Stiefel trace-min, small -------------------------------------------------
    for (n, yn) in ((512, 0.1), (1000, 0.1), (512, 0.3)):
        prob = P.make_stiefel(n, 32, y_noise=yn)
        rs = R.stiefel(prob)
        tag = f"stiefel{n}_yn{int(yn * 10)}"
        for name, kw in (("tight", dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)),
                         ("default", dict(Delta=1.0, max_iterations=1000, kappa_fgr=.1, theta=.5)),
                         ("boundary", dict(Delta=0.05, max_iterations=1000, kappa_fgr=1e-6, theta=.5))):
            s, mn, it = rs.stpcg(prob.Y0, prob.g, **kw)
            out[f"{tag}_{name}"] = dict(update_step_M_norm=mn, num_iterations=it, args=kw,
                                        problem=f"make_stiefel({n}, 32, y_noise={yn})")
            arrays[f"{tag}_{name}_s"] = s
        if n == 512:
            S, f, grad = rs.model(prob.Y0)
            arrays[f"{tag}_S"] = S
            arrays[f"{tag}_grad"] = grad
            arrays[f"{tag}_hess_g"] = rs.hess(prob.Y0, S, prob.g)
            out[f"{tag}_f"] = f
            r = rs.tnt(prob.Y0, default_tnt_params())
            arrays[f"{tag}_tnt_x"] = r.pop("x")
            r["params"] = default_tnt_params()
            out[f"{tag}_tnt"] = r
    for n in (512, 1000):
        prob = P.make_stiefel_critical(n, 32)
        rs = R.stiefel(prob)
        for name, kw in (("tight", dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)),
                         ("default", dict(Delta=1e3, max_iterations=1000, kappa_fgr=.1, theta=.5)),
                         ("boundary", dict(Delta=5.0, max_iterations=1000, kappa_fgr=1e-6, theta=.5))):
            s, mn, it = rs.stpcg(prob.Y0, prob.g, **kw)
            out[f"stiefelcrit{n}_{name}"] = dict(update_step_M_norm=mn, num_iterations=it, args=kw,
                                                 problem=f"make_stiefel_critical({n}, 32)")
            arrays[f"stiefelcrit{n}_{name}_s"] = s
    # --- sphere Rayleigh quotient, diag + low-rank A (configs C1 / C2 shape) ------------
    for (n, k) in ((1000, 16), (4099, 5)):
        prob = P.make_sphere_critical(n, k)
        gn = float(np.linalg.norm(prob.g))
        tag = f"spherecrit{n}_k{k}"
        for name, kw in (("tight", dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)),
                         ("default", dict(Delta=10.0 * gn, max_iterations=1000, kappa_fgr=.1, theta=.5)),
                         ("boundary", dict(Delta=0.2 * gn, max_iterations=1000, kappa_fgr=1e-6, theta=.5))):
            s, mn, it = R.sphere_stpcg(prob, prob.x0, prob.g, **kw)
            out[f"{tag}_{name}"] = dict(update_step_M_norm=mn, num_iterations=it, args=kw,
                                        problem=f"make_sphere_critical({n}, {k})")
            arrays[f"{tag}_{name}_s"] = s
    prob = P.make_sphere(100, 16)          # C1: n = 100, random start, default TNTParams
    r = R.sphere_tnt(prob, prob.x0, default_tnt_params())
    arrays["sphere100_tnt_x"] = r.pop("x")
    r["params"] = default_tnt_params()
    out["sphere100_tnt"] = r
    # --- constraint-preconditioned (projected) STPCG (tests/IterativeSolvers_unit_test.cpp:316-496 shape) ----
    h, m, A, g = P.make_projected(50, 3)
    for name, kw in (("projected_exact", dict(Delta=DBL_MAX, max_iterations=250, kappa_fgr=1e-8, theta=.7)),
                     ("projected_trunc", dict(Delta=1e-4, max_iterations=250, kappa_fgr=.1, theta=.7))):
        s, mn, it = R.stpcg_projected(h, m, A, g, **kw)
        out[name] = dict(update_step_M_norm=mn, num_iterations=it, args=kw, problem="make_projected(50, 3)")
        arrays[name + "_s"] = s
    # --- LSQR (reference IterativeSolvers.h:552-855; tests/IterativeSolvers_unit_test.cpp:517-700 cases) -------
    for name, (Am, bv, kw) in P.lsqr_cases().items():
        x, xn, it = R.lsqr(Am, bv, **kw)
        out[name] = dict(xnorm=xn, num_iterations=it, args=kw, x=x.tolist())
    # --- TNLS (reference Riemannian/TNLS.h; tests/TNLS_unit_test.cpp problem) ---------------------------------
    for name, (tt, yy, kw) in P.tnls_sine_cases().items():
        out[name] = dict(R.tnls_sine(tt, yy, [1.0, 1.0], **kw), args=kw)
    # --- GradientDescent (reference GradientDescent.h; tests/GradientDescent_unit_test.cpp shape) ----------
    r = R.s2_gd(x0, Ppt, max_iterations=1000, gradient_tolerance=1e-6)
    r["x"] = r["x"].tolist()
    out["s2_gd"] = r
    r = R.sphere_gd(prob, prob.x0, max_iterations=60, gradient_tolerance=1e-6)      # prob = make_sphere(100, 16)
    arrays["sphere100_gd_x"] = r.pop("x")
    out["sphere100_gd"] = r
    # --- TNT with a preconditioner (reference TNT.h:247, adapter l.413-426) on the device model shapes ----------
    prob = P.make_sphere(100, 16)
    r = R.sphere_tnt(prob, prob.x0, default_tnt_params(), minv=1.0 / (2.0 * prob.d))      # pointwise Jacobi
    arrays["sphere100_tnt_jacobi_x"] = r.pop("x")
    out["sphere100_tnt_jacobi"] = r
    prob = P.make_stiefel(512, 32, y_noise=.1)
    r = R.stiefel(prob).tnt(prob.Y0, default_tnt_params(), minv=P.stiefel_row_scaling(512, 32))   # P_Y(minv o V)
    arrays["stiefel512_yn1_tnt_pjacobi_x"] = r.pop("x")
    out["stiefel512_yn1_tnt_pjacobi"] = r
    # --- LSQR / TNLS on device-shaped (pointwise) operators: tests/host/lsq_device_check.cpp ---------------------
    case = P.device_lsq_case(5000)
    x, xn, it = R.lsqr_diag(case["d"], case["b"], **case["lsqr"])
    out["lsqr_diag5000"] = dict(xnorm=xn, num_iterations=it, args=case["lsqr"])
    arrays["lsqr_diag5000_x"] = x
    r = R.tnls_elem(case["d"], case["c"], case["x0"], **case["tnls"])
    arrays["tnls_elem5000_x"] = r.pop("x")
    r["args"] = case["tnls"]
    out["tnls_elem5000"] = r
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **arrays)
    print("wrote", len(out), "records,", len(arrays), "arrays")


if __name__ == "__main__":
    main()
