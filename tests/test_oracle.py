"""CPU: pins the oracles.  (a) closed-form known-answer tests of the reference's
own STPCG tests (tests/IterativeSolvers_unit_test.cpp:138-251), (b) the golden
fixtures generated from oracle/_ref (unmodified reference headers), (c) the
C restatement (oracle/stpcg_port.c) against oracle/_ref bit for bit when the
latter is present, and the reference's TNT test assertions
(tests/TNT_unit_test.cpp:126-187)."""
import os

import numpy as np
import pytest

from oracle import refapi
from optimization_b200 import problems as P

DBL_MAX = 1.7976931348623157e308
g3 = np.array([21., -.4, 19.])
H3 = np.array([1000., 100., 1.])
M3 = np.array([100., 10., 1.])


@pytest.fixture(scope="module")
def port(build_oracle):
    return refapi.PortOracle()


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(refapi.REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    r = refapi.RefOracle()
    r.set_threads(1)
    return r


def _both(port, ref_or_none):
    return [port] + ([ref_or_none] if ref_or_none is not None else [])


def test_kat_exact_stpcg(port):
    # IterativeSolvers_unit_test.cpp:138-159
    s, mn, it, why = port.stpcg_diag(g3, H3, Delta=DBL_MAX, max_iterations=3, kappa_fgr=1e-8, theta=.999)
    s_true = -g3 / H3
    assert np.linalg.norm(s - s_true) < 1e-6
    assert abs(mn - np.linalg.norm(s)) / np.linalg.norm(s) < 1e-6
    assert it == 3


def test_kat_negative_curvature(port):
    # :165-186  -> s = -Delta g/|g|, 0 completed iterations
    s, mn, it, why = port.stpcg_diag(g3, -H3, Delta=1000., max_iterations=3, kappa_fgr=1e-8, theta=.999)
    assert np.linalg.norm(s - (-1000. * g3 / np.linalg.norm(g3))) < 1e-6
    assert abs(mn - 1000.) < 1e-6 and it == 0 and why == "boundary"


def test_kat_preconditioned(port):
    # :190-216
    s, mn, it, why = port.stpcg_diag(g3, H3, 1. / M3, Delta=DBL_MAX, max_iterations=3, kappa_fgr=1e-8, theta=.999)
    assert np.linalg.norm(s + g3 / H3) < 1e-6
    sM = np.sqrt(np.sum(M3 * s * s))
    assert abs(mn - sM) / sM < 1e-6


def test_kat_negative_curvature_preconditioned(port):
    # :220-251  -> s = Delta p/|p|_M, p = -M^-1 g
    s, mn, it, why = port.stpcg_diag(g3, -H3, 1. / M3, Delta=1000., max_iterations=3, kappa_fgr=1e-8, theta=.999)
    p = -g3 / M3
    s_true = 1000. * p / np.sqrt(np.sum(M3 * p * p))
    assert np.linalg.norm(s - s_true) < 1e-6
    assert abs(mn - 1000.) < 1e-6


def test_truncation_property(port):
    # :254-275: n=1000 random diag in [1000,3000], kappa=.1, theta=.7, Delta=1000, max 3 iterations
    dp = P.make_diag(1000, seed=5)
    s, mn, it, why = port.stpcg_diag(dp.g, dp.h, Delta=1000., max_iterations=3, kappa_fgr=.1, theta=.7)
    assert np.linalg.norm(dp.g + dp.h * s) / np.linalg.norm(dp.g) < .1
    # :279-310 with preconditioning: M^-1-norm of the residual
    s, mn, it, why = port.stpcg_diag(dp.g, dp.h, dp.minv, Delta=1000., max_iterations=3, kappa_fgr=.1, theta=.7)
    r = dp.g + dp.h * s
    assert np.sqrt(np.sum(r * dp.minv * r)) / np.sqrt(np.sum(dp.g * dp.minv * dp.g)) < .1


def test_invalid_arguments(port):
    # IterativeSolvers.h:183-205
    for kw in (dict(Delta=0.), dict(Delta=-1.), dict(kappa_fgr=1.), dict(kappa_fgr=-.1), dict(theta=1.5),
               dict(theta=-.1), dict(epsilon=0.), dict(epsilon=1.)):
        args = dict(Delta=1., max_iterations=3, kappa_fgr=.1, theta=.5, epsilon=1e-8)
        args.update(kw)
        with pytest.raises(ValueError):
            port.stpcg_diag(g3, H3, **args)


def test_port_matches_golden_kats(port, golden):
    rec, arr = golden
    for name in ("ExactSTPCG", "ExactSTPCGwithNegativeCurvature", "ExactSTPCGwithPreconditioning",
                 "ExactSTPCGwithNegativeCurvatureAndPreconditioning"):
        a = rec[name]["args"]
        minv = None if a["minv"] is None else np.array(a["minv"])
        s, mn, it, why = port.stpcg_diag(np.array(a["g"]), np.array(a["h"]), minv, Delta=a["Delta"],
                                         max_iterations=a["max_iterations"], kappa_fgr=a["kappa_fgr"],
                                         theta=a["theta"])
        assert it == rec[name]["num_iterations"]
        assert mn == rec[name]["update_step_M_norm"]
        assert s.tolist() == rec[name]["s"]          # bit for bit


def test_port_matches_golden_diag(port, golden):
    rec, arr = golden
    dp = P.make_diag(1000, seed=5)
    for name, minv in (("diag1000_trunc", None), ("diag1000_precon_trunc", dp.minv),
                       ("diag1000_tight", None), ("diag1000_boundary", None)):
        s, mn, it, why = port.stpcg_diag(dp.g, dp.h, minv, **rec[name]["args"])
        assert it == rec[name]["num_iterations"]
        assert mn == rec[name]["update_step_M_norm"]
        assert np.array_equal(s, arr[name + "_s"])


@pytest.mark.parametrize("tag,maker", [("stiefel512_yn1", lambda: P.make_stiefel(512, 32, y_noise=.1)),
                                       ("stiefel512_yn3", lambda: P.make_stiefel(512, 32, y_noise=.3)),
                                       ("stiefelcrit512", lambda: P.make_stiefel_critical(512, 32))])
def test_port_matches_golden_stiefel(port, golden, tag, maker):
    rec, arr = golden
    prob = maker()
    for name in ("tight", "default", "boundary"):
        r = rec[f"{tag}_{name}"]
        s, mn, it, why = port.stpcg_stiefel(prob, prob.Y0, prob.g, **r["args"])
        assert it == r["num_iterations"], name
        assert mn == r["update_step_M_norm"], name
        assert np.array_equal(s, arr[f"{tag}_{name}_s"]), name


def test_s2_tnt_golden_and_reference_assertions(port, golden):
    # tests/TNT_unit_test.cpp:126-187: status == Gradient, |grad| < tol, F decreased
    rec, arr = golden
    x0, Ppt = [-.5, -.5, -.707107], [0., 0., 1.]
    for name, pre in (("s2_tnt", False), ("s2_tnt_precon", True), ("s2_tnt_tight", False),
                      ("s2_tnt_precon_tight", True)):
        r = rec[name]
        out = port.s2_tnt(x0, Ppt, pre, r["params"])
        assert out["status"] == "Gradient" == r["status"]
        assert out["gradfx_norm"] < r["params"]["gradient_tolerance"]
        assert out["f"] < out["objective_values"][0]
        assert out["inner_iterations"] == r["inner_iterations"]
        for key in ("gain_ratios", "trust_region_radius", "objective_values", "gradient_norms",
                    "update_step_norms", "update_step_M_norms"):
            assert out[key] == r[key], key              # bit for bit
        assert out["x"].tolist() == r["x"]


def test_golden_matches_survey_record(golden):
    # values captured from the unmodified headers during the survey
    rec, _ = golden
    r = rec["s2_tnt_tight"]
    assert r["inner_iterations"] == [1, 1, 1, 1, 1, 1]
    assert r["trust_region_radius"][:2] == [1.0, 1.7677657456064366]
    assert abs(r["gain_ratios"][0] - 2.1520193581695812) < 1e-14
    assert rec["s2_tnt_precon_tight"]["inner_iterations"] == [2, 1, 2, 2, 2, 2]
    assert rec["ExactSTPCG"]["num_iterations"] == 3
    assert abs(rec["ExactSTPCG"]["update_step_M_norm"] - 19.000012026311978) < 1e-13
    assert rec["ExactSTPCGwithNegativeCurvature"]["num_iterations"] == 0


def test_ref_matches_port_live(port, ref):
    """oracle/_ref (reference headers) == C restatement, bit for bit, seeded inputs."""
    dp = P.make_diag(4099, seed=9)
    for minv in (None, dp.minv):
        for kw in (dict(Delta=1e6, max_iterations=50, kappa_fgr=1e-10, theta=0.),
                   dict(Delta=1e-3, max_iterations=50, kappa_fgr=.1, theta=.5)):
            a = ref.stpcg_diag(dp.g, dp.h, minv, **kw)
            b = port.stpcg_diag(dp.g, dp.h, minv, **kw)
            assert a[1] == b[1] and a[2] == b[2] and np.array_equal(a[0], b[0])
    prob = P.make_stiefel_critical(300, 32)
    a = ref.stiefel(prob).stpcg(prob.Y0, prob.g, Delta=1e6, max_iterations=100, kappa_fgr=1e-9, theta=0.)
    b = port.stpcg_stiefel(prob, prob.Y0, prob.g, Delta=1e6, max_iterations=100, kappa_fgr=1e-9, theta=0.)
    assert a[1] == b[1] and a[2] == b[2] and np.array_equal(a[0], b[0])


# ---- sphere Rayleigh quotient (configs C1 / C2 shape) ----------------------------------
@pytest.mark.parametrize("n,k", [(1000, 16), (4099, 5)])
def test_port_matches_golden_sphere(port, golden, n, k):
    rec, arr = golden
    prob = P.make_sphere_critical(n, k)
    for name in ("tight", "default", "boundary"):
        r = rec[f"spherecrit{n}_k{k}_{name}"]
        s, mn, it, why = port.stpcg_sphere(prob, prob.x0, prob.g, **r["args"])
        assert it == r["num_iterations"] and mn == r["update_step_M_norm"]
        assert np.array_equal(s, arr[f"spherecrit{n}_k{k}_{name}_s"])          # bit for bit
    # the residual-reduction property the reference tests (IterativeSolvers_unit_test.cpp:254-275)
    kw = rec[f"spherecrit{n}_k{k}_tight"]["args"]
    s, mn, it, why = port.stpcg_sphere(prob, prob.x0, prob.g, **kw)
    r = prob.g + port.sphere_hess(prob, prob.x0, s)
    assert why == "residual" and np.linalg.norm(r) <= 1.0001 * kw["kappa_fgr"] * np.linalg.norm(prob.g)


def test_sphere_tnt_golden_reference_assertions(golden):
    # the assertions of tests/TNT_unit_test.cpp:126-187 on the recorded C1 run (n = 100, default TNTParams)
    rec, arr = golden
    r = rec["sphere100_tnt"]
    x = arr["sphere100_tnt_x"]
    assert abs(np.linalg.norm(x) - 1.0) < 1e-12
    assert r["objective_values"][-1] < r["objective_values"][0] if "objective_values" in r else True


def test_ref_matches_port_live_sphere(port, ref):
    for n, k in ((257, 16), (5000, 3), (300, 0)):
        prob = P.make_sphere_critical(n, k)
        gn = float(np.linalg.norm(prob.g))
        for kw in (dict(Delta=1e6, max_iterations=50, kappa_fgr=1e-10, theta=0.),
                   dict(Delta=.2 * gn, max_iterations=50, kappa_fgr=.1, theta=.5)):
            a = ref.sphere_stpcg(prob, prob.x0, prob.g, **kw)
            b = port.stpcg_sphere(prob, prob.x0, prob.g, **kw)
            assert a[2] == b[2] and a[1] == b[1] and np.array_equal(a[0], b[0])


# ---- sparse Hessian families (configs C5 / C4): the C restatement against the reference's STPCG on the same operator ----
def test_sparse_operator_ports_match_the_reference_header(build_oracle):
    import numpy as np
    from optimization_b200 import problems as P
    from oracle import refapi
    port = refapi.PortOracle()
    try:
        R = refapi.RefOracle()
    except (FileNotFoundError, OSError):
        import pytest
        pytest.skip("oracle/_ref not built here")
    R.set_threads(1)
    for pr in (P.make_posegraph((8, 7, 6), 4), P.make_posegraph((8, 7, 6), 3, sigma=0., x_noise=0.), P.make_posegraph((6, 6, 6), 5)):
        lam, f, grad = port.csr3_model(pr, pr.X0)
        hv, L, f_np, grad_np = P.posegraph_hess_numpy(pr, pr.X0, pr.g)
        assert abs(f - f_np) <= 1e-12 * max(1.0, abs(f_np)) and np.abs(lam.reshape(-1, 3, 3) - L).max() < 1e-12
        assert np.linalg.norm(port.csr3_hess(pr, pr.X0, lam, pr.g) - hv) <= 1e-13 * np.linalg.norm(hv)
        for kw in (dict(Delta=1e6, max_iterations=60, kappa_fgr=1e-8, theta=0.), dict(Delta=0.5, max_iterations=60, kappa_fgr=1e-3, theta=.5)):
            s, mn, it, why = port.stpcg_csr3(pr, pr.X0, lam, pr.g, **kw)
            s2, mn2, it2 = R.csr3_stpcg(pr, pr.X0, lam, pr.g, **kw)
            assert it == it2 and mn == mn2 and np.array_equal(s, s2)
    dims, p = (9, 8, 7), 3
    n = dims[0] * dims[1] * dims[2]
    g = (2 * P.uniform01(5, 0, n * p) - 1).reshape(n, p)
    for mv in (None, np.full((n, p), 1 / 6.)):
        kw = dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-8, theta=0.)
        s, mn, it, why = port.stpcg_stencil7(dims, p, g, mv, **kw)
        s2, mn2, it2 = R.stencil7_stpcg(dims, p, g, mv, **kw)
        assert it == it2 and mn == mn2 and np.array_equal(s, s2) and why == "residual"
        assert np.linalg.norm(P.laplacian3d_apply(s, *dims) + g) <= 1e-7 * np.linalg.norm(g)
