"""The C++ drop-in header layer (include/Optimization/...: same include paths, names and
semantics as the reference, written from scratch).
CPU: control flow of OUR TNT.h / IterativeSolvers.h on the reference's own test problems
     (tests/TNT_unit_test.cpp:63-187, tests/IterativeSolvers_unit_test.cpp:138-251) with a
     host vector type, compared bit for bit with the golden traces from the unmodified reference.
GPU: TNT<DeviceMatrix, DeviceMatrix, double> end to end on the Stiefel trace-min problem (fused
     device tCG inside) against the golden TNT run of the reference."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from optimization_b200 import problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")


def _compile(name, link):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host", name + ".cpp"), "-o", exe]
    if link:
        cmd += ["-L" + os.path.join(ROOT, "optimization_b200"), "-loptimization_b200",
                "-Wl,-rpath," + os.path.join(ROOT, "optimization_b200")]
    subprocess.run(cmd, check=True)
    return exe


def _lines(out):
    return {d["case"]: d for d in (json.loads(l) for l in out.splitlines() if l.startswith("{"))}


def test_host_headers_match_reference_golden(golden):
    rec, _ = golden
    exe = _compile("tnt_host_check", link=False)
    got = _lines(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    for name in ("s2_tnt", "s2_tnt_precon", "s2_tnt_tight", "s2_tnt_precon_tight"):
        g, r = got[name], rec[name]
        assert g["status_code"] == r["status_code"] == 0               # TNTStatus::Gradient
        assert g["inner_iterations"] == r["inner_iterations"]
        for key in ("gain_ratios", "trust_region_radius", "objective_values", "gradient_norms",
                    "update_step_M_norms", "x"):
            assert g[key] == r[key], (name, key)                        # bit for bit
        assert g["f"] == r["f"] and g["gradfx_norm"] == r["gradfx_norm"]
    for name in ("ExactSTPCG", "ExactSTPCGwithNegativeCurvature", "ExactSTPCGwithPreconditioning",
                 "ExactSTPCGwithNegativeCurvatureAndPreconditioning"):
        assert got[name]["s"] == rec[name]["s"]
        assert got[name]["num_iterations"] == rec[name]["num_iterations"]
        assert got[name]["update_step_M_norm"] == rec[name]["update_step_M_norm"]
    assert got["invalid_argument"]["thrown"] == 1


def test_gradient_descent_header_matches_reference_golden(golden):
    """OUR Riemannian/GradientDescent.h (+ EuclideanGradientDescent / EuclideanTNT with the reference's signatures)."""
    rec, _ = golden
    exe = _compile("gd_host_check", link=False)
    got = _lines(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)
    g, r = got["s2_gd"], rec["s2_gd"]
    assert g["status_code"] == r["status_code"] == 0 and g["iterations"] == r["iterations"]
    assert g["f"] == r["f"] and g["gradfx_norm"] == r["gradfx_norm"] and g["x"] == r["x"]      # bit for bit
    assert g["hook_calls"] == g["accepted"] == r["iterations"] - 1        # user function: once per accepted step
    assert got["gd_invalid_argument"]["thrown"] == 4                      # GradientDescent.h:141-161
    e = got["euclidean"]
    assert e["tnt_status"] == 0 and e["tnt_err"] < 1e-12 and e["gd_err"] < 1e-6


def test_projected_stpcg_matches_reference_golden(golden, tmp_path):
    """Constraint-preconditioned (projected) path of OUR IterativeSolvers.h (P + At, Multiplier type) and the
    per-iteration user hook; assertions of the reference's tests/IterativeSolvers_unit_test.cpp:316-496."""
    rec, arr = golden
    exe = _compile("projected_host_check", link=False)
    h, m, A, g = P.make_projected(50, 3)
    f = tmp_path / "proj.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<QQ", 50, 3))
        for a in (h, m, A, g):
            fh.write(np.ascontiguousarray(a).tobytes())
    got = _lines(subprocess.run([exe, str(f)], check=True, capture_output=True, text=True).stdout)
    for name in ("projected_exact", "projected_trunc"):
        gt, r = got[name], rec[name]
        assert gt["num_iterations"] == r["num_iterations"] and gt["update_step_M_norm"] == r["update_step_M_norm"]
        assert np.array_equal(np.array(gt["s"]), arr[name + "_s"])                      # bit for bit
        assert gt["As_norm"] < 1e-9                                                     # s in ker(A)  (:400)
        assert abs(gt["s_M_norm"] - gt["update_step_M_norm"]) <= 1e-6 * gt["s_M_norm"]  # M-norm recurrence (:408)
    # exact solve == primal part of the KKT solution [H A^T; A 0][s; l] = [-g; 0]  (:326-404)
    n, mc = 50, 3
    K = np.zeros((n + mc, n + mc))
    K[:n, :n] = np.diag(h)
    K[:n, n:] = A.T
    K[n:, :n] = A
    s_gt = np.linalg.solve(K, np.concatenate([-g, np.zeros(mc)]))[:n]
    s = np.array(got["projected_exact"]["s"])
    assert np.linalg.norm(s - s_gt) / np.linalg.norm(s_gt) < 1e-6
    assert got["projected_trunc"]["update_step_M_norm"] == 1e-4                         # truncated at the boundary
    assert got["user_hook"] == {"case": "user_hook", "calls": 3, "num_iterations": 2}  # hook stops the loop (l.365-369)


def test_lsqr_header_matches_reference_golden(golden, tmp_path):
    """OUR LSQR (LinearAlgebra/IterativeSolvers.h) on the five cases of the reference's LSQR unit tests
    (tests/IterativeSolvers_unit_test.cpp:517-700) + a seeded 60 x 40 system: bit for bit against the reference
    header, plus that test file's own assertions."""
    rec, _ = golden
    exe = _compile("lsqr_host_check", link=False)
    cases = P.lsqr_cases()
    big = float(np.sqrt(np.finfo(np.float64).max))
    f = tmp_path / "lsqr.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<Q", len(cases)))
        for name, (A, b, kw) in cases.items():
            m, n = A.shape
            fh.write(struct.pack("<QQ", m, n))
            fh.write(np.ascontiguousarray(A).tobytes())
            fh.write(np.ascontiguousarray(b).tobytes())
            fh.write(struct.pack("<Q5d", kw.get("max_iterations", 1000), kw.get("lam", 0.0), kw.get("btol", 1e-6),
                                 kw.get("Atol", 1e-6), kw.get("cond_limit", 1e8), kw.get("Delta", big)))
    out = subprocess.run([exe, str(f)], check=True, capture_output=True, text=True).stdout
    lines = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    got = {name: d for name, d in zip(cases, lines)}
    for name in cases:
        assert got[name]["num_iterations"] == rec[name]["num_iterations"], name
        assert got[name]["xnorm"] == rec[name]["xnorm"] and got[name]["x"] == rec[name]["x"], name     # bit for bit
    A, b = P.LSQR_A43, P.LSQR_B4
    # :517-557 x = 0 is stationary: immediate return
    assert got["lsqr_trivial"]["num_iterations"] == 0 and got["lsqr_trivial"]["xnorm"] == 0
    # :560-597 consistent system: relative residual, reported norm, iteration count
    x = np.array(got["lsqr_consistent"]["x"])
    bc = A @ np.array([1., 2., 3.])
    assert np.linalg.norm(A @ x - bc) < 1e-6 * np.linalg.norm(bc)
    assert abs(got["lsqr_consistent"]["xnorm"] - np.linalg.norm(x)) < 1e-6 * np.linalg.norm(x)
    assert got["lsqr_consistent"]["num_iterations"] < 12
    # :600-633 inconsistent system: least-squares solution
    xls = np.linalg.lstsq(A, b, rcond=None)[0]
    x = np.array(got["lsqr_inconsistent"]["x"])
    assert np.linalg.norm(x - xls) < 1e-5 * np.linalg.norm(xls)
    # :637-690 trust region binding: terminates on the boundary and still reduces the residual
    x = np.array(got["lsqr_trust_region"]["x"])
    Delta = np.linalg.norm(xls) / 2
    assert abs(got["lsqr_trust_region"]["xnorm"] - Delta) < 1e-9 and abs(np.linalg.norm(x) - Delta) < 1e-9
    assert np.linalg.norm(A @ x - b) < np.linalg.norm(b)
    # :693-735 Tikhonov: normal equations (A^T A + lambda I) x = A^T b
    x = np.array(got["lsqr_tikhonov"]["x"])
    xt = np.linalg.solve(A.T @ A + np.eye(3), A.T @ b)
    assert np.linalg.norm(x - xt) < 1e-5 * np.linalg.norm(xt)
    assert lines[-1] == {"case": "lsqr_invalid_argument", "thrown": 5}                  # IterativeSolvers.h:568-587


def test_tnls_header_matches_reference_golden(golden, tmp_path):
    """OUR Riemannian/TNLS.h (EuclideanTNLS -> TNLS -> LSQR) on the curve-fitting problem of the reference's
    tests/TNLS_unit_test.cpp (root finding, noisy fit, noisy fit with a right preconditioner)."""
    rec, _ = golden
    exe = _compile("tnls_host_check", link=False)
    cases = P.tnls_sine_cases()
    f = tmp_path / "tnls.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<Q", len(cases)))
        for name, (t, y, kw) in cases.items():
            fh.write(struct.pack("<Q", t.size))
            fh.write(np.ascontiguousarray(t).tobytes())
            fh.write(np.ascontiguousarray(y).tobytes())
            fh.write(struct.pack("<QQ5d", 1 if kw.get("use_precon") else 0, 100, kw["root_tol"], kw["grad_tol"],
                                 kw["rel_tol"], kw["step_tol"], kw["Delta_tol"]))
    out = subprocess.run([exe, str(f)], check=True, capture_output=True, text=True).stdout
    lines = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    for (name, (t, y, kw)), g in zip(cases.items(), lines):
        r = rec[name]
        assert g["status_code"] == r["status_code"] and g["inner_iterations"] == r["inner_iterations"], name
        # floating-point traces: identical arithmetic up to libm's sin / cos (same image => bit for bit in practice)
        for key in ("rho", "trust_region_radius", "objective_values", "x"):
            assert np.allclose(g[key], r[key], rtol=1e-11, atol=1e-13), (name, key)
        assert abs(g["f"] - r["f"]) <= 1e-11 * abs(r["f"]) + 1e-15
    # assertions of the reference's tests (TNLS_unit_test.cpp:118-205)
    root, fit, fitp = lines
    t, y, _ = cases["tnls_root"]
    assert root["status_code"] == 0                                                    # TNLSStatus::Root
    assert np.linalg.norm(y - np.sin(root["x"][0] * t + root["x"][1])) < 1e-6
    t, yn, _ = cases["tnls_fit"]
    noise = np.linalg.norm(yn - y)
    for g in (fit, fitp):
        assert g["status_code"] == 1 and g["gradfx_norm"] < 1e-6                       # TNLSStatus::Gradient
        assert np.linalg.norm(yn - np.sin(g["x"][0] * t + g["x"][1])) < noise          # better than the planted signal


def test_header_layer_has_reference_layout():
    for rel in ("Optimization/Base/Concepts.h", "Optimization/Riemannian/Concepts.h",
                "Optimization/Riemannian/TNT.h", "Optimization/Riemannian/GradientDescent.h",
                "Optimization/Riemannian/TNLS.h", "Optimization/LinearAlgebra/Concepts.h",
                "Optimization/LinearAlgebra/IterativeSolvers.h", "Optimization/LinearAlgebra/LOBPCG.h",
                "Optimization/Util/Stopwatch.h",
                "Optimization/b200/Device.h", "optimization_b200.h"):
        assert os.path.exists(os.path.join(ROOT, "include", rel)), rel


@pytest.mark.gpu
def test_device_tnt_matches_reference_golden(golden, tmp_path):
    rec, arr = golden
    exe = _compile("tnt_device_check", link=True)
    prob = P.make_stiefel(512, 32, y_noise=.1)
    f = tmp_path / "prob.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<QQ", prob.n, prob.p))
        fh.write(np.ascontiguousarray(prob.A_bf16).tobytes())
        fh.write(np.ascontiguousarray(prob.Y0).tobytes())
        fh.write(np.ascontiguousarray(prob.g).tobytes())
    xo, so = tmp_path / "x.bin", tmp_path / "s.bin"
    out = subprocess.run([exe, str(f), str(xo), str(so)], check=True, capture_output=True, text=True).stdout
    got = _lines(out)
    r = rec["stiefel512_yn1_tnt"]
    g = got["tnt"]
    assert g["status_code"] == r["status_code"]                          # bit-exact termination status
    assert g["inner_iterations"] == r["inner_iterations"]                # bit-exact iteration counts
    # rho = (f - f_new) / predicted: the cancellation in f - f_new amplifies rounding by |f| / df (1e7 in the
    # last iteration), so rho is only defined to ~1e-7 there; decisions (accept / radius) still agree exactly
    assert np.allclose(g["gain_ratios"], r["gain_ratios"], rtol=1e-6, atol=0)
    assert np.allclose(g["trust_region_radius"], r["trust_region_radius"], rtol=1e-10, atol=0)
    assert np.allclose(g["objective_values"], r["objective_values"], rtol=1e-12, atol=0)
    x = np.fromfile(xo).reshape(prob.n, prob.p)
    x_ref = arr["stiefel512_yn1_tnt_x"]
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-10     # final iterate
    assert g["last_path"] in (0, 1, 2)                                   # fp64 MMA kernel, v6 or v4 tcgen05 kernel
    # direct STPCG on descriptor functors: fused path (a handful of launches) == generic loop
    s = got["stpcg"]
    assert s["num_iterations"] == s["generic_iterations"] == rec["stiefel512_yn1_tight"]["num_iterations"]
    assert s["rel_diff"] < 1e-10
    assert s["fused_launches"] <= 8 < s["generic_launches"]
    s_dev = np.fromfile(so).reshape(prob.n, prob.p)
    s_ref = arr["stiefel512_yn1_tight_s"]
    if np.all(np.isfinite(s_ref)):
        assert np.linalg.norm(s_dev - s_ref) / np.linalg.norm(s_ref) < 1e-10


@pytest.mark.gpu
def test_device_sphere_tnt_matches_reference_golden(golden, tmp_path):
    """Config C1 (sphere, n = 100, default TNTParams) through TNT<DeviceMatrix, DeviceMatrix, double> with the
    SphereRayleigh functor set: fused device tCG inside, model / retraction on the device."""
    rec, arr = golden
    exe = _compile("tnt_sphere_check", link=True)
    prob = P.make_sphere(100, 16)
    f = tmp_path / "sphere.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<QQ", prob.n, prob.k))
        for a in (prob.d, prob.U, prob.sigma, prob.x0):
            fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    xo, xg = tmp_path / "x.bin", tmp_path / "xgd.bin"
    out = subprocess.run([exe, str(f), str(xo), str(xg)], check=True, capture_output=True, text=True).stdout
    gd, rgd = _lines(out)["sphere_gd"], rec["sphere100_gd"]
    # GradientDescent<DeviceMatrix>: same status, iteration count and line-search trial count as the reference run
    assert (gd["status_code"], gd["iterations"], gd["linesearch_total"]) == \
        (rgd["status_code"], rgd["iterations"], rgd["linesearch_total"])
    assert abs(gd["f"] - rgd["f"]) <= 1e-12 * abs(rgd["f"])
    xgd = np.fromfile(xg)
    assert np.linalg.norm(xgd - arr["sphere100_gd_x"]) / np.linalg.norm(arr["sphere100_gd_x"]) < 1e-10
    g = _lines(out)["sphere_tnt"]
    r = rec["sphere100_tnt"]
    assert g["status_code"] == r["status_code"]                          # bit-exact termination status
    assert g["inner_iterations"] == r["inner_iterations"]                # bit-exact iteration counts
    assert np.allclose(g["trust_region_radius"], r["trust_region_radius"], rtol=1e-10, atol=0)
    assert np.allclose(g["objective_values"], r["objective_values"], rtol=1e-10, atol=1e-13)
    assert np.allclose(g["gain_ratios"], r["gain_ratios"], rtol=1e-6, atol=0)
    x = np.fromfile(xo)
    x_ref = arr["sphere100_tnt_x"]
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-10     # final iterate


@pytest.mark.gpu
def test_device_tnt_with_preconditioner_matches_reference_golden(golden, tmp_path):
    """TNT + preconditioner on device matrices (reference TNT.h:247, adapter l.413-426): sphere model with the pointwise
    Jacobi descriptor (stays on the fused tCG path) and Stiefel model with the tangent-space preserving projected Jacobi
    functor (generic loop over device kernels), both against runs of the unmodified reference headers."""
    rec, arr = golden
    exe = _compile("tnt_precon_check", link=True)
    sp = P.make_sphere(100, 16)
    st = P.make_stiefel(512, 32, y_noise=.1)
    f = tmp_path / "precon.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<QQ", sp.n, sp.k))
        for a in (sp.d, sp.U, sp.sigma, sp.x0, 1.0 / (2.0 * sp.d)):
            fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        fh.write(struct.pack("<QQ", st.n, st.p))
        fh.write(np.ascontiguousarray(st.A_bf16).tobytes())
        fh.write(np.ascontiguousarray(st.Y0).tobytes())
        fh.write(P.stiefel_row_scaling(st.n, st.p).tobytes())
    xs, xy = tmp_path / "xs.bin", tmp_path / "xy.bin"
    out = subprocess.run([exe, str(f), str(xs), str(xy)], check=True, capture_output=True, text=True).stdout
    got = _lines(out)
    for case, key, xfile, shape in (("sphere_tnt_jacobi", "sphere100_tnt_jacobi", xs, (sp.n,)),
                                    ("stiefel_tnt_pjacobi", "stiefel512_yn1_tnt_pjacobi", xy, (st.n, st.p))):
        g, r = got[case], rec[key]
        assert g["status_code"] == r["status_code"]                          # bit-exact termination status
        assert g["inner_iterations"] == r["inner_iterations"]                # bit-exact iteration counts
        assert np.allclose(g["trust_region_radius"], r["trust_region_radius"], rtol=1e-10, atol=0)
        assert np.allclose(g["objective_values"], r["objective_values"], rtol=1e-10, atol=1e-13)
        assert np.allclose(g["gain_ratios"], r["gain_ratios"], rtol=1e-6, atol=0)
        x = np.fromfile(xfile).reshape(shape)
        x_ref = arr[key + "_x"]
        assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-10     # final iterate
    # the Jacobi descriptor keeps the solves fused: a handful of launches per inner solve instead of ~10 per CG iteration
    n_outer = len(rec["sphere100_tnt_jacobi"]["inner_iterations"])
    assert got["sphere_tnt_jacobi"]["launches"] < 40 * n_outer


@pytest.mark.gpu
def test_device_lsqr_and_tnls_match_reference_golden(golden, tmp_path):
    """LSQR<DeviceMatrix> and EuclideanTNLS<DeviceMatrix> (reference IterativeSolvers.h:552-855, TNLS.h:265-729) on
    operators made of device level-1 kernels, against the unmodified reference headers on the same operators."""
    rec, arr = golden
    exe = _compile("lsq_device_check", link=True)
    case = P.device_lsq_case(5000)
    f = tmp_path / "lsq.bin"
    lk, tk = case["lsqr"], case["tnls"]
    with open(f, "wb") as fh:
        fh.write(struct.pack("<Q", case["d"].size))
        for a in (case["d"], case["b"], case["c"], case["x0"]):
            fh.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        fh.write(struct.pack("<Qdddd", lk["max_iterations"], lk["lam"], lk["btol"], lk["Atol"], lk["cond_limit"]))
        fh.write(struct.pack("<Qddddd", tk["max_iterations"], tk["root_tol"], tk["grad_tol"], tk["rel_tol"], tk["step_tol"],
                             tk["Delta_tol"]))
    xl, xt = tmp_path / "xl.bin", tmp_path / "xt.bin"
    out = subprocess.run([exe, str(f), str(xl), str(xt)], check=True, capture_output=True, text=True).stdout
    got = _lines(out)
    g, r = got["lsqr_diag"], rec["lsqr_diag5000"]
    assert g["num_iterations"] == r["num_iterations"]
    assert abs(g["xnorm"] - r["xnorm"]) <= 1e-10 * r["xnorm"]
    x, x_ref = np.fromfile(xl), arr["lsqr_diag5000_x"]
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-10
    assert np.linalg.norm(case["d"] * x - case["b"]) < 1e-8 * np.linalg.norm(case["b"])      # it solved the system
    g, r = got["tnls_elem"], rec["tnls_elem5000"]
    assert g["status_code"] == r["status_code"]
    assert g["inner_iterations"] == r["inner_iterations"]
    assert np.allclose(g["trust_region_radius"], r["trust_region_radius"], rtol=1e-10, atol=0)
    # f = |F|^2 / 2 falls from 71 to 1e-13: the last values are sums of squares of residuals at the rounding level of F
    assert np.allclose(g["objective_values"], r["objective_values"], rtol=1e-8, atol=1e-12 * r["objective_values"][0])
    x, x_ref = np.fromfile(xt), arr["tnls_elem5000_x"]
    assert np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref) < 1e-10
    assert np.linalg.norm(x - case["xstar"]) < 1e-6 * np.linalg.norm(case["xstar"])           # the root it was built from


@pytest.mark.gpu
def test_device_lobpcg_header_entry_point(tmp_path):
    """LinearAlgebra::LOBPCG<std::vector<double>, DeviceMatrix>(A, B, T, X0, nev, max_iters, num_iters, nc, tau) with
    block-operator descriptor functors: the diagonal problems of the reference's LOBPCG unit tests against the CPU
    restatement and the exact spectrum."""
    from oracle import lobpcg_port as L
    exe = _compile("lobpcg_device_check", link=True)
    n, nx, nev = 1000, 10, 5
    ad, bd = np.linspace(-.5 * n, .5 * n, n), np.linspace(1.0, n, n)
    X0 = (2.0 * P.uniform01(91, 0, n * nx) - 1.0).reshape(n, nx)
    f = tmp_path / "lob.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<QQ", n, nx))
        for a in (ad, bd, X0):
            fh.write(np.ascontiguousarray(a).tobytes())
    got = _lines(subprocess.run([exe, str(f)], check=True, capture_output=True, text=True).stdout)
    # the reference's own header (compiled against the Eigen stand-in) on the same X0; the header layer draws the
    # reference's Gaussian probe block itself, so iteration counts are pinned exactly
    from oracle import refapi
    R = refapi.RefLobpcg()
    for name, Bop, exact in (("diag", None, ad[:nev]), ("diag_generalized", ("diag", bd), np.sort(ad / bd)[:nev])):
        th_ref, _, it_ref, nc_ref = R.lobpcg(("diag", ad), Bop, ("diag", np.abs(ad)), X0, nev, n, 1e-8)
        g = got[name]
        assert g["nc"] == nc_ref == nev and g["x_cols"] == nev                         # m x nev, like the reference
        assert np.linalg.norm(np.array(g["theta"]) - exact) < 1e-4                     # the reference tests' bar
        assert np.allclose(g["theta"], th_ref, rtol=1e-9, atol=1e-9)
        assert abs(g["num_iters"] - it_ref) <= 1        # (Gram rounding on a different machine may move the test by one)
    assert got["apply"]["max_err"] == 0.0
    assert got["invalid_argument"]["thrown"] == 2


def test_host_lobpcg_header_matches_reference_header(tmp_path):
    """OUR LOBPCG.h (dense path: RayleighRitz, soft locking, user function, random-start overload) against the
    reference's own LOBPCG.h, both compiled against the same Eigen stand-in (oracle/eigen_shim): iteration counts,
    converged counts, eigenvalues and eigenvectors bit for bit on the four problems of the reference's
    tests/LOBPCG_unit_test.cpp:123-208."""
    from oracle import refapi
    try:
        R = refapi.RefLobpcg()
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built here")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "lobpcg_host_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "oracle", "eigen_shim"),
                    os.path.join(ROOT, "tests", "host", "lobpcg_host_check.cpp"), "-o", exe], check=True)
    n, nx, nev = 1000, 10, 5
    ad, bd = np.linspace(-.5 * n, .5 * n, n), np.linspace(1.0, n, n)
    X0 = (2.0 * P.uniform01(91, 0, n * nx) - 1.0).reshape(n, nx)
    f = tmp_path / "lob.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<QQ", n, nx))
        for a in (ad, bd, X0):
            fh.write(np.ascontiguousarray(a).tobytes())
    got = _lines(subprocess.run([exe, str(f), str(tmp_path) + "/x_"], check=True, capture_output=True, text=True).stdout)
    for name, (gen, pre) in {"plain": (False, False), "precon": (False, True), "generalized_precon": (True, True),
                             "generalized": (True, False)}.items():
        th, X, it, nc = R.lobpcg(("diag", ad), ("diag", bd) if gen else None, ("diag", np.abs(ad)) if pre else None, X0,
                                 nev, 10 * n, 1e-8)
        g = got[name]
        assert (g["num_iters"], g["nc"], g["hook_calls"]) == (it, nc, it)
        assert np.array_equal(np.array(g["theta"]), th)
        assert np.array_equal(np.fromfile(str(tmp_path) + f"/x_{name}.bin").reshape(n, nev), X)
        exact = np.sort(ad / bd)[:nev] if gen else ad[:nev]
        assert np.linalg.norm(th - exact) < 1e-4                                       # the reference tests' bar
    assert got["hook_stop"]["num_iters"] == 7 and got["hook_stop"]["hook_calls"] == 7
    assert got["random_start"]["nc"] == 3 and got["random_start"]["x_cols"] == 3
    assert np.linalg.norm(np.array(got["random_start"]["theta"]) - ad[:3]) < 1e-4
    assert got["invalid_argument"]["thrown"] == 2
