"""GPU: parity of the fused CUDA tCG path (through the C ABI) against the CPU
oracle (oracle/stpcg_port.c, pinned by tests/test_oracle.py) and the committed
golden fixtures from the unmodified reference headers.

Bar (BASELINE.json north_star): iteration counts and exit reasons bit-exact,
iterates within 1e-10 relative.  RTOL below is that tolerance."""
import math

import numpy as np
import pytest

from optimization_b200 import problems as P

pytestmark = pytest.mark.gpu
RTOL = 1e-10
DBL_MAX = 1.7976931348623157e308


@pytest.fixture(scope="module")
def ctx():
    from optimization_b200.device import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def port(build_oracle):
    from oracle import refapi
    return refapi.PortOracle()


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def run_diag(ctx, g, h, minv=None, **kw):
    H = ctx.diag_operator(ctx.to_device(h))
    out = ctx.stpcg(ctx.to_device(g), H, minv=None if minv is None else ctx.to_device(minv), **kw)
    return out.s.cpu().numpy(), out


# ---- the reference's own known-answer tests -----------------------------------
def test_kats_against_golden(ctx, golden):
    rec, _ = golden
    for name in ("ExactSTPCG", "ExactSTPCGwithNegativeCurvature", "ExactSTPCGwithPreconditioning",
                 "ExactSTPCGwithNegativeCurvatureAndPreconditioning"):
        a = rec[name]["args"]
        minv = None if a["minv"] is None else np.array(a["minv"])
        s, out = run_diag(ctx, np.array(a["g"]), np.array(a["h"]), minv, Delta=a["Delta"],
                          max_iterations=a["max_iterations"], kappa_fgr=a["kappa_fgr"], theta=a["theta"])
        assert out.num_iterations == rec[name]["num_iterations"], name
        assert rel(s, rec[name]["s"]) < RTOL, name
        assert abs(out.update_step_M_norm - rec[name]["update_step_M_norm"]) <= RTOL * rec[name]["update_step_M_norm"]


def test_kat_closed_forms(ctx):
    g3, H3 = np.array([21., -.4, 19.]), np.array([1000., 100., 1.])
    s, out = run_diag(ctx, g3, H3, Delta=DBL_MAX, max_iterations=3, kappa_fgr=1e-8, theta=.999)
    assert np.linalg.norm(s + g3 / H3) < 1e-6 and out.num_iterations == 3
    s, out = run_diag(ctx, g3, -H3, Delta=1000., max_iterations=3, kappa_fgr=1e-8, theta=.999)
    assert np.linalg.norm(s + 1000. * g3 / np.linalg.norm(g3)) < 1e-6
    assert out.num_iterations == 0 and out.exit_reason == "boundary" and out.update_step_M_norm == 1000.


# ---- diagonal operator, seeded data --------------------------------------------
def test_diag_golden(ctx, golden):
    rec, arr = golden
    dp = P.make_diag(1000, seed=5)
    for name, minv in (("diag1000_trunc", None), ("diag1000_precon_trunc", dp.minv),
                       ("diag1000_tight", None), ("diag1000_boundary", None)):
        s, out = run_diag(ctx, dp.g, dp.h, minv, **rec[name]["args"])
        assert out.num_iterations == rec[name]["num_iterations"], name
        assert rel(s, arr[name + "_s"]) < RTOL, name
        assert abs(out.update_step_M_norm - rec[name]["update_step_M_norm"]) <= RTOL * rec[name]["update_step_M_norm"]


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 4099, 100003, 1 << 20])
@pytest.mark.parametrize("precon", [False, True])
def test_diag_vs_oracle_ragged(ctx, port, n, precon):
    dp = P.make_diag(n, seed=7 + n % 13)
    minv = dp.minv if precon else None
    for kw in (dict(Delta=1e6, max_iterations=40, kappa_fgr=1e-10, theta=0.),
               dict(Delta=1e-3, max_iterations=40, kappa_fgr=.1, theta=.5),
               dict(Delta=1e6, max_iterations=3, kappa_fgr=1e-10, theta=0.)):
        s_ref, mn_ref, it_ref, why_ref = port.stpcg_diag(dp.g, dp.h, minv, **kw)
        s, out = run_diag(ctx, dp.g, dp.h, minv, **kw)
        assert out.num_iterations == it_ref
        assert out.exit_reason == why_ref
        assert rel(s, s_ref) < RTOL
        assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)


@pytest.mark.parametrize("precon", ["none", "jacobi", "callback"])
def test_host_callback_operator_runs_the_reference_loop(ctx, port, precon):
    """OB200_OP_HOST_CALLBACK / OB200_PRECON_HOST_CALLBACK (the unfused fallback for arbitrary Hessian /
    preconditioner functors): ob200_stpcg runs IterativeSolvers.h:211-424 on the host over device vectors, the functors
    being host callbacks.  Same iteration counts / exits / steps as the oracle and as the fused diagonal kernel, on a
    positive definite, an indefinite and a singular operator."""
    import torch
    n = 4099
    dp = P.make_diag(n, seed=5)
    g = ctx.to_device(dp.g)
    minv_np = dp.minv if precon != "none" else None
    minv = ctx.to_device(dp.minv)
    for h_np, kws in ((dp.h, (dict(Delta=1e6, max_iterations=40, kappa_fgr=1e-10, theta=0.),
                              dict(Delta=1e-3, max_iterations=40, kappa_fgr=.1, theta=.5),
                              dict(Delta=1e6, max_iterations=3, kappa_fgr=1e-10, theta=0.))),
                      (np.where(np.arange(n) % 7 == 0, -dp.h, dp.h), (dict(Delta=50., max_iterations=100, kappa_fgr=1e-8, theta=0.),)),
                      (np.zeros(n), (dict(Delta=50., max_iterations=100, kappa_fgr=1e-8, theta=0.),))):
        h = ctx.to_device(h_np)
        calls = {"H": 0, "P": 0}

        def apply_H(v, out, h=h):
            calls["H"] += 1
            torch.mul(h.view_as(v), v, out=out)

        def apply_P(r, out):
            calls["P"] += 1
            torch.mul(minv.view_as(r), r, out=out)

        Hcb = ctx.callback_operator(n, 1, apply_H)
        pc = {"none": None, "jacobi": minv, "callback": ctx.callback_precon(n, 1, apply_P)}[precon]
        for kw in kws:
            s_ref, mn_ref, it_ref, why_ref = port.stpcg_diag(dp.g, h_np, minv_np, **kw)
            out = ctx.stpcg(g, Hcb, minv=pc, **kw)
            assert ctx.last_path == "generic"
            assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref), kw
            assert rel(out.s.cpu().numpy().ravel(), s_ref) < RTOL
            assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)
            fused = ctx.stpcg(g, ctx.diag_operator(h), minv=(minv if precon != "none" else None), **kw)
            assert (fused.num_iterations, fused.exit_reason) == (out.num_iterations, out.exit_reason)
            assert rel(out.s.cpu().numpy().ravel(), fused.s.cpu().numpy().ravel()) < RTOL
        assert calls["H"] > 0 and (calls["P"] > 0) == (precon == "callback")


def test_diag_indefinite_and_kernel(ctx, port):
    n = 5000
    dp = P.make_diag(n, seed=3)
    h = dp.h.copy()
    h[::7] *= -1.0                     # indefinite: negative curvature exit
    kw = dict(Delta=50., max_iterations=100, kappa_fgr=1e-8, theta=0.)
    s_ref, mn_ref, it_ref, why_ref = port.stpcg_diag(dp.g, h, None, **kw)
    s, out = run_diag(ctx, dp.g, h, None, **kw)
    assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
    assert rel(s, s_ref) < RTOL and out.update_step_M_norm == mn_ref == 50.
    # H = 0: p lies in ker(H) (IterativeSolvers.h:305-337)
    z = np.zeros(n)
    s_ref, mn_ref, it_ref, why_ref = port.stpcg_diag(dp.g, z, None, **kw)
    s, out = run_diag(ctx, dp.g, z, None, **kw)
    assert why_ref == "kernel" and (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
    assert rel(s, s_ref) < RTOL


def test_invalid_arguments_raise(ctx):
    g3, H3 = np.array([21., -.4, 19.]), np.array([1000., 100., 1.])
    for kw in (dict(Delta=0.), dict(Delta=-1.), dict(kappa_fgr=1.), dict(kappa_fgr=-.1), dict(theta=1.5),
               dict(theta=-.1), dict(epsilon=0.), dict(epsilon=1.)):
        args = dict(Delta=1., max_iterations=3, kappa_fgr=.1, theta=.5, epsilon=1e-8)
        args.update(kw)
        with pytest.raises(ValueError):
            run_diag(ctx, g3, H3, **args)


# ---- exact reductions ---------------------------------------------------------------
def test_dot_is_correctly_rounded(ctx):
    rng = np.random.default_rng(0)
    for n in (1, 31, 256, 1000, 65537, 3_200_000):
        a = rng.standard_normal(n) * np.exp(rng.uniform(-20, 20, n))
        b = rng.standard_normal(n)
        got = ctx.dot(ctx.to_device(a), ctx.to_device(b))
        # unit partials are fma-chains of 256 elements: compare with fsum of exact products (tolerance 2 ulp of
        # the partial-sum magnitude), and check bitwise run-to-run determinism
        want = math.fsum((a * b).tolist())
        scale = math.fsum(np.abs(a * b).tolist())
        assert abs(got - want) <= 1e-13 * scale
        assert got == ctx.dot(ctx.to_device(a), ctx.to_device(b))


# ---- Stiefel block-diagonal Hessian ----------------------------------------------------
def stiefel_setup(ctx, prob):
    import torch
    A = torch.from_numpy(prob.A_bf16.astype(np.int16)).to("cuda:0")   # bit pattern
    Y = ctx.to_device(prob.Y0)
    return A, Y, ctx.stiefel_operator(A, Y)


@pytest.mark.parametrize("tag,maker", [("stiefel512_yn1", lambda: P.make_stiefel(512, 32, y_noise=.1)),
                                       ("stiefel512_yn3", lambda: P.make_stiefel(512, 32, y_noise=.3)),
                                       ("stiefel1000_yn1", lambda: P.make_stiefel(1000, 32, y_noise=.1)),
                                       ("stiefelcrit512", lambda: P.make_stiefel_critical(512, 32)),
                                       ("stiefelcrit1000", lambda: P.make_stiefel_critical(1000, 32))])
def test_stiefel_stpcg_golden(ctx, golden, tag, maker):
    rec, arr = golden
    prob = maker()
    A, Y, H = stiefel_setup(ctx, prob)
    if tag + "_S" in arr:
        assert rel(H.S, arr[tag + "_S"]) < RTOL
        assert abs(H.f - rec[tag + "_f"]) <= RTOL * abs(rec[tag + "_f"])
    for name in ("tight", "default", "boundary"):
        r = rec[f"{tag}_{name}"]
        out = ctx.stpcg(ctx.to_device(prob.g), H, **r["args"])
        assert out.num_iterations == r["num_iterations"], name
        s_ref = arr[f"{tag}_{name}_s"]
        if np.all(np.isfinite(s_ref)):
            assert rel(out.s.cpu().numpy(), s_ref) < RTOL, name
        assert abs(out.update_step_M_norm - r["update_step_M_norm"]) <= RTOL * r["update_step_M_norm"], name


@pytest.mark.parametrize("n", [128, 300, 4096, 20000])
def test_stiefel_stpcg_vs_oracle(ctx, port, n):
    for prob in (P.make_stiefel_critical(n, 32), P.make_stiefel(n, 32, y_noise=.2)):
        A, Y, H = stiefel_setup(ctx, prob)
        for kw in (dict(Delta=1e6, max_iterations=60, kappa_fgr=1e-9, theta=0.),
                   dict(Delta=3.0, max_iterations=60, kappa_fgr=1e-3, theta=.5)):
            s_ref, mn_ref, it_ref, why_ref = port.stpcg_stiefel(prob, prob.Y0, prob.g, **kw)
            out = ctx.stpcg(ctx.to_device(prob.g), H, **kw)
            assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
            assert rel(out.s.cpu().numpy(), s_ref) < RTOL
            assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)


def test_stiefel_hvp_model_retract_golden(ctx, golden):
    rec, arr = golden
    tag = "stiefel512_yn1"
    prob = P.make_stiefel(512, 32, y_noise=.1)
    A, Y, H = stiefel_setup(ctx, prob)
    hv = ctx.hvp(H, ctx.to_device(prob.g)).cpu().numpy()
    assert rel(hv, arr[tag + "_hess_g"]) < RTOL
    S, f, grad, bound = ctx.stiefel_model(A, Y)
    assert rel(S, arr[tag + "_S"]) < RTOL and rel(grad.cpu().numpy(), arr[tag + "_grad"]) < RTOL
    # retraction: orthonormal columns, same column space as Y + V, first-order agreement
    V = ctx.to_device(0.01 * prob.g)
    Q = ctx.stiefel_retract(Y, V).cpu().numpy()
    assert np.linalg.norm(Q.T @ Q - np.eye(32)) < 1e-12
    Z = prob.Y0 + 0.01 * prob.g
    Qn, Rn = np.linalg.qr(Z)
    Qn = Qn * np.sign(np.diag(Rn))[None, :]
    assert rel(Q, Qn) < 1e-10


# ---- BASELINE size: size-independent properties ------------------------------------------
def test_stiefel_full_size_properties(ctx):
    import torch
    prob = P.make_stiefel_critical(100000, 32)
    A, Y, H = stiefel_setup(ctx, prob)
    g = ctx.to_device(prob.g)
    kw = dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.)
    o1 = ctx.stpcg(g, H, **kw)
    s1 = o1.s.clone()
    o2 = ctx.stpcg(g, H, **kw)
    assert o1.exit_reason == "residual" and 20 < o1.num_iterations < 200
    # bitwise run-to-run determinism (exact reductions)
    assert o1.num_iterations == o2.num_iterations and o1.update_step_M_norm == o2.update_step_M_norm
    assert torch.equal(s1, o2.s)
    # residual reduction ||g + H s|| <= kappa ||g||  (the property the reference tests, :254-275)
    Hs = ctx.hvp(H, s1)
    r = (g + Hs)
    assert math.sqrt(ctx.dot(r, r)) <= 1.0001 * 1e-9 * math.sqrt(ctx.dot(g, g))
    # reported M-norm equals the Frobenius norm of s (recurrence, IterativeSolvers.h:415-424)
    assert abs(o1.update_step_M_norm - math.sqrt(ctx.dot(s1, s1))) <= 1e-9 * o1.update_step_M_norm
    # s is tangent at Y: sym(Y^T s) = 0
    G = (Y.T @ s1).cpu().numpy()
    assert np.linalg.norm(G + G.T) <= 1e-9 * np.linalg.norm(s1.cpu().numpy())
    # linearity of the HVP
    v = ctx.to_device(P.make_stiefel(100000, 32, y_noise=.1).g)
    lhs = ctx.hvp(H, ctx.axpby(2.0, g, -3.0, v)).cpu().numpy()
    rhs = 2.0 * ctx.hvp(H, g).cpu().numpy() - 3.0 * ctx.hvp(H, v).cpu().numpy()
    assert rel(lhs, rhs) < 1e-12
    # host-buffer entry gives the same result as the device entry
    oh = ctx.stpcg(prob.g, H, host=True, **kw)
    assert oh.num_iterations == o1.num_iterations and np.array_equal(oh.s, s1.cpu().numpy())


def _host_oracle():
    """The reference's own STPCG header (oracle/_ref, all host cores) when the prebuilt library travelled with the
    snapshot, else the single-threaded C restatement."""
    import os
    from oracle import refapi
    try:
        R = refapi.RefOracle()
        R.set_threads(len(os.sched_getaffinity(0)))
        return R, None
    except (FileNotFoundError, OSError):
        return None, refapi.PortOracle()


@pytest.mark.parametrize("which", ["critical", "noisy"])
def test_stiefel_full_size_vs_reference(ctx, build_oracle, which):
    """BASELINE size (config C3, Stiefel(100000, 32)): the fused kernel against the reference's STPCG on the host --
    iteration count and exit bit-exact, iterate and M-norm within 1e-10 (reference IterativeSolvers.h:285-422)."""
    prob = P.make_stiefel_critical(100000, 32) if which == "critical" else P.make_stiefel(100000, 32, y_noise=.2)
    A, Y, H = stiefel_setup(ctx, prob)
    cases = [dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.)] if which == "critical" else \
            [dict(Delta=3.0, max_iterations=60, kappa_fgr=1e-3, theta=.5), dict(Delta=1e6, max_iterations=25, kappa_fgr=1e-9, theta=0.)]
    R, port = _host_oracle()
    rs = R.stiefel(prob) if R else None
    for kw in cases:
        if rs:
            s_ref, mn_ref, it_ref = rs.stpcg(prob.Y0, prob.g, **kw)
        else:
            s_ref, mn_ref, it_ref, _ = port.stpcg_stiefel(prob, prob.Y0, prob.g, **kw)
        out = ctx.stpcg(ctx.to_device(prob.g), H, **kw)
        assert ctx.last_path.startswith("tcgen05")
        assert out.num_iterations == it_ref
        assert rel(out.s.cpu().numpy(), s_ref) < RTOL
        assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)
        if which == "critical":
            assert out.exit_reason == "residual" and 20 < it_ref < 200


@pytest.mark.parametrize("n", [512, 4099, 20000])
def test_stiefel_projected_jacobi_vs_reference(ctx, build_oracle, n):
    """OB200_PRECON_STIEFEL_PROJECTED_JACOBI (v = P_Y(minv o r), the preconditioner TNT's adapter TNT.h:413-426 hands
    to STPCG for a tangent-space preserving precon): the C ABI's unfused loop with the one-launch device HVP against
    the reference's own STPCG header with the same preconditioner functor -- counts exact, iterate within 1e-10 -- and
    against the same loop driven through host callbacks."""
    import torch
    R, _ = _host_oracle()
    if R is None:
        pytest.skip("oracle/_ref (the compiled reference headers) did not travel with this snapshot")
    minv_np = P.stiefel_row_scaling(n, 32)
    for prob, kws in ((P.make_stiefel_critical(n, 32), (dict(Delta=1e6, max_iterations=60, kappa_fgr=1e-9, theta=0.),
                                                         dict(Delta=1e6, max_iterations=3, kappa_fgr=1e-9, theta=0.))),
                      (P.make_stiefel(n, 32, y_noise=.2), (dict(Delta=3.0, max_iterations=60, kappa_fgr=1e-3, theta=.5),))):
        A, Y, H = stiefel_setup(ctx, prob)
        minv = ctx.to_device(minv_np)
        g = ctx.to_device(prob.g)
        rs = R.stiefel(prob)

        def apply_H(v, out, H=H):
            out.copy_(ctx.hvp(H, v.contiguous()))

        def apply_P(r, out, Y=Y):
            z = minv * r
            G = Y.t() @ z
            out.copy_(z - Y @ (0.5 * (G + G.t())))

        for kw in kws:
            s_ref, mn_ref, it_ref = rs.stpcg(prob.Y0, prob.g, minv=minv_np, projected=True, **kw)
            out = ctx.stpcg(g, H, minv=ctx.projected_jacobi(minv), **kw)
            assert ctx.last_path == "generic"
            assert out.num_iterations == it_ref, kw
            assert rel(out.s.cpu().numpy(), s_ref) < RTOL
            assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)
            # the step stays in the tangent space at Y: sym(Y^T s) = 0
            YtS = (Y.t() @ out.s).cpu().numpy()
            assert np.abs(YtS + YtS.T).max() <= 1e-12 * max(1.0, float(out.s.abs().max()))
            cb = ctx.stpcg(g, ctx.callback_operator(n, 32, apply_H), minv=ctx.callback_precon(n, 32, apply_P), **kw)
            # (the callbacks round differently -- torch matmul instead of the exact reductions -- so a count may move by
            # one at a residual threshold)
            assert abs(cb.num_iterations - out.num_iterations) <= 1
            if cb.num_iterations == out.num_iterations:
                assert rel(cb.s.cpu().numpy(), out.s.cpu().numpy()) < 1e-6
    # wrong pairing is refused, not mis-computed
    Hd = ctx.diag_operator(ctx.to_device(np.ones(n * 32)))
    with pytest.raises(Exception):
        ctx.stpcg(g, Hd, minv=ctx.projected_jacobi(minv), Delta=1.0)


def test_sphere_full_size_vs_reference(ctx, build_oracle):
    """BASELINE size (config C2, n = 2^24, k = 16): same inputs on both sides (generated on the device, copied to the
    host), CG iterations capped so the CPU side finishes in seconds; counts / exits exact, iterate within 1e-10."""
    import types
    n, k = 1 << 24, 16
    d, Ut, sigma, x0, g = P.make_sphere_critical_device(n, k, device="cuda:0")
    H = ctx.sphere_operator(d, None, sigma, x0, Ut=Ut)
    prob = types.SimpleNamespace(n=n, k=k, d=d.cpu().numpy(), U=np.ascontiguousarray(Ut.t().contiguous().cpu().numpy()),
                                 sigma=sigma, x0=x0.cpu().numpy(), g=g.cpu().numpy())
    gn = float(np.linalg.norm(prob.g))
    R, port = _host_oracle()
    for kw in (dict(Delta=1e6 * gn, max_iterations=6, kappa_fgr=1e-10, theta=0.),
               dict(Delta=.2 * gn, max_iterations=40, kappa_fgr=.1, theta=.5)):
        if R:
            s_ref, mn_ref, it_ref = R.sphere_stpcg(prob, prob.x0, prob.g, **kw)
        else:
            s_ref, mn_ref, it_ref, _ = port.stpcg_sphere(prob, prob.x0, prob.g, None, **kw)
        out = ctx.stpcg(g, H, **kw)
        assert out.num_iterations == it_ref
        assert rel(out.s.cpu().numpy(), s_ref) < RTOL
        assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)


@pytest.mark.parametrize("n", [128, 300, 1000, 4096 + 32])
def test_block_apply_tcgen05_readbacks_match_fp64(ctx, n):
    """A V for the block-diagonal bf16 A: the exact digit-plane tcgen05 contraction with both TMEM read-back
    arrangements (one lane per thread; 16-lane fragments, the persistent kernel's) against the fp64 tensor-core
    product and numpy.  Exercises the row order of the A images (tc_row_of_lane)."""
    import ctypes as C
    import torch
    from optimization_b200.device import _ptr
    prob = P.make_stiefel(n, 32, y_noise=.2)
    A, Y, H = stiefel_setup(ctx, prob)
    V = ctx.to_device(prob.g * np.exp(P.uniform01(5, 0, n)[:, None] * 8.0 - 4.0))       # row scales over e^8
    outs = []
    for mode in (0, 1, 2):
        o = torch.zeros_like(V)
        ctx._check(ctx.lib.ob200_debug_block_apply(ctx.h, n, _ptr(A), _ptr(V), _ptr(o), mode))
        outs.append(o.cpu().numpy())
    Ad = P.from_bf16_bits(prob.A_bf16)
    Vh = V.cpu().numpy()
    want = np.zeros_like(Vh)
    for b in range(prob.nblk):
        r0, r1 = 128 * b, min(n, 128 * b + 128)
        want[r0:r1] = Ad[b, :r1 - r0, :r1 - r0] @ Vh[r0:r1]
    for o in outs:
        assert rel(o, want) < 1e-14
    assert np.array_equal(outs[1], outs[2])          # same integers, same recombination: bit-identical


@pytest.mark.parametrize("n", [33, 64, 65, 127, 129, 191, 193, 1000, 4099, 18977])
def test_stiefel_v6_ragged_sizes_vs_oracle(ctx, port, n):
    """The default (warp-specialised, rotated-basis) kernel on row counts around the 64-row half / 128-row block / 8-row
    strip boundaries: partial halves, a single half, last block of 1 .. 127 rows."""
    for prob, kw in ((P.make_stiefel_critical(n, 32), dict(Delta=1e6, max_iterations=80, kappa_fgr=1e-9, theta=0.)),
                     (P.make_stiefel(n, 32, y_noise=.2), dict(Delta=3.0, max_iterations=60, kappa_fgr=1e-3, theta=.5))):
        A, Y, H = stiefel_setup(ctx, prob)
        s_ref, mn_ref, it_ref, why_ref = port.stpcg_stiefel(prob, prob.Y0, prob.g, **kw)
        out = ctx.stpcg(ctx.to_device(prob.g), H, **kw)
        assert ctx.last_path == "tcgen05"
        assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
        assert rel(out.s.cpu().numpy(), s_ref) < RTOL
        assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)


def test_stiefel_v6_is_bit_reproducible_and_rotation_cache_follows_the_point(ctx, port):
    """Repeated solves agree in every bit (exact integer reductions; this is the check that found the cross-proxy race
    of the stage, tools/v6_stress.py), and the cached rotation (Q, Lambda, Y Q) is dropped when the point changes under
    the same device address."""
    import torch
    n = 20000
    p1 = P.make_stiefel_critical(n, 32, seed=21)
    p2 = P.make_stiefel(n, 32, y_noise=.2)
    kw = dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.)
    A = torch.from_numpy(p1.A_bf16.astype(np.int16)).to("cuda:0")
    Y = ctx.to_device(p1.Y0)
    H = ctx.stiefel_operator(A, Y)
    g = ctx.to_device(p1.g)
    first = ctx.stpcg(g, H, **kw)
    for _ in range(8):
        o = ctx.stpcg(g, H, **kw)
        assert o.num_iterations == first.num_iterations and torch.equal(o.s, first.s)
    s_ref, _, it_ref, why_ref = port.stpcg_stiefel(p1, p1.Y0, p1.g, **kw)
    assert (first.num_iterations, first.exit_reason) == (it_ref, why_ref) and rel(first.s.cpu().numpy(), s_ref) < RTOL
    # same buffers, new point (and new A): the rotated copy of Y must be rebuilt
    A.copy_(torch.from_numpy(p2.A_bf16.astype(np.int16)))
    Y.copy_(torch.from_numpy(p2.Y0).to(Y.device))
    H2 = ctx.stiefel_operator(A, Y)
    kw2 = dict(Delta=3.0, max_iterations=60, kappa_fgr=1e-3, theta=.5)
    o2 = ctx.stpcg(ctx.to_device(p2.g), H2, **kw2)
    s2, _, it2, why2 = port.stpcg_stiefel(p2, p2.Y0, p2.g, **kw2)
    assert (o2.num_iterations, o2.exit_reason) == (it2, why2) and rel(o2.s.cpu().numpy(), s2) < RTOL


def test_stiefel_non_finite_input_is_reported(ctx):
    """A NaN in g poisons the exact accumulators: the solve must end with an error status, never with a silent answer."""
    prob = P.make_stiefel_critical(1000, 32)
    A, Y, H = stiefel_setup(ctx, prob)
    g = prob.g.copy()
    g[517, 3] = np.nan
    with pytest.raises(Exception):
        out = ctx.stpcg(ctx.to_device(g), H, Delta=1e6, max_iterations=20, kappa_fgr=1e-9, theta=0.)
        assert not np.all(np.isfinite(out.s.cpu().numpy()))     # (reached only if no status was raised)
        raise RuntimeError("non-finite solve returned a status of success")


def test_planes_cache_follows_the_matrix(ctx, port):
    """The cached tcgen05 digit planes are keyed by pointer, size AND content: an A rewritten in place at the same
    address (what a caching allocator produces when one problem replaces another) must not meet stale planes."""
    import torch
    p1 = P.make_stiefel_critical(1024, 32, seed=21)
    p2 = P.make_stiefel_critical(1024, 32, seed=33)
    kw = dict(Delta=1e6, max_iterations=60, kappa_fgr=1e-9, theta=0.)
    A = torch.from_numpy(p1.A_bf16.astype(np.int16)).to("cuda:0")
    for prob in (p1, p2, p1):
        A.copy_(torch.from_numpy(prob.A_bf16.astype(np.int16)))          # same device address, new content
        Y = ctx.to_device(prob.Y0)
        H = ctx.stiefel_operator(A, Y)
        out = ctx.stpcg(ctx.to_device(prob.g), H, **kw)
        s_ref, mn_ref, it_ref, why_ref = port.stpcg_stiefel(prob, prob.Y0, prob.g, **kw)
        assert ctx.last_path.startswith("tcgen05")
        assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
        assert rel(out.s.cpu().numpy(), s_ref) < RTOL
    # the stand-alone HVP validates its planes on the device: rewrite A again and call it FIRST
    A.copy_(torch.from_numpy(p2.A_bf16.astype(np.int16)))
    H2 = ctx.stiefel_operator(A, ctx.to_device(p2.Y0))
    hv = ctx.hvp(H2, ctx.to_device(p2.g)).cpu().numpy()
    assert rel(hv, P.stiefel_hess_numpy(p2, p2.Y0, p2.g)) < 1e-12


@pytest.mark.parametrize("n", [64, 128, 300, 1000, 20000])
def test_stiefel_hvp_fused_vs_numpy(ctx, n):
    """ob200_hvp for the Stiefel operator: ONE persistent launch (contraction + Gram | projection), against the dense
    numpy Hessian and against the fp64 tensor-core path (ob200_set_option('tcgen05', 0))."""
    for prob in (P.make_stiefel(n, 32, y_noise=.2), P.make_stiefel_critical(n, 32)):
        A, Y, H = stiefel_setup(ctx, prob)
        V = ctx.to_device(prob.g)
        hv = ctx.hvp(H, V).cpu().numpy()
        want = P.stiefel_hess_numpy(prob, prob.Y0, prob.g)
        assert rel(hv, want) < 1e-12
        ctx.set_option("tcgen05", 0)
        try:
            hv0 = ctx.hvp(H, V).cpu().numpy()
        finally:
            ctx.set_option("tcgen05", 1)
        assert rel(hv0, want) < 1e-12
        z = ctx.hvp(H, ctx.to_device(np.zeros_like(prob.g))).cpu().numpy()
        assert np.all(z == 0.0)


# ---- sphere Rayleigh-quotient Hessian, A = diag + low rank (configs C1 / C2) ------------------
def sphere_setup(ctx, prob):
    d, U, x = ctx.to_device(prob.d), ctx.to_device(prob.U), ctx.to_device(prob.x0)
    return ctx.sphere_operator(d, U, prob.sigma, x)


@pytest.mark.parametrize("n,k", [(1000, 16), (4099, 5)])
def test_sphere_stpcg_golden(ctx, golden, n, k):
    rec, arr = golden
    prob = P.make_sphere_critical(n, k)
    H = sphere_setup(ctx, prob)
    for name in ("tight", "default", "boundary"):
        r = rec[f"spherecrit{n}_k{k}_{name}"]
        out = ctx.stpcg(ctx.to_device(prob.g), H, **r["args"])
        assert out.num_iterations == r["num_iterations"], name
        assert rel(out.s.cpu().numpy(), arr[f"spherecrit{n}_k{k}_{name}_s"]) < RTOL, name
        assert abs(out.update_step_M_norm - r["update_step_M_norm"]) <= RTOL * r["update_step_M_norm"], name


@pytest.mark.parametrize("n,k", [(1, 0), (2, 1), (255, 16), (256, 16), (257, 3), (4099, 16), (100003, 16),
                                 (1 << 20, 16), (70000, 0)])
@pytest.mark.parametrize("precon", [False, True])
def test_sphere_stpcg_vs_oracle_ragged(ctx, port, n, k, precon):
    prob = P.make_sphere_critical(n, k) if n > 2 else P.make_sphere(n, k)
    H = sphere_setup(ctx, prob)
    minv = 1.0 / (1.0 + P.uniform01(77, 0, n)) if precon else None
    gn = float(np.linalg.norm(prob.g)) or 1.0
    for kw in (dict(Delta=1e6 * gn, max_iterations=40, kappa_fgr=1e-10, theta=0.),
               dict(Delta=.2 * gn, max_iterations=40, kappa_fgr=.1, theta=.5),
               dict(Delta=1e6 * gn, max_iterations=3, kappa_fgr=1e-10, theta=0.)):
        s_ref, mn_ref, it_ref, why_ref = port.stpcg_sphere(prob, prob.x0, prob.g, minv, **kw)
        out = ctx.stpcg(ctx.to_device(prob.g), H, minv=None if minv is None else ctx.to_device(minv), **kw)
        assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
        assert rel(out.s.cpu().numpy(), s_ref) < RTOL
        assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)


def test_sphere_indefinite_exits_like_oracle(ctx, port):
    # random base point: the Rayleigh Hessian is indefinite, tCG leaves through the boundary (l.347-361)
    prob = P.make_sphere(20000, 16)
    H = sphere_setup(ctx, prob)
    for Delta in (0.5, 50.0, 1e4):
        kw = dict(Delta=Delta, max_iterations=100, kappa_fgr=1e-8, theta=0.)
        s_ref, mn_ref, it_ref, why_ref = port.stpcg_sphere(prob, prob.x0, prob.g, **kw)
        out = ctx.stpcg(ctx.to_device(prob.g), H, **kw)
        assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
        assert rel(out.s.cpu().numpy(), s_ref) < RTOL and out.update_step_M_norm == mn_ref


def test_sphere_model_hvp_retract(ctx, port):
    prob = P.make_sphere(30011, 16)
    H = sphere_setup(ctx, prob)
    f_ref, Ax_ref, grad_ref = port.sphere_model(prob, prob.x0)
    Ax, f, grad = ctx.sphere_model(ctx.to_device(prob.d), H.Ut, prob.sigma, ctx.to_device(prob.x0))
    assert abs(f - f_ref) <= RTOL * abs(f_ref)
    assert rel(Ax.cpu().numpy(), Ax_ref) < RTOL and rel(grad.cpu().numpy(), grad_ref) < RTOL
    hv = ctx.hvp(H, ctx.to_device(prob.g)).cpu().numpy()
    assert rel(hv, port.sphere_hess(prob, prob.x0, prob.g)) < RTOL
    # projection retraction (reference example: (x + v) / |x + v|)
    v = 0.01 * prob.g
    q = ctx.sphere_retract(ctx.to_device(prob.x0), ctx.to_device(v)).cpu().numpy()
    z = prob.x0 + v
    assert rel(q, z / np.linalg.norm(z)) < 1e-14 and abs(np.linalg.norm(q) - 1.0) < 1e-14


def test_sphere_full_size_properties(ctx):
    # config C2: n = 2^24, k = 16 (U alone is 2 GB), generated on the device
    import torch
    n, k = 1 << 24, 16
    d, Ut, sigma, x0, g = P.make_sphere_critical_device(n, k, device="cuda:0")
    H = ctx.sphere_operator(d, None, sigma, x0, Ut=Ut)
    gn = math.sqrt(ctx.dot(g, g))
    kw = dict(Delta=1e6 * gn, max_iterations=200, kappa_fgr=1e-9, theta=0.)
    o1 = ctx.stpcg(g, H, **kw)
    s1 = o1.s.clone()
    o2 = ctx.stpcg(g, H, **kw)
    assert o1.exit_reason == "residual" and 5 < o1.num_iterations < 100
    assert o1.num_iterations == o2.num_iterations and o1.update_step_M_norm == o2.update_step_M_norm
    assert torch.equal(s1, o2.s)                               # bitwise run-to-run determinism
    r = g + ctx.hvp(H, s1)                                     # residual reduction (:254-275)
    assert math.sqrt(ctx.dot(r, r)) <= 1.0001 * 1e-9 * gn
    assert abs(o1.update_step_M_norm - math.sqrt(ctx.dot(s1, s1))) <= 1e-9 * o1.update_step_M_norm
    assert abs(ctx.dot(x0, s1)) <= 1e-9 * o1.update_step_M_norm    # s is tangent at x
    v = torch.roll(g, 12345)
    v = v - x0 * ctx.dot(x0, v)
    lhs = ctx.hvp(H, ctx.axpby(2.0, g, -3.0, v)).cpu().numpy()
    rhs = 2.0 * ctx.hvp(H, g).cpu().numpy() - 3.0 * ctx.hvp(H, v).cpu().numpy()
    assert rel(lhs, rhs) < 1e-12                               # linearity of the HVP


# ---- the fp64 tensor-core fallback of the Stiefel kernel ------------------------------------------
@pytest.mark.parametrize("n", [300, 4096])
def test_stiefel_fp64_mma_path_vs_oracle(ctx, port, n):
    """ob200_set_option("tcgen05", 0): tcg_stiefel_kernel (A p on the fp64 tensor cores) against the oracle and
    against the tcgen05 digit-plane kernel."""
    prob = P.make_stiefel_critical(n, 32)
    A, Y, H = stiefel_setup(ctx, prob)
    kw = dict(Delta=1e6, max_iterations=60, kappa_fgr=1e-9, theta=0.)
    s_ref, mn_ref, it_ref, why_ref = port.stpcg_stiefel(prob, prob.Y0, prob.g, **kw)
    o_tc = ctx.stpcg(ctx.to_device(prob.g), H, **kw)
    assert ctx.last_path == "tcgen05"       # default: the warp-specialised v6 kernel (solve in the eigenbasis of S)
    ctx.set_option("tcgen05", 2)            # the previous generation of the tcgen05 kernel (kept as an option)
    o_v5 = ctx.stpcg(ctx.to_device(prob.g), H, **kw)
    assert ctx.last_path == "tcgen05_v4"
    ctx.set_option("tcgen05", 0)
    try:
        o_mm = ctx.stpcg(ctx.to_device(prob.g), H, **kw)
        assert ctx.last_path == "dmma"
    finally:
        ctx.set_option("tcgen05", 1)
    for o in (o_tc, o_v5, o_mm):
        assert (o.num_iterations, o.exit_reason) == (it_ref, why_ref)
        assert rel(o.s.cpu().numpy(), s_ref) < RTOL
        assert abs(o.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)
    assert rel(o_tc.s.cpu().numpy(), o_mm.s.cpu().numpy()) < RTOL


def test_stiefel_wide_range_A_takes_fallback(ctx, port):
    """A block whose entries span more than 22 bits is not block-fixed-point: the library must notice and run the
    fp64 tensor-core kernel by itself (no option set), still matching the oracle."""
    import dataclasses
    prob = P.make_stiefel_critical(1024, 32)
    Ad = P.from_bf16_bits(prob.A_bf16).copy()          # (nblk, 128, 128) doubles on the bf16 grid
    Ad[0, 5, 9] = Ad[0, 9, 5] = 2.0 ** -40             # exact in bf16, 2^-40 relative to the block's O(1) entries
    prob2 = dataclasses.replace(prob, A_bf16=P.to_bf16_bits(Ad))
    A, Y, H = stiefel_setup(ctx, prob2)
    kw = dict(Delta=1e6, max_iterations=60, kappa_fgr=1e-9, theta=0.)
    s_ref, mn_ref, it_ref, why_ref = port.stpcg_stiefel(prob2, prob2.Y0, prob2.g, **kw)
    out = ctx.stpcg(ctx.to_device(prob2.g), H, **kw)
    assert ctx.last_path == "dmma"
    assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
    assert rel(out.s.cpu().numpy(), s_ref) < RTOL


# ---- LOBPCG (reference LinearAlgebra/LOBPCG.h; BASELINE config C4) ------------------------------------------
LOB_N, LOB_M, LOB_NEV, LOB_TAU = 1000, 10, 5, 1e-8


def _lob_x0(m, nx, seed=91):
    return (2.0 * P.uniform01(seed, 0, m * nx) - 1.0).reshape(m, nx)


@pytest.mark.parametrize("generalized,precon", [(False, False), (False, True), (True, True), (True, False)])
def test_lobpcg_reference_unit_test_problems(ctx, generalized, precon):
    """The four diagonal problems of the reference's tests/LOBPCG_unit_test.cpp:123-208 (n = 1000, block 10, nev 5,
    tau 1e-8) against the CPU restatement (same X0 and probe block) and the exact spectrum."""
    from oracle import refapi
    R = refapi.RefLobpcg()      # the reference's own LOBPCG.h compiled against the Eigen stand-in (oracle/_ref)
    adiag, bdiag = np.linspace(-.5 * LOB_N, .5 * LOB_N, LOB_N), np.linspace(1.0, LOB_N, LOB_N)
    X0 = _lob_x0(LOB_N, LOB_M)
    Om = R.omega(LOB_N, LOB_M)  # the Gaussian probe block the reference draws (same generator, same order)
    th_ref, X_ref, it_ref, nc_ref = R.lobpcg(("diag", adiag), ("diag", bdiag) if generalized else None,
                                             ("diag", np.abs(adiag)) if precon else None, X0, LOB_NEV, 10 * LOB_N, LOB_TAU)
    dA = ctx.block_diag(ctx.to_device(adiag))
    dB = ctx.block_diag(ctx.to_device(bdiag)) if generalized else None
    dT = ctx.block_diag(ctx.to_device(np.abs(adiag))) if precon else None
    th, X, it, nc = ctx.lobpcg(dA, dB, dT, ctx.to_device(X0), LOB_NEV, 10 * LOB_N, LOB_TAU, Omega=ctx.to_device(Om))
    exact = np.sort(adiag / bdiag)[:LOB_NEV] if generalized else adiag[:LOB_NEV]
    assert nc == nc_ref == LOB_NEV
    assert np.linalg.norm(th - exact) < 1e-4 and np.linalg.norm(th_ref - exact) < 1e-4       # the reference's bar
    assert np.allclose(th, th_ref, rtol=1e-9, atol=1e-9)
    assert abs(it - it_ref) <= 1          # pinned to the reference header (Gram rounding may move the test by one)


def test_lobpcg_small_problem_with_literal_x0(ctx):
    # LOBPCG_unit_test.cpp:94-120
    lam = np.array([1., 2., 3., 4.])
    X0 = np.array([[0.8147, 0.6324], [0.9058, 0.0975], [0.1270, 0.2785], [0.9134, 0.5469]])
    th, X, it, nc = ctx.lobpcg(ctx.block_diag(ctx.to_device(lam)), None, None, ctx.to_device(X0), 2, 1000, 1e-8)
    assert nc == 2 and np.linalg.norm(th - lam[:2]) < 1e-3
    with pytest.raises(ValueError):
        ctx.lobpcg(ctx.block_diag(ctx.to_device(lam)), None, None, ctx.to_device(X0), 3, 10)     # nev > nx (l.148)


def test_lobpcg_laplacian_vs_oracle_and_exact_spectrum(ctx):
    """Config C4 at small size: 7-point Dirichlet Laplacian on a 12 x 10 x 9 grid, Jacobi T = 1/6, block 16."""
    from oracle import lobpcg_port as L
    gx, gy, gz, nx, nev = 12, 10, 9, 16, 6
    m = gx * gy * gz
    lam = lambda g: 2.0 - 2.0 * np.cos(np.arange(1, g + 1) * np.pi / (g + 1))
    exact = np.sort((lam(gz)[:, None, None] + lam(gy)[None, :, None] + lam(gx)[None, None, :]).ravel())
    X0 = _lob_x0(m, nx, seed=31)
    dA = ctx.block_laplacian3d(gx, gy, gz)
    # the stencil kernel against the numpy operator
    AX = ctx.block_apply(dA, ctx.to_device(X0)).cpu().numpy()
    assert rel(AX, P.laplacian3d_apply(X0, gx, gy, gz)) < 1e-14
    from oracle import refapi
    R = refapi.RefLobpcg()
    th_ref, _, it_ref, nc_ref = R.lobpcg(("laplacian", (gx, gy, gz)), None, ("scalar", 1.0 / 6.0), X0, nev, 500, 1e-8)
    th, X, it, nc = ctx.lobpcg(dA, None, ctx.block_scalar(1.0 / 6.0), ctx.to_device(X0), nev, 500, 1e-8,
                               Omega=ctx.to_device(R.omega(m, nx)))
    assert nc == nc_ref == nev
    assert np.allclose(th, exact[:nev], rtol=1e-7) and np.allclose(th, th_ref, rtol=1e-9)
    assert abs(it - it_ref) <= 1
    Xh = X.cpu().numpy()
    R = P.laplacian3d_apply(Xh, gx, gy, gz) - Xh * th[None, :]
    assert np.all(np.linalg.norm(R, axis=0) <= 1e-6 * np.linalg.norm(Xh, axis=0))


def test_lobpcg_full_size_properties(ctx):
    """Config C4: 160^3 Laplacian (m = 4 096 000), block 64, Jacobi 1/6, a fixed number of iterations; size-independent
    checks against the analytic spectrum: Ritz values are upper bounds of the exact eigenvalues (Cauchy interlacing),
    every Ritz value lies within its residual norm of an exact eigenvalue, B-orthonormal Ritz vectors, determinism."""
    import torch
    g, nx, nev, iters = 160, 64, 32, 8
    m = g ** 3
    X0 = (2.0 * P._torch_uniform01(31, 0, m * nx, "cuda:0") - 1.0).view(m, nx)
    A, T = ctx.block_laplacian3d(g, g, g), ctx.block_scalar(1.0 / 6.0)
    th, X, it, nc = ctx.lobpcg(A, None, T, X0, nev, iters, 1e-6)
    th2, X2, it2, nc2 = ctx.lobpcg(A, None, T, X0, nev, iters, 1e-6)
    assert it == it2 == iters and np.array_equal(th, th2) and torch.equal(X, X2)         # bitwise reproducible
    lam1 = 2.0 - 2.0 * np.cos(np.arange(1, g + 1) * np.pi / (g + 1))
    exact = np.sort((lam1[:, None, None] + lam1[None, :, None] + lam1[None, None, :]).ravel())
    assert np.all(np.diff(th) >= 0) and np.all(th >= exact[:nev] * (1 - 1e-12))           # interlacing
    Xc = X.contiguous()
    R = ctx.block_apply(A, Xc) - Xc * torch.from_numpy(th).to(Xc.device)[None, :]
    rn = (torch.linalg.norm(R, dim=0) / torch.linalg.norm(Xc, dim=0)).cpu().numpy()
    dist = np.min(np.abs(th[:, None] - exact[None, :4096]), axis=1)
    assert np.all(dist <= rn * (1 + 1e-9))                                                # |theta - lambda| <= |r| / |x|
    G = (Xc.T @ Xc).cpu().numpy()
    assert np.linalg.norm(G - np.eye(nev)) < 1e-9                                         # X^T B X = I (B = I)


# ---- sparse Hessian families: rotation synchronisation on St(3,r)^N (config C5), 7-point Laplacian (config C4) --------
def _pose_dev(ctx, prob):
    import torch
    rp = torch.from_numpy(prob.rowptr.astype(np.int64)).cuda()
    ci = torch.from_numpy(prob.colidx.astype(np.int32)).cuda()
    bl = ctx.to_device(prob.blocks)
    X = ctx.to_device(prob.X0)
    return rp, ci, bl, X


@pytest.mark.parametrize("dims,r,consistent", [((8, 7, 6), 4, False), ((8, 7, 6), 3, True), ((6, 6, 6), 5, False),
                                               ((3, 1, 1), 8, False), ((20, 20, 25), 4, False), ((50, 50, 40), 4, True)])
def test_csr3_model_hvp_stpcg_vs_oracle(ctx, port, dims, r, consistent):
    """Block-CSR 3x3 Hessian of the rotation-synchronisation cost (SE-Sync shape): model (Lambda, f, gradient), stand-alone
    HVP and the fused tCG against the C restatement (itself bit-identical to the reference's STPCG on the same operator,
    tests/test_oracle.py), N = 3 ... 1e5 poses, r = 3 ... 8, with and without the Jacobi preconditioner."""
    prob = P.make_posegraph(dims, r, sigma=0.0, x_noise=0.0) if consistent else P.make_posegraph(dims, r)
    rp, ci, bl, X = _pose_dev(ctx, prob)
    lam_ref, f_ref, grad_ref = port.csr3_model(prob, prob.X0)
    lam, f, grad = ctx.csr3_model(rp, ci, bl, X)
    assert abs(f - f_ref) <= 1e-12 * max(abs(f_ref), float(prob.N))
    assert np.abs(lam.cpu().numpy() - lam_ref).max() <= 1e-13 * max(1.0, np.abs(lam_ref).max())
    assert rel(grad.cpu().numpy(), grad_ref) < 1e-13 or np.linalg.norm(grad_ref) < 1e-10
    H = ctx.csr3_operator(rp, ci, bl, X)
    lam_h = H.Lambda.cpu().numpy()
    hv_ref = port.csr3_hess(prob, prob.X0, lam_h, prob.g)
    hv = ctx.hvp(H, ctx.to_device(prob.g)).cpu().numpy()
    assert rel(hv, hv_ref) < 1e-13
    # right-hand side: in the range of H for the consistent problem (the gauge directions X_i Omega are its kernel)
    g = hv_ref if consistent else prob.g
    gn = float(np.linalg.norm(g))
    minv = 1.0 / (1.0 + P.uniform01(77, 0, g.size)).reshape(g.shape)
    for kw, mv in ((dict(Delta=1e6 * gn, max_iterations=40, kappa_fgr=1e-9, theta=0.), None),
                   (dict(Delta=.3 * gn, max_iterations=40, kappa_fgr=1e-3, theta=.5), None),
                   (dict(Delta=1e6 * gn, max_iterations=25, kappa_fgr=1e-9, theta=0.), minv)):
        s_ref, mn_ref, it_ref, why_ref = port.stpcg_csr3(prob, prob.X0, lam_h, g, mv, **kw)
        out = ctx.stpcg(ctx.to_device(g), H, minv=None if mv is None else ctx.to_device(mv), **kw)
        assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
        assert rel(out.s.cpu().numpy(), s_ref) < RTOL
        assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)
    # retraction: every pose block has orthonormal rows, first-order agreement with X + V
    V = ctx.to_device(0.01 * prob.g)
    Xr = ctx.csr3_retract(X, V).cpu().numpy().reshape(prob.N, 3, r)
    assert np.abs(Xr @ Xr.transpose(0, 2, 1) - np.eye(3)).max() < 1e-13
    Z = (prob.X0 + 0.01 * prob.g).reshape(prob.N, 3, r)
    Qn = np.stack([np.linalg.qr(z.T)[0] * np.sign(np.diag(np.linalg.qr(z.T)[1]))[None, :] for z in Z[:50]])
    assert np.abs(Xr[:50] - Qn.transpose(0, 2, 1)).max() < 1e-12


@pytest.mark.parametrize("dims,p", [((9, 8, 7), 3), ((1, 1, 5), 1), ((32, 32, 32), 4), ((64, 48, 40), 8)])
def test_stencil7_operator_stpcg_vs_oracle(ctx, port, dims, p):
    """The 7-point Dirichlet Laplacian (config C4's operator) as a tCG Hessian descriptor, with the Jacobi
    preconditioner 1/6 of that configuration."""
    n = dims[0] * dims[1] * dims[2]
    g = (2.0 * P.uniform01(5, 0, n * p) - 1.0).reshape(n, p)
    H = ctx.stencil7_operator(*dims, p)
    hv = ctx.hvp(H, ctx.to_device(g)).cpu().numpy()
    assert np.array_equal(hv, port.stencil7_apply(dims, p, g))
    assert rel(hv, P.laplacian3d_apply(g, *dims)) < 1e-14
    minv = np.full((n, p), 1.0 / 6.0)
    for kw, mv in ((dict(Delta=1e9, max_iterations=60, kappa_fgr=1e-9, theta=0.), None),
                   (dict(Delta=1e9, max_iterations=60, kappa_fgr=1e-9, theta=0.), minv),
                   (dict(Delta=.5, max_iterations=60, kappa_fgr=1e-3, theta=.5), minv)):
        s_ref, mn_ref, it_ref, why_ref = port.stpcg_stencil7(dims, p, g, mv, **kw)
        out = ctx.stpcg(ctx.to_device(g), H, minv=None if mv is None else ctx.to_device(mv), **kw)
        assert (out.num_iterations, out.exit_reason) == (it_ref, why_ref)
        assert rel(out.s.cpu().numpy(), s_ref) < RTOL
        assert abs(out.update_step_M_norm - mn_ref) <= RTOL * abs(mn_ref)
