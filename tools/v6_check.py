"""v6 kernel bring-up: parity against the C port at small sizes, then timing against v4 at the BASELINE size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimization_b200 import problems as P
from optimization_b200.device import Context
from oracle import refapi
import subprocess
subprocess.run(["make", "-s", "-C", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"), "port"], check=True)
port = refapi.PortOracle()
ctx = Context(0)
ctx.set_option("tcgen05", 1)
def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
sizes = [int(a) for a in sys.argv[1:]] or [128, 64, 300, 1000, 4096, 20000]
ok = True
for n in sizes:
    for prob, kws in ((P.make_stiefel_critical(n, 32), [dict(Delta=1e6, max_iterations=60, kappa_fgr=1e-9, theta=0.), dict(Delta=1e6, max_iterations=3, kappa_fgr=1e-9, theta=0.)]),
                      (P.make_stiefel(n, 32, y_noise=.2), [dict(Delta=3.0, max_iterations=60, kappa_fgr=1e-3, theta=.5)])):
        A = torch.from_numpy(prob.A_bf16.view(np.int16)).cuda(); Y = ctx.to_device(prob.Y0); g = ctx.to_device(prob.g)
        H = ctx.stiefel_operator(A, Y)
        for kw in kws:
            s_ref, mn_ref, it_ref, why_ref = port.stpcg_stiefel(prob, prob.Y0, prob.g, **kw)
            o = ctx.stpcg(g, H, **kw)
            o2 = ctx.stpcg(g, H, **kw)
            e = rel(o.s.cpu().numpy(), s_ref)
            good = (o.num_iterations, o.exit_reason) == (it_ref, why_ref) and e < 1e-10 and torch.equal(o.s, o2.s)
            ok &= good
            print(f"n={n} path={ctx.last_path} it={o.num_iterations}/{it_ref} exit={o.exit_reason}/{why_ref} rel={e:.2e} det={torch.equal(o.s, o2.s)} {'OK' if good else 'FAIL'}", flush=True)
print("V6_CHECK", "PASS" if ok else "FAIL", flush=True)
if "--notime" not in sys.argv:
    n = 100000
    prob = P.make_stiefel_critical(n, 32)
    A = torch.from_numpy(prob.A_bf16.view(np.int16)).cuda(); Y = ctx.to_device(prob.Y0); g = ctx.to_device(prob.g)
    H = ctx.stiefel_operator(A, Y)
    kw = dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)
    import ctypes as C
    for opt in (1, 2):
        ctx.set_option("tcgen05", opt)
        for _ in range(3): out = ctx.stpcg(g, H, **kw)
        ms = [ctx.stpcg(g, H, **kw).solve_kernel_ms for _ in range(10)]
        mx = (C.c_uint64 * 4)(); mn = (C.c_uint64 * 4)()
        ctx.lib.ob200_debug_phase_times(ctx.h, 1, None, None)
        o = ctx.stpcg(g, H, **kw)
        ctx.lib.ob200_debug_phase_times(ctx.h, 0, mx, mn)
        k = o.num_iterations
        print(f"opt={opt} path={ctx.last_path} iters={k} kernel_ms={np.median(ms):.3f} us/iter={1e3*np.median(ms)/k:.2f} GB/s={H.step_bytes()*k/np.median(ms)/1e6:.0f}")
        print("   per-iter us  workA/waitA/workB/waitB  max:", [round(v / k / 1e3, 2) for v in mx], " min:", [round(v / k / 1e3, 2) for v in mn], flush=True)
    ctx.set_option("tcgen05", 1)
