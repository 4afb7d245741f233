import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimization_b200 import problems as P
from optimization_b200.device import Context
from oracle import refapi
port = refapi.PortOracle()
ctx = Context(0)
def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
for n in [int(a) for a in sys.argv[1:]] or [128, 64, 256, 1000]:
    prob = P.make_stiefel_critical(n, 32)
    A = torch.from_numpy(prob.A_bf16.view(np.int16)).cuda(); Y = ctx.to_device(prob.Y0); g = ctx.to_device(prob.g)
    H = ctx.stiefel_operator(A, Y)
    for mi in (1, 2, 60):
        kw = dict(Delta=1e6, max_iterations=mi, kappa_fgr=1e-9, theta=0.)
        s_ref, mn_ref, it_ref, why_ref = port.stpcg_stiefel(prob, prob.Y0, prob.g, **kw)
        for opt in (2, 1, 1, 1):
            ctx.set_option("tcgen05", opt)
            try:
                o = ctx.stpcg(g, H, **kw)
                print(f"n={n} maxit={mi} opt={opt} path={ctx.last_path} it={o.num_iterations}/{it_ref} exit={o.exit_reason}/{why_ref} rel={rel(o.s.cpu().numpy(), s_ref):.3e} mnorm={o.update_step_M_norm:.6e}/{mn_ref:.6e}", flush=True)
            except Exception as e:
                print(f"n={n} maxit={mi} opt={opt} EXC {e}", flush=True)
