"""Run-to-run bit reproducibility of the Stiefel tCG kernels (option from argv[1], default 1 = v6): repeated solves must
agree in every bit of s (exact integer reductions make the result independent of scheduling)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimization_b200 import problems as P
from optimization_b200.device import Context
opt = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sizes = [int(a) for a in sys.argv[2:]] or [5000, 20000, 100000]
ctx = Context(0)
ctx.set_option("tcgen05", opt)
bad = 0
for n in sizes:
    prob = P.make_stiefel_critical(n, 32)
    A = torch.from_numpy(prob.A_bf16.view(np.int16)).cuda(); Y = ctx.to_device(prob.Y0); g = ctx.to_device(prob.g)
    H = ctx.stiefel_operator(A, Y)
    kw = dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)
    ref = ctx.stpcg(g, H, **kw)
    mism = []
    for r in range(12):
        o = ctx.stpcg(g, H, **kw)
        if not (torch.equal(o.s, ref.s) and o.num_iterations == ref.num_iterations):
            d = (o.s - ref.s).abs()
            rows = torch.nonzero(d.amax(dim=1) > 0).flatten()
            mism.append((r, o.num_iterations, float(d.max()), int(rows.numel()), rows[:6].tolist()))
    bad += len(mism)
    print(f"n={n} opt={opt} path={ctx.last_path} iters={ref.num_iterations} mismatching runs: {len(mism)}/12 {mism[:3]}", flush=True)
print("V6_STRESS", "PASS" if bad == 0 else "FAIL", flush=True)
