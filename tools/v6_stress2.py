import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimization_b200 import problems as P
from optimization_b200.device import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
prob = P.make_stiefel_critical(n, 32)
kw = dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)
def setup(ctx):
    A = torch.from_numpy(prob.A_bf16.view(np.int16)).cuda(); Y = ctx.to_device(prob.Y0); g = ctx.to_device(prob.g)
    return g, ctx.stiefel_operator(A, Y), (A, Y)
ctx = Context(0)
g, H, keep = setup(ctx)
ctx.set_option("tcgen05", 2)
ref = ctx.stpcg(g, H, **kw).s.clone()
ctx.set_option("tcgen05", 1)
def rel(x): return float((x - ref).norm() / ref.norm())
print("same context, repeated v6 solves vs v4:", [f"{rel(ctx.stpcg(g, H, **kw).s):.1e}" for _ in range(8)], flush=True)
for mi in (1, 2, 3, 5, 10, 20):
    k2 = dict(kw, max_iterations=mi)
    ctx.set_option("tcgen05", 2); r2 = ctx.stpcg(g, H, **k2).s.clone(); ctx.set_option("tcgen05", 1)
    print(f"max_iterations={mi}:", [f"{float((ctx.stpcg(g, H, **k2).s - r2).norm() / r2.norm()):.1e}" for _ in range(6)], flush=True)
outs = []
for _ in range(4):
    c2 = Context(0); g2, H2, keep2 = setup(c2); c2.set_option("tcgen05", 1)
    outs.append([f"{rel(c2.stpcg(g2, H2, **kw).s):.1e}" for _ in range(2)])
    c2.close()
print("fresh context each time (first, second solve):", outs, flush=True)
