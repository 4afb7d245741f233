import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from optimization_b200 import problems as P
from optimization_b200.device import Context
ctx = Context(0)
ctx.set_option("tcgen05", 1)
n = 100000
prob = P.make_stiefel_critical(n, 32)
A = torch.from_numpy(prob.A_bf16.view(np.int16)).cuda(); Y = ctx.to_device(prob.Y0); g = ctx.to_device(prob.g)
H = ctx.stiefel_operator(A, Y)
kw = dict(Delta=1e6, max_iterations=int(os.environ.get("MAXIT", "12")), kappa_fgr=1e-9, theta=0.0)
for _ in range(3): out = ctx.stpcg(g, H, **kw)
mx = (C.c_uint64 * 4)(); mn = (C.c_uint64 * 4)()
ctx.lib.ob200_debug_phase_times(ctx.h, 1, None, None)
o = ctx.stpcg(g, H, **kw)
ctx.lib.ob200_debug_phase_times(ctx.h, 0, mx, mn)
k = o.num_iterations
print(f"iters={k} kernel_ms={o.solve_kernel_ms:.3f} us/iter={1e3*o.solve_kernel_ms/k:.2f}")
print("   per-iter us  workA/waitA/workB/waitB  max:", [round(v / k / 1e3, 2) for v in mx], " min:", [round(v / k / 1e3, 2) for v in mn], flush=True)
