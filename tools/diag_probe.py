"""Throughput of the fused diagonal-Hessian tCG kernel (the reference's own STPCG test shape, scaled up)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimization_b200.device import Context

ctx = Context(0)
for n in [int(a) for a in sys.argv[1:]] or [1 << 24, 1 << 26]:
    g = torch.rand(n, dtype=torch.float64, device="cuda") * 2 - 1
    h = 1000.0 + 2000.0 * torch.rand(n, dtype=torch.float64, device="cuda")
    minv = 1.0 / (1000.0 + 2000.0 * torch.rand(n, dtype=torch.float64, device="cuda"))
    H = ctx.diag_operator(h)
    for name, mv in (("plain", None), ("jacobi", minv)):
        kw = dict(Delta=1e300, max_iterations=30, kappa_fgr=1e-300, theta=0.0, minv=mv)
        for _ in range(2):
            out = ctx.stpcg(g, H, **kw)
        ms = sorted(ctx.stpcg(g, H, **kw).solve_kernel_ms for _ in range(5))[2]
        sb = H.step_bytes(precon=mv is not None)
        print(f"diag n={n} {name}: {out.num_iterations} iterations ({out.exit_reason}), {1e3 * ms / out.num_iterations:.1f} us/iteration, "
              f"{sb / 1e6:.0f} MB/step -> {sb * out.num_iterations / ms / 1e6:.0f} GB/s", flush=True)
ctx.close()
