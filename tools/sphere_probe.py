"""Throughput of the fused sphere tCG kernel at config C2 (n = 2^24, k = 16) and of the stand-alone HVP."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from optimization_b200 import problems as P
from optimization_b200.device import Context

ctx = Context(0)
for n in [int(a) for a in sys.argv[1:]] or [1 << 24]:
    k = 16
    d, Ut, sigma, x0, g = P.make_sphere_critical_device(n, k, device="cuda:0")
    H = ctx.sphere_operator(d, None, sigma, x0, Ut=Ut)
    gn = math.sqrt(ctx.dot(g, g))
    kw = dict(Delta=1e6 * gn, max_iterations=200, kappa_fgr=1e-9, theta=0.)
    for _ in range(2):
        out = ctx.stpcg(g, H, **kw)
    ms = sorted(ctx.stpcg(g, H, **kw).solve_kernel_ms for _ in range(5))[2]
    it = out.num_iterations
    sb = H.step_bytes()
    print(f"sphere n={n} k={k}: {it} iterations, kernel {ms:.3f} ms, {1e3 * ms / it:.1f} us/iteration, "
          f"algorithmic {sb / 1e6:.1f} MB/step -> {sb * it / ms / 1e6:.0f} GB/s", flush=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hv = ctx.hvp(H, g)
    torch.cuda.synchronize()
    ev0.record(ctx.stream)
    for _ in range(5):
        ctx.hvp(H, g, out=hv)
    ev1.record(ctx.stream)
    torch.cuda.synchronize()
    t = ev0.elapsed_time(ev1) / 5
    print(f"   stand-alone HVP: {t:.3f} ms, {H.hvp_bytes() / 1e6:.1f} MB -> {H.hvp_bytes() / t / 1e6:.0f} GB/s", flush=True)
ctx.close()
