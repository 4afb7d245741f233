"""LOBPCG at BASELINE config C4: 160^3 7-point Laplacian (m = 4 096 000), block nx = 64, nev = 32, Jacobi T = 1/6."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimization_b200 import problems as P
from optimization_b200.device import Context

g = int(sys.argv[1]) if len(sys.argv) > 1 else 160
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nx, nev = 64, 32
m = g ** 3
ctx = Context(0)
X0 = (2.0 * P._torch_uniform01(31, 0, m * nx, "cuda:0") - 1.0).view(m, nx)
A, T = ctx.block_laplacian3d(g, g, g), ctx.block_scalar(1.0 / 6.0)
ctx.lobpcg(A, None, T, X0, nev, 3, 1e-6)            # warm-up (allocations, cuSOLVER handle)
torch.cuda.synchronize()
l0 = ctx.kernel_launches
t0 = time.perf_counter()
th, X, it, nc = ctx.lobpcg(A, None, T, X0, nev, iters, 1e-6)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ns = 3 * nx
flops = it * (2 * 2 * m * ns * ns + 2 * 2 * m * ns * nx)        # two Grams + two block updates per iteration (full basis)
bytes_ = it * 8 * m * (ns * 2 + ns * 2 * 2 + ns * 2 + nx * 8)      # apply r/w, two Grams read S+Z, updates, residual (algorithmic)
lam1 = 2.0 - 2.0 * np.cos(np.arange(1, g + 1) * np.pi / (g + 1))
exact = np.sort((lam1[:, None, None] + lam1[None, :, None] + lam1[None, None, :]).ravel())[:nev]
print(f"LOBPCG C4 g={g} m={m} nx={nx} nev={nev}: {it} iterations in {dt * 1e3:.1f} ms = {dt / it * 1e3:.1f} ms/iteration "
      f"({it / dt:.1f} iterations/s), {flops / dt / 1e12:.2f} TFLOP/s fp64, ~{bytes_ / dt / 1e9:.0f} GB/s, converged {nc}/{nev}, "
      f"launches {ctx.kernel_launches - l0}")
print("   theta[:4] =", th[:4], " exact[:4] =", exact[:4], flush=True)
ctx.close()
