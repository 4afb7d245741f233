"""Validate the tcgen05 digit-plane contraction W = A V against the fp64 tensor-core
path and numpy (run on the GPU box)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optimization_b200 import problems as P
from optimization_b200.device import Context, _ptr

ctx = Context(0)
ok = True
for n in (128, 1000, 4096, 100000):
    prob = P.make_stiefel(n, 32, y_noise=.2)
    A = torch.from_numpy(prob.A_bf16.view(np.int16)).cuda()
    rng = np.random.default_rng(n)
    for name, V in (("tangent", prob.g),
                    ("wide", rng.standard_normal((n, 32)) * np.exp(rng.uniform(-40, 40, (n, 1)))),
                    ("tiny", prob.g * 1e-200), ("huge", prob.g * 1e200),
                    ("sparse", prob.g * (rng.uniform(size=(n, 32)) < 0.05))):
        V = np.ascontiguousarray(V)
        Vd = ctx.to_device(V)
        o_tc = torch.empty_like(Vd); o_dm = torch.empty_like(Vd)
        rc = ctx.lib.ob200_debug_block_apply(ctx.h, n, _ptr(A), _ptr(Vd), _ptr(o_tc), 1); assert rc == 0, ctx.lib.ob200_last_error(ctx.h)
        rc = ctx.lib.ob200_debug_block_apply(ctx.h, n, _ptr(A), _ptr(Vd), _ptr(o_dm), 0); assert rc == 0
        tcv, dm = o_tc.cpu().numpy(), o_dm.cpu().numpy()
        # reference + magnitude scale per entry: sum |a||v|
        Ab = prob.A_dense_blocks()
        ref = np.zeros_like(V); mag = np.zeros_like(V)
        if n <= 4096:
            for b in range(prob.nblk):
                r0, r1 = b * 128, min(n, (b + 1) * 128)
                ref[r0:r1] = Ab[b, :r1 - r0, :r1 - r0] @ V[r0:r1]
                mag[r0:r1] = np.abs(Ab[b, :r1 - r0, :r1 - r0]) @ np.abs(V[r0:r1])
        else:
            ref, mag = dm, np.abs(dm) + 1e-300
        mag = np.maximum(mag, 1e-300)
        e_tc = np.max(np.abs(tcv - ref) / mag); e_dm = np.max(np.abs(dm - ref) / mag)
        # block-max relative error bound of the digit scheme: 2^-56 * 128 * max|a| * blockmax / mag
        print(f"n={n:6d} {name:8s} max err/sum|a||v|: tcgen05 {e_tc:.2e}  dmma {e_dm:.2e}  "
              f"rel diff tc-dmma {np.linalg.norm(tcv - dm) / max(np.linalg.norm(dm), 1e-300):.2e}", flush=True)
        if not np.isfinite(e_tc) or np.linalg.norm(tcv - dm) > 1e-13 * np.linalg.norm(dm):
            ok = False
print("TC_CHECK", "PASS" if ok else "FAIL")
