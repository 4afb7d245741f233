"""Multi-GPU check of the row-sharded pose-graph operator (run on the GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29521 tests/mgpu_posegraph_check.py
N-shard == 1-shard bit identity of the fused tCG solve on the block-CSR 3x3 Hessian (halo rows of p pushed over
NVLink inside the kernel, exact integer reductions), plus timing at N = 1e6 poses with --big."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optimization_b200 import problems as P  # noqa: E402
from optimization_b200.device import Context  # noqa: E402
from optimization_b200.sharded import ShardedPoseGraph, SinglePoseGraph  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = Context(local)
    ctx.connect(rank, world)
    ok = True
    cases = [((16, 16, 8), 4, False), ((32, 32, 16), 4, True), ((20, 20, 20), 5, False)]
    if "--big" in sys.argv:
        cases.append(((100, 100, 100), 4, True))
    for dims, r, consistent in cases:
        prob = P.make_posegraph(dims, r, sigma=0.0, x_noise=0.0) if consistent else P.make_posegraph(dims, r)
        g = prob.g
        if consistent:                       # right-hand side in the range of H
            from oracle import refapi
            port = refapi.PortOracle()
            lam, _, _ = port.csr3_model(prob, prob.X0)
            g = port.csr3_hess(prob, prob.X0, lam, prob.g) if prob.N <= 20000 else prob.g
        gn = float(np.linalg.norm(g))
        for kw in (dict(Delta=1e6 * gn, max_iterations=30, kappa_fgr=1e-9, theta=0.0),
                   dict(Delta=0.3 * gn, max_iterations=30, kappa_fgr=1e-3, theta=.5)):
            sh = ShardedPoseGraph(ctx, prob, rank, world, g=g)
            out = sh.solve_device(**kw)
            out = sh.solve_device(**kw)
            parts = [None] * world
            dist.all_gather_object(parts, (sh.lo, out.s.cpu().numpy(), out.num_iterations, out.exit_reason,
                                           out.update_step_M_norm, out.solve_kernel_ms))
            if rank == 0:
                s_full = np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])], axis=0)
                c1 = Context(local)
                one = SinglePoseGraph(c1, prob, g=g)
                o1 = one.solve_device(**kw)
                o1 = one.solve_device(**kw)
                s1 = o1.s.cpu().numpy()
                same = (o1.num_iterations == out.num_iterations and o1.exit_reason == out.exit_reason
                        and o1.update_step_M_norm == out.update_step_M_norm and np.array_equal(s1, s_full))
                kms = max(p[5] for p in parts)
                print(f"posegraph N={prob.N} r={r} world={world} iters={out.num_iterations}/{o1.num_iterations} "
                      f"exit={out.exit_reason} bit-identical={same} maxdiff={np.abs(s1 - s_full).max():.3e} "
                      f"us/iter sharded={1e3 * kms / max(out.num_iterations, 1):.1f} "
                      f"single={1e3 * o1.solve_kernel_ms / max(o1.num_iterations, 1):.1f}", flush=True)
                ok = ok and same
                c1.close()
            dist.barrier()
    if rank == 0:
        print("MGPU_POSEGRAPH_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
