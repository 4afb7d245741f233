"""Multi-GPU check (run on the GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py [n]
N-shard == 1-shard bit identity of the fused tCG solve (exact integer reductions
exchanged over NVLink peer memory), plus parity with the CPU oracle at small n."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optimization_b200 import problems as P  # noqa: E402
from optimization_b200.device import Context  # noqa: E402
from optimization_b200.sharded import ShardedSphere, ShardedStiefel, SingleSphere, SingleStiefel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ok = True
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 20000, 100000]
    ctx = Context(local)
    ctx.connect(rank, world)
    for n in sizes:
        for maker, kw in ((P.make_stiefel_critical, dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)),
                          (lambda n_: P.make_stiefel(n_, 32, y_noise=.2), dict(Delta=3.0, max_iterations=60, kappa_fgr=1e-3, theta=.5))):
            prob = maker(n)
            sh = ShardedStiefel(ctx, prob, rank, world)
            out = sh.solve_device(**kw)
            s_loc = out.s.cpu().numpy()
            # gather the shards on rank 0
            parts = [None] * world
            dist.all_gather_object(parts, (sh.lo, sh.hi, s_loc, out.num_iterations, out.exit_reason, out.update_step_M_norm))
            if rank == 0:
                s_full = np.concatenate([p[2] for p in sorted(parts)], axis=0)
                its = {p[3] for p in parts}
                assert len(its) == 1
                c1 = Context(local)                       # independent single-GPU context, whole problem
                single = SingleStiefel(c1, prob)
                o1 = single.solve_device(**kw)
                s1 = o1.s.cpu().numpy()
                same = (o1.num_iterations == out.num_iterations and o1.exit_reason == out.exit_reason
                        and o1.update_step_M_norm == out.update_step_M_norm and np.array_equal(s1, s_full))
                print(f"n={n} world={world} iters={out.num_iterations}/{o1.num_iterations} exit={out.exit_reason} "
                      f"bit-identical={same} maxdiff={np.abs(s1 - s_full).max():.3e}", flush=True)
                ok = ok and same
                c1.close()
            dist.barrier()
    # sphere Rayleigh Hessian (three exact reductions per CG step, 16 low-rank sums in the first one)
    for n in [4096, 100003, 1 << 20]:
        prob = P.make_sphere_critical(n, 16)
        gn = float(np.linalg.norm(prob.g))
        for kw in (dict(Delta=1e6 * gn, max_iterations=60, kappa_fgr=1e-10, theta=0.0),
                   dict(Delta=0.3 * gn, max_iterations=60, kappa_fgr=1e-3, theta=.5)):
            sh = ShardedSphere(ctx, prob, rank, world)
            out = sh.solve_device(**kw)
            parts = [None] * world
            dist.all_gather_object(parts, (sh.lo, out.s.cpu().numpy(), out.num_iterations, out.exit_reason,
                                           out.update_step_M_norm))
            if rank == 0:
                s_full = np.concatenate([p[1] for p in sorted(parts, key=lambda t: t[0])])
                c1 = Context(local)
                o1 = SingleSphere(c1, prob).solve_device(**kw)
                s1 = o1.s.cpu().numpy()
                same = (o1.num_iterations == out.num_iterations and o1.exit_reason == out.exit_reason
                        and o1.update_step_M_norm == out.update_step_M_norm and np.array_equal(s1, s_full))
                print(f"sphere n={n} world={world} iters={out.num_iterations}/{o1.num_iterations} exit={out.exit_reason} "
                      f"bit-identical={same} maxdiff={np.abs(s1 - s_full).max():.3e}", flush=True)
                ok = ok and same
                c1.close()
            dist.barrier()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
