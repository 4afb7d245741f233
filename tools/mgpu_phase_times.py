"""Per-phase times of the row-sharded Stiefel tCG kernel (run under torchrun on N GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/mgpu_phase_times.py
Every rank prints, per CG iteration and over its CTAs (max / min): work in phase A, wait at reduction A (arrival ->
release, i.e. local barrier + machine-wide exchange + skew between the GPUs), work in the scalar stage + phase B, wait at
reduction B."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optimization_b200 import problems as P  # noqa: E402
from optimization_b200.device import Context  # noqa: E402
from optimization_b200.sharded import ShardedStiefel, SingleStiefel  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
ctx = Context(local)
prob = P.make_stiefel_critical(100000, 32)
kw = dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)
if world > 1:
    ctx.connect(rank, world)
    sh = ShardedStiefel(ctx, prob, rank, world)
else:
    sh = SingleStiefel(ctx, prob)
for _ in range(3):
    out = sh.solve_device(**kw)
mx = (C.c_uint64 * 4)()
mn = (C.c_uint64 * 4)()
if world > 1:
    dist.barrier()
ctx.lib.ob200_debug_phase_times(ctx.h, 1, None, None)
o = sh.solve_device(**kw)
ctx.lib.ob200_debug_phase_times(ctx.h, 0, mx, mn)
k = o.num_iterations
print(f"rank {rank}/{world}: iters={k} kernel {1e3 * o.solve_kernel_ms / k:.2f} us/iter; per-iteration us "
      f"[workA, waitA, workB, waitB] max over CTAs {[round(v / k / 1e3, 2) for v in mx]} min {[round(v / k / 1e3, 2) for v in mn]}",
      flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
ctx.close()
