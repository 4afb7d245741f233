"""Per-phase timers of the sharded persistent kernel (run under torchrun on the GPU box)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from optimization_b200 import problems as P
from optimization_b200.device import Context
from optimization_b200.sharded import ShardedStiefel
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
ctx = Context(local); ctx.connect(rank, world)
prob = P.make_stiefel_critical(100000, 32)
sh = ShardedStiefel(ctx, prob, rank, world)
kw = dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)
for _ in range(3): out = sh.solve_device(**kw)
ms = [sh.solve_device(**kw).solve_kernel_ms for _ in range(10)]
mx = (C.c_uint64 * 4)(); mn = (C.c_uint64 * 4)()
ctx.lib.ob200_debug_phase_times(ctx.h, 1, None, None)
o = sh.solve_device(**kw)
ctx.lib.ob200_debug_phase_times(ctx.h, 0, mx, mn)
k = o.num_iterations
if rank == 0:
    print(f"world={world} iters={k} kernel_ms={np.median(ms):.3f} us/iter={1e3*np.median(ms)/k:.2f}")
    print("   per-iter us  workA/waitA/workB/waitB  max:", [round(v / k / 1e3, 2) for v in mx], " min:", [round(v / k / 1e3, 2) for v in mn], flush=True)
dist.barrier(); dist.destroy_process_group(); ctx.close()
