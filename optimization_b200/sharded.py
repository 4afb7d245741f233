"""Workload drivers used by bench.py and the multi-GPU tests: one Stiefel tCG
problem resident on one GPU (`SingleStiefel`) or row-sharded over the ranks of
a torch.distributed job (`ShardedStiefel`).  Host-side plumbing only."""
from __future__ import annotations

import numpy as np
import torch


def _bits(A_bf16: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(A_bf16).view(np.int16)).to(device)


class SingleStiefel:
    kernel_name = "tcg_stiefel_tc_kernel (persistent fused tCG, whole solve; tcgen05 int8 digit planes + fp64 MMA)"

    def __init__(self, ctx, prob):
        self.ctx, self.prob = ctx, prob
        dev = f"cuda:{ctx.device}"
        self.A = _bits(prob.A_bf16, dev)
        self.Y = ctx.to_device(prob.Y0)
        self.g = ctx.to_device(prob.g)
        self.s = torch.empty_like(self.g)
        self.H = ctx.stiefel_operator(self.A, self.Y)
        self.g_host = torch.from_numpy(prob.g).pin_memory()
        self.s_host = torch.empty_like(self.g_host).pin_memory()

    def solve_device(self, **kw):
        return self.ctx.stpcg(self.g, self.H, s_out=self.s, **kw)

    def solve_host(self, **kw):
        return self.ctx.stpcg(self.g_host, self.H, s_out=self.s_host, host=True, **kw)

    def step_bytes_total(self):
        return self.H.step_bytes()

    def ncu_traffic_per_launch(self, iters_per_launch):
        """DRAM bytes per launch of the persistent kernel from the committed `ncu --set full` capture
        (profiles/traffic.json), scaled to this run's iteration count; None if no capture matches."""
        import json
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
        try:
            with open(path) as fh:
                rec = json.load(fh)["tcg_stiefel_tc_kernel"]
        except Exception:
            return None
        if rec.get("n") != self.prob.n or getattr(self, "world", 1) != 1:
            return None
        return rec["dram_bytes_per_launch"] / rec["cg_iterations_per_launch"] * iters_per_launch

    def hvp_device(self):
        return self.ctx.hvp(self.H, self.g)


def row_partition(n: int, world: int, nb: int = 128):
    """Contiguous row blocks, aligned to the operator's nb x nb blocks:
    rank r owns rows [lo[r], hi[r])."""
    nblk = (n + nb - 1) // nb
    lo = [min(n, nb * (nblk * r // world)) for r in range(world)]
    hi = [min(n, nb * (nblk * (r + 1) // world)) for r in range(world)]
    return lo, hi


class ShardedStiefel(SingleStiefel):
    """One Stiefel tCG problem row-sharded over the ranks of a torch.distributed
    job; the reductions travel through NVLink peer memory inside the kernels."""
    kernel_name = "tcg_stiefel_tc_kernel (persistent fused tCG + in-kernel NVLink peer exchange)"

    def __init__(self, ctx, prob, rank, world):
        self.ctx, self.prob, self.rank, self.world = ctx, prob, rank, world
        if getattr(ctx, "world", 1) != world:
            ctx.connect(rank, world)
        lo, hi = row_partition(prob.n, world, prob.nb)
        self.lo, self.hi = lo[rank], hi[rank]
        b0, b1 = self.lo // prob.nb, (self.hi + prob.nb - 1) // prob.nb
        dev = f"cuda:{ctx.device}"
        self.A = _bits(prob.A_bf16[b0:b1], dev)
        self.Y = ctx.to_device(prob.Y0[self.lo:self.hi])
        self.g = ctx.to_device(prob.g[self.lo:self.hi])
        self.s = torch.empty_like(self.g)
        self.H = ctx.stiefel_operator(self.A, self.Y)          # S, bound: global (exchanged)
        self.g_host = torch.from_numpy(np.ascontiguousarray(prob.g[self.lo:self.hi])).pin_memory()
        self.s_host = torch.empty_like(self.g_host).pin_memory()

    def step_bytes_total(self):
        # whole-problem algorithmic bytes of one CG step = sum over ranks
        n, p = self.prob.n, self.prob.p
        return 12 * 8 * n * p + self.prob.nblk * self.prob.nb * self.prob.nb * 2


class SingleSphere:
    """Sphere Rayleigh-quotient tCG problem (config C1 / C2 family) resident on one GPU."""

    def __init__(self, ctx, prob, lo=0, hi=None):
        hi = prob.n if hi is None else hi
        self.ctx, self.prob, self.lo, self.hi = ctx, prob, lo, hi
        self.d = ctx.to_device(prob.d[lo:hi])
        self.U = ctx.to_device(np.ascontiguousarray(prob.U[lo:hi]))
        self.x = ctx.to_device(prob.x0[lo:hi])
        self.g = ctx.to_device(prob.g[lo:hi])
        self.H = ctx.sphere_operator(self.d, self.U, prob.sigma, self.x)      # A x, x^T A x: global (exchanged)

    def solve_device(self, **kw):
        return self.ctx.stpcg(self.g, self.H, **kw)


class ShardedSphere(SingleSphere):
    """The same problem row-sharded over the ranks (shard boundaries at multiples of 256 elements, the unit
    of the exact reduction, so N-GPU runs are bit-identical to 1-GPU runs)."""

    def __init__(self, ctx, prob, rank, world):
        if getattr(ctx, "world", 1) != world:
            ctx.connect(rank, world)
        lo, hi = row_partition(prob.n, world, 256)
        super().__init__(ctx, prob, lo[rank], hi[rank])
