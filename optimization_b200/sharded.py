"""Workload drivers used by bench.py and the multi-GPU tests: one Stiefel tCG
problem resident on one GPU (`SingleStiefel`) or row-sharded over the ranks of
a torch.distributed job (`ShardedStiefel`).  Host-side plumbing only."""
from __future__ import annotations

import numpy as np
import torch


def _bits(A_bf16: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(A_bf16).view(np.int16)).to(device)


class SingleStiefel:
    kernel_name = "tcg_stiefel_v6_kernel (persistent fused tCG, whole solve; warp-specialised roles, tcgen05 int8 digit planes, TMA-staged operands, solve in the eigenbasis of S)"

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
                rec = json.load(fh)["tcg_stiefel_v6_kernel"]
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
    kernel_name = "tcg_stiefel_v6_kernel<MULTI> (persistent fused tCG + in-kernel NVLink peer exchange)"

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


# ---- BASELINE config C5: the pose-graph Hessian row-sharded by pose ranges, halo exchange of p over NVLink ------------
def pose_partition(N: int, world: int, unit: int = 256):
    """Contiguous pose ranges, boundaries at multiples of `unit` (the 256-element runs and the 8-pose groups of the
    exact reductions never straddle a shard): rank q owns poses [lo[q], hi[q])."""
    nunits = (N + unit - 1) // unit
    lo = [min(N, unit * (nunits * q // world)) for q in range(world)]
    hi = [min(N, unit * (nunits * (q + 1) // world)) for q in range(world)]
    return lo, hi


def shard_posegraph(prob, rank: int, world: int, unit: int = 256):
    """Host-side sharding plan of rank `rank` (pure numpy; every rank can compute every rank's plan):
    local block-CSR with LOCAL column indices ([0, n_local) own poses, then the halo poses in ascending global order,
    entry order inside a row unchanged), the halo pose list grouped by owner, and what this rank must send to whom."""
    lo, hi = pose_partition(prob.N, world, unit)
    l0, l1 = lo[rank], hi[rank]
    e0, e1 = int(prob.rowptr[l0]), int(prob.rowptr[l1])
    cols = prob.colidx[e0:e1].astype(np.int64)
    remote = (cols < l0) | (cols >= l1)
    halo = np.unique(cols[remote])                                   # ascending global pose index
    local_cols = np.where(remote, (l1 - l0) + np.searchsorted(halo, cols), cols - l0)
    owner = np.searchsorted(np.asarray(hi), halo, side="right")      # rank owning each halo pose
    halo_ptr = np.searchsorted(owner, np.arange(world + 1))          # halo[halo_ptr[q]:halo_ptr[q+1]] owned by q
    return dict(lo=l0, hi=l1, rowptr=(prob.rowptr[l0:l1 + 1] - prob.rowptr[l0]).astype(np.uint64),
                colidx=local_cols.astype(np.uint32), blocks=np.ascontiguousarray(prob.blocks[e0:e1]),
                halo=halo, halo_ptr=halo_ptr, lo_all=lo, hi_all=hi)


def halo_send_plan(prob, rank: int, world: int, unit: int = 256):
    """What `rank` pushes: for every destination q the own poses q needs (as local indices, in the order of q's halo
    list) and the slot in q's halo buffer where they start."""
    send_idx, send_ptr, dst_off = [], [0], []
    for q in range(world):
        if q == rank:
            send_ptr.append(send_ptr[-1])
            dst_off.append(0)
            continue
        plan_q = shard_posegraph(prob, q, world, unit)
        a, b = plan_q["halo_ptr"][rank], plan_q["halo_ptr"][rank + 1]
        mine = plan_q["halo"][a:b]                                   # my poses that q needs
        send_idx.append(mine - plan_q["lo_all"][rank])
        send_ptr.append(send_ptr[-1] + mine.size)
        dst_off.append(int(a))
    idx = np.concatenate(send_idx).astype(np.uint32) if send_idx else np.zeros(0, np.uint32)
    return idx, np.asarray(send_ptr, dtype=np.uint64), np.asarray(dst_off, dtype=np.uint64)


class SinglePoseGraph:
    def __init__(self, ctx, prob, g=None):
        self.ctx, self.prob = ctx, prob
        self.rp = torch.from_numpy(prob.rowptr.astype(np.int64)).cuda()
        self.ci = torch.from_numpy(prob.colidx.astype(np.int32)).cuda()
        self.bl, self.X = ctx.to_device(prob.blocks), ctx.to_device(prob.X0)
        self.H = ctx.csr3_operator(self.rp, self.ci, self.bl, self.X)
        self.g = ctx.to_device(prob.g if g is None else g)

    def solve_device(self, **kw):
        return self.ctx.stpcg(self.g, self.H, **kw)


class ShardedPoseGraph:
    """One pose-graph tCG problem row-sharded over the ranks of a torch.distributed job.  Every rank holds its poses'
    rows of Q, X for own + halo poses, and a halo buffer that the owners of the halo poses fill (peer stores over
    NVLink, inside the persistent kernel) before every operator apply."""

    def __init__(self, ctx, prob, rank, world, g=None):
        import ctypes as C
        import torch.distributed as dist
        from . import capi
        self.ctx, self.prob, self.rank, self.world = ctx, prob, rank, world
        if getattr(ctx, "world", 1) != world:
            ctx.connect(rank, world)
        plan = shard_posegraph(prob, rank, world)
        self.lo, self.hi = plan["lo"], plan["hi"]
        n_loc, n_halo, r = self.hi - self.lo, int(plan["halo"].size), prob.r
        # halo buffer + handle exchange
        buf = (C.c_ubyte * capi.COMM_HANDLE_BYTES)()
        ctx._check(ctx.lib.ob200_halo_create(ctx.h, max(n_halo, 1) * 3 * r * 8, buf))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(buf))
        ctx._check(ctx.lib.ob200_halo_connect(ctx.h, b"".join(handles)))
        Xb = prob.X0.reshape(prob.N, 3 * r)
        Xext = np.concatenate([Xb[self.lo:self.hi], Xb[plan["halo"]]], axis=0).reshape(-1, r)
        self.rp = torch.from_numpy(plan["rowptr"].astype(np.int64)).cuda()
        self.ci = torch.from_numpy(plan["colidx"].astype(np.int32)).cuda()
        self.bl, self.Xext = ctx.to_device(plan["blocks"]), ctx.to_device(np.ascontiguousarray(Xext))
        idx, ptr, off = halo_send_plan(prob, rank, world)
        self.send_idx = torch.from_numpy(idx.astype(np.int32)).cuda() if idx.size else torch.zeros(1, dtype=torch.int32).cuda()
        # Lambda of the own poses (X of the halo poses is part of Xext), f summed over the ranks
        N_loc = n_loc
        lam = torch.empty((N_loc, 9), dtype=torch.float64, device=self.Xext.device)
        f = C.c_double(0)
        ctx._check(ctx.lib.ob200_csr3_model(ctx.h, N_loc, r, self.rp.data_ptr(), self.ci.data_ptr(), self.bl.data_ptr(),
                                            self.Xext.data_ptr(), lam.data_ptr(), C.byref(f), None))
        op = capi.Operator()
        op.kind = capi.OP_BLOCK_CSR3
        op.n, op.p = 3 * N_loc, r
        op.Y_dev = self.Xext.data_ptr()
        op.csr_rowptr_dev, op.csr_colidx_dev = self.rp.data_ptr(), self.ci.data_ptr()
        op.csr_blocks_dev, op.csr_lambda_dev = self.bl.data_ptr(), lam.data_ptr()
        op.csr_nnz = int(self.ci.numel())
        op.csr_n_halo = n_halo
        op.halo_send_idx_dev = self.send_idx.data_ptr()
        for q in range(world + 1):
            op.halo_send_ptr[q] = int(ptr[q])
        for q in range(world):
            op.halo_dst_off[q] = int(off[q])
        from .device import OperatorHandle
        self.H = OperatorHandle(ctx, op, [self.rp, self.ci, self.bl, self.Xext, lam, self.send_idx])
        self.f, self.Lambda = f.value, lam
        gg = prob.g if g is None else g
        self.g = ctx.to_device(np.ascontiguousarray(gg[3 * self.lo:3 * self.hi]))

    def solve_device(self, **kw):
        return self.ctx.stpcg(self.g, self.H, **kw)


# ---- BASELINE config C4 sharded: LOBPCG with block vectors row-sharded over the ranks -----------------------------------
def lobpcg_halo_bytes(world: int, nx: int, plane_rows: int = 0) -> int:
    """Size of the halo buffer ob200_lobpcg needs in row-sharded mode (same formula as csrc/capi.cu)."""
    nsmax = 3 * nx
    return 8 * (512 + 2 * world * (nsmax * nsmax + 64) + 4 * plane_rows * nsmax)


def setup_halo(ctx, world: int, nbytes: int):
    """Allocate this rank's halo buffer and connect to the peers' (handles exchanged with torch.distributed)."""
    import ctypes as C
    import torch.distributed as dist
    from . import capi
    buf = (C.c_ubyte * capi.COMM_HANDLE_BYTES)()
    ctx._check(ctx.lib.ob200_halo_create(ctx.h, int(nbytes), buf))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(buf))
    ctx._check(ctx.lib.ob200_halo_connect(ctx.h, b"".join(handles)))


def z_partition(gz: int, world: int):
    """z-planes [lo[q], hi[q]) of rank q (contiguous slabs of the gx x gy x gz grid, x fastest)."""
    lo = [gz * q // world for q in range(world)]
    hi = [gz * (q + 1) // world for q in range(world)]
    return lo, hi


class ShardedLaplacianLobpcg:
    """LOBPCG on the 7-point Laplacian of a gx x gy x gz grid with the block vectors sharded in z-slabs: Grams, norms and
    residual norms all-reduced in rank order, Rayleigh-Ritz replicated, one ghost plane exchanged with each neighbour
    before every operator apply -- all through NVLink peer stores (no NCCL call on the data path)."""

    def __init__(self, ctx, gx, gy, gz, nx, rank, world):
        self.ctx, self.rank, self.world = ctx, rank, world
        if getattr(ctx, "world", 1) != world:
            ctx.connect(rank, world)
        lo, hi = z_partition(gz, world)
        self.z0, self.z1 = lo[rank], hi[rank]
        self.rows = slice(self.z0 * gx * gy, self.z1 * gx * gy)
        setup_halo(ctx, world, lobpcg_halo_bytes(world, nx, gx * gy))
        self.A = ctx.block_laplacian3d(gx, gy, self.z1 - self.z0)          # LOCAL planes
        self.T = ctx.block_scalar(1.0 / 6.0)

    def solve(self, X0_local, nev, max_iters, tau, Omega_local):
        return self.ctx.lobpcg(self.A, None, self.T, X0_local, nev, max_iters, tau, Omega=Omega_local)
