"""CPU (gloo, world_size 2): host-side logic of the multi-GPU path -- row
partition, IPC-handle exchange plumbing, and the property the design rests on:
exact integer accumulators summed across ranks equal the single-rank result bit
for bit (the kernels exchange exactly these integer words over NVLink)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optimization_b200.sharded import row_partition

KUL_BIAS, KUL_LIMBS = 1088, 68


def kulisch_accumulate(x: np.ndarray) -> np.ndarray:
    """Python twin of csrc/common.cuh:kul_decompose / kul_add_host (test only)."""
    acc = np.zeros(KUL_LIMBS + 4, dtype=object)
    for v in x.tolist():
        if v == 0.0:
            continue
        m, e = np.frexp(v)                    # v = m * 2^e, 0.5 <= |m| < 1
        mant = int(m * (1 << 53))             # exact
        shift = int(e) - 53 + KUL_BIAS
        j, off = shift >> 5, shift & 31
        mag = abs(mant) << off
        sign = -1 if mant < 0 else 1
        acc[j] += sign * (mag & 0xFFFFFFFF)
        acc[j + 1] += sign * ((mag >> 32) & 0xFFFFFFFF)
        acc[j + 2] += sign * (mag >> 64)
    return np.array([int(a) for a in acc[:KUL_LIMBS]], dtype=np.int64)


def kulisch_value(limbs: np.ndarray):
    from fractions import Fraction
    tot = sum(int(l) << (32 * j) for j, l in enumerate(limbs.tolist()))
    return float(Fraction(tot, 1 << KUL_BIAS))    # correctly rounded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = row_partition(n, world, 128)
        rng = np.random.default_rng(7)
        a = rng.standard_normal((n, 4)) * np.exp(rng.uniform(-30, 30, (n, 4)))
        local = kulisch_accumulate((a[lo[rank]:hi[rank]] ** 2).ravel())
        t = torch.from_numpy(local.copy())
        dist.all_reduce(t)                                   # what the kernels do over NVLink
        # handle exchange plumbing used by Context.connect
        mine = bytes([rank]) * 64
        got = [None] * world
        dist.all_gather_object(got, mine)
        out_q.put((rank, t.numpy().tolist(), [g[0] for g in got], lo, hi))
    finally:
        dist.destroy_process_group()


def test_row_partition_covers_and_aligns():
    for n in (1, 127, 128, 129, 1000, 100000, 12345):
        for world in (1, 2, 3, 4, 8):
            lo, hi = row_partition(n, world, 128)
            assert lo[0] == 0 and hi[-1] == n
            for r in range(world):
                assert lo[r] <= hi[r] and lo[r] % 128 == 0
                if r + 1 < world:
                    assert hi[r] == lo[r + 1]
    # sphere shards: boundaries at multiples of the 256-element reduction unit
    for n in (1, 255, 257, 100003, 1 << 20):
        for world in (1, 2, 8):
            lo, hi = row_partition(n, world, 256)
            assert lo[0] == 0 and hi[-1] == n and all(l % 256 == 0 for l in lo)
            assert all(hi[r] == lo[r + 1] for r in range(world - 1))
    lo, hi = row_partition(100000, 8, 128)
    assert max(h - l for l, h in zip(lo, hi)) - min(h - l for l, h in zip(lo, hi)) <= 128 + 96


@pytest.mark.timeout(180)
def test_two_rank_exact_reduction_matches_single_rank():
    world, n = 2, 700
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    rng = np.random.default_rng(7)
    a = rng.standard_normal((n, 4)) * np.exp(rng.uniform(-30, 30, (n, 4)))
    single = kulisch_accumulate((a ** 2).ravel())
    import math
    want = math.fsum((a ** 2).ravel().tolist())
    for rank, limbs, handles, lo, hi in res:
        # integer limb sums need not be limb-wise equal (carries are not normalised), the VALUE is
        assert kulisch_value(np.array(limbs, dtype=np.int64)) == kulisch_value(single) == want
        assert handles == [0, 1]


def test_posegraph_sharding_plan_is_consistent():
    """Host-side plan of the row-sharded pose-graph operator (config C5): local column indices address own poses then halo
    poses, entry order inside a row is unchanged, every halo pose is pushed by its owner into the right slot."""
    import numpy as np
    from optimization_b200 import problems as P
    from optimization_b200.sharded import halo_send_plan, pose_partition, shard_posegraph
    prob = P.make_posegraph((16, 16, 8), 4)
    for world in (1, 2, 3, 8):
        lo, hi = pose_partition(prob.N, world)
        assert lo[0] == 0 and hi[-1] == prob.N and all(a % 256 == 0 for a in lo) and lo[1:] == hi[:-1]
        plans = [shard_posegraph(prob, q, world) for q in range(world)]
        sends = [halo_send_plan(prob, q, world) for q in range(world)]
        for q, pl in enumerate(plans):
            n_loc = pl["hi"] - pl["lo"]
            e0 = int(prob.rowptr[pl["lo"]])
            glob = prob.colidx[e0:e0 + pl["colidx"].size].astype(np.int64)
            back = np.where(pl["colidx"] < n_loc, pl["colidx"].astype(np.int64) + pl["lo"],
                            pl["halo"][np.maximum(pl["colidx"].astype(np.int64) - n_loc, 0)] if pl["halo"].size else 0)
            assert np.array_equal(back, glob)                         # same entries, same order
            assert np.all((pl["halo"] < pl["lo"]) | (pl["halo"] >= pl["hi"]))
            # emulate the pushes of every owner into q's halo buffer
            got = np.full(pl["halo"].size, -1, dtype=np.int64)
            for src in range(world):
                idx, ptr, off = sends[src]
                mine = idx[int(ptr[q]):int(ptr[q + 1])].astype(np.int64) + plans[src]["lo"]
                got[int(off[q]):int(off[q]) + mine.size] = mine
            assert np.array_equal(got, pl["halo"])
