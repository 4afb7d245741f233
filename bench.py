#!/usr/bin/env python
"""bench.py -- tCG iterations/second on Stiefel(1e5,32) (BASELINE.json metric).

A "step" is one whole Steihaug-Toint truncated-CG solve (reference
IterativeSolvers.h:166-426) of the trace-min Hessian system at the
Stiefel(100000, 32) workload `make_stiefel_critical` (block-diagonal A stored in
bf16, fp64 tangent vectors), stopped at relative residual 1e-9 (natural CG
convergence, about 40 iterations).  value = tCG iterations completed / second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torch.distributed.run (one rank per GPU); the tangent
vectors are row-sharded and every rank works on ONE problem (strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "tCG iters/sec on Stiefel(1e5,32)"
UNIT = "iterations/s"
SOLVE = dict(Delta=1e6, max_iterations=200, kappa_fgr=1e-9, theta=0.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c2", action="store_true", help="skip the secondary measurements (sphere C2, LOBPCG C4)")
    return ap.parse_args()


def workload_name(n):
    return (f"Stiefel({n},32) trace-min Hessian tCG, make_stiefel_critical(seed=21), block-diag A bf16 "
            f"(128x128 blocks), fp64 vectors, Delta=1e6 kappa_fgr=1e-9 theta=0 max_iterations=200")


# ----------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(index), "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def samples(self):
        """Sample lines written so far (nvidia-smi takes a few hundred ms to start on a fresh box)."""
        try:
            with open(self.f.name) as fh:
                return sum(1 for line in fh if line.count(",") >= 8)
        except OSError:
            return 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------
# CPU reference arm: the reference's own STPCG header (oracle/_ref) on host cores
# ----------------------------------------------------------------------------------
def host_cores():
    """Host cores this process may use (NOT OMP_NUM_THREADS: torch.distributed.run forces that to 1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference(prob, max_iterations, threads=None, repeats=1):
    """Times the reference CPU path on a bounded sample: ONE solve of the same workload capped at
    `max_iterations` CG iterations on `threads` OpenMP threads (default: every host core, set through the
    stand-in type's own num_threads() clause, so the launcher's OMP_NUM_THREADS does not matter)."""
    from oracle import refapi
    kind = "reference"
    kw = dict(SOLVE, max_iterations=max_iterations)
    try:
        R = refapi.RefOracle()
        threads = host_cores() if threads is None else int(threads)
        R.set_threads(threads)
        rs = R.stiefel(prob)

        def run():
            return rs.stpcg(prob.Y0, prob.g, **kw)
    except (FileNotFoundError, OSError):
        kind = "port"
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
        Pt = refapi.PortOracle()
        threads = 1

        def run():
            return Pt.stpcg_stiefel(prob, prob.Y0, prob.g, **kw)[:3]
    best, out = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = run()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    s_ref, mnorm, its = out[0], out[1], out[2]
    return dict(value=its / best, unit=UNIT, cores=threads, kind=kind, seconds=best, iterations=its, s=s_ref,
                update_step_M_norm=mnorm,
                sample=f"one solve capped at {max_iterations} CG iterations ({its} run) of the same workload, "
                       f"{threads} OpenMP thread(s) of {host_cores()} host cores; HostMat stand-in for Eigen")


def cpu_baseline_rows(prob, full=True):
    """all-core row (a full solve to the natural residual target when `full`: it doubles as the parity
    reference) and the Eigen-faithful 1-thread row (capped at 12 iterations)."""
    allc = cpu_reference(prob, SOLVE["max_iterations"] if full else 12)
    one = cpu_reference(prob, 12, threads=1)
    row = {k: allc[k] for k in ("value", "unit", "cores", "kind", "sample")}
    row["one_thread"] = {k: one[k] for k in ("value", "unit", "cores", "sample")}
    return row, allc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from optimization_b200 import problems as P
    prob = P.make_stiefel_critical(args.n, 32)
    cap = 12
    for _ in range(args.warmup):
        cpu_reference(prob, 2)
    t0 = time.perf_counter()
    its = 0
    info = None
    for _ in range(args.steps):
        info = cpu_reference(prob, cap)
        its += info["iterations"]
    dt = time.perf_counter() - t0
    v = its / dt
    one = cpu_reference(prob, cap, threads=1)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(args.n),
                                            "step": f"one reference STPCG solve capped at {cap} iterations"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                             "sample": info["sample"],
                             "one_thread": {k: one[k] for k in ("value", "unit", "cores", "sample")}},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def sphere_c2(ctx, n=1 << 24, k=16):
    """Secondary workload (BASELINE config C2, not the headline): fused tCG on the sphere Rayleigh Hessian,
    A = diag + rank-16, n = 2^24, generated on the device; same timing rules (CUDA events, warm-up)."""
    import math
    import torch
    from optimization_b200 import problems as P
    try:
        d, Ut, sigma, x0, g = P.make_sphere_critical_device(n, k, device=f"cuda:{ctx.device}")
        H = ctx.sphere_operator(d, None, sigma, x0, Ut=Ut)
        gn = math.sqrt(ctx.dot(g, g))
        kw = dict(Delta=1e6 * gn, max_iterations=200, kappa_fgr=1e-9, theta=0.0)
        s = torch.empty_like(g)
        for _ in range(3):
            ctx.stpcg(g, H, s_out=s, **kw)
        its, kms = 0, 0.0
        for _ in range(5):
            o = ctx.stpcg(g, H, s_out=s, **kw)
            its += o.num_iterations
            kms += o.solve_kernel_ms
        peak, _ = measured_peak()
        sb = H.step_bytes()
        ach = sb * its / kms / 1e6
        return {"workload": f"sphere S^(n-1) Rayleigh-quotient Hessian tCG, n=2^24, A = diag + rank-{k}, fp64 "
                            f"(make_sphere_critical_device), kappa_fgr=1e-9",
                "value": its / (kms * 1e-3), "unit": UNIT, "cg_iterations_per_solve": its / 5,
                "kernel": "tcg_sphere_kernel", "algorithmic_bytes_per_cg_step": sb,
                "achieved_GBps": ach, "frac_of_measured_peak": ach / peak}
    except Exception as e:  # never let the secondary measurement break the headline line
        return {"error": repr(e)}


def lobpcg_c4(ctx, g=160, nx=64, nev=32, iters=10):
    """Secondary workload (BASELINE config C4, not the headline): LOBPCG on the 160^3 7-point Laplacian, block 64,
    Jacobi preconditioner, a fixed number of iterations; time through the public C-ABI call (host-synchronous)."""
    import time
    import torch
    from optimization_b200 import problems as P
    try:
        m = g ** 3
        X0 = (2.0 * P._torch_uniform01(31, 0, m * nx, f"cuda:{ctx.device}") - 1.0).view(m, nx)
        A, T = ctx.block_laplacian3d(g, g, g), ctx.block_scalar(1.0 / 6.0)
        ctx.lobpcg(A, None, T, X0, nev, 3, 1e-6)            # warm-up: workspace, cuSOLVER handle
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th, X, it, nc = ctx.lobpcg(A, None, T, X0, nev, iters, 1e-6)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ns = 3 * nx
        flops = it * (2 * 2 * m * ns * ns + 2 * m * ns * nx)   # two Grams (counted in full) + the fused block update
        return {"workload": f"LOBPCG, 7-point Laplacian {g}^3 (m = {m}), block nx = {nx}, nev = {nev}, Jacobi 1/6, "
                            f"{it} iterations", "value": it / dt, "unit": "iterations/s", "ms_per_iteration": 1e3 * dt / it,
                "bound": "fp64 compute", "achieved_TFLOPs": flops / dt / 1e12,
                "kernels": "blk_gram_kernel / blk_update_kernel (mma.sync.m8n8k4.f64), blk_stencil7_kernel, cusolverDnDsygvd"}
    except Exception as e:  # never let the secondary measurement break the headline line
        return {"error": repr(e)}


def so3_c5(ctx, dims=(100, 100, 100), r=4):
    """Secondary workload (BASELINE config C5, single GPU): fused tCG on the block-CSR 3x3 Hessian of the rotation-
    synchronisation cost, N = 1e6 poses (3-D grid odometry + one random loop closure per pose), rank r = 4,
    consistent measurements (the ground truth is an exact minimiser: positive semidefinite Hessian, right-hand
    side in its range), a fixed number of CG iterations."""
    import torch
    from optimization_b200 import problems as P
    try:
        prob = P.make_posegraph(dims, r, sigma=0.0, x_noise=0.0)
        rp = torch.from_numpy(prob.rowptr.astype(np.int64)).cuda()
        ci = torch.from_numpy(prob.colidx.astype(np.int32)).cuda()
        bl, X = ctx.to_device(prob.blocks), ctx.to_device(prob.X0)
        H = ctx.csr3_operator(rp, ci, bl, X)
        g = ctx.hvp(H, ctx.to_device(prob.g))                    # in the range of H (gauge directions are its kernel)
        gn = float(torch.linalg.norm(g))
        kw = dict(Delta=1e6 * gn, max_iterations=30, kappa_fgr=1e-12, theta=0.0)
        s = torch.empty_like(g)
        for _ in range(2):
            ctx.stpcg(g, H, s_out=s, **kw)
        its, kms = 0, 0.0
        for _ in range(3):
            o = ctx.stpcg(g, H, s_out=s, **kw)
            its += o.num_iterations
            kms += o.solve_kernel_ms
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            ctx.hvp(H, g, out=s)
        ev1.record()
        torch.cuda.synchronize()
        t_hvp = ev0.elapsed_time(ev1) / 5
        peak, _ = measured_peak()
        sb, hb = H.step_bytes(), H.hvp_bytes()
        ach = sb * its / kms / 1e6
        return {"workload": f"SO(3)^N rotation synchronisation (St(3,{r})^N relaxation), N = {prob.N} poses on a "
                            f"{dims[0]}x{dims[1]}x{dims[2]} grid + 1 loop closure per pose, {prob.nnz} 3x3 blocks, "
                            f"tCG capped at 30 iterations", "value": its / (kms * 1e-3), "unit": UNIT,
                "cg_iterations_per_solve": its / 3, "exit_reason": o.exit_reason, "kernel": "tcg_sparse_kernel",
                "algorithmic_bytes_per_cg_step": sb, "achieved_GBps": ach, "frac_of_measured_peak": ach / peak,
                "hvp": {"GBps": hb / t_hvp / 1e6, "ms": t_hvp, "algorithmic_bytes": hb}}
    except Exception as e:  # never let the secondary measurement break the headline line
        return {"error": repr(e)}


def stiefel_fallbacks(ctx, prob, n):
    """The headline kernel needs every block of A to be 22-bit block-fixed-point (the synthetic A is, by
    construction).  Secondary measurements of what other operators get: (i) the same A with the tcgen05 path switched
    off (fp64 tensor-core kernel), (ii) a generic bf16 A (bell-shaped off-diagonals rounded to bf16: exponent spread
    of ~26 bits per block), which the library routes to the fp64 tensor-core kernel by itself."""
    import torch
    from optimization_b200 import problems as P
    from optimization_b200.sharded import SingleStiefel
    out = {}
    try:
        peak, _ = measured_peak()

        def rate(solver):
            for _ in range(2):
                solver.solve_device(**SOLVE)
            its, kms = 0, 0.0
            for _ in range(5):
                o = solver.solve_device(**SOLVE)
                its += o.num_iterations
                kms += o.solve_kernel_ms
            ach = solver.step_bytes_total() * its / kms / 1e6
            return {"value": its / (kms * 1e-3), "unit": UNIT, "cg_iterations_per_solve": its / 5,
                    "kernel_path": ctx.last_path, "achieved_GBps": ach, "frac_of_measured_peak": ach / peak}
        s0 = SingleStiefel(ctx, prob)
        ctx.set_option("tcgen05", 0)
        try:
            out["stiefel_dmma_fallback"] = dict(rate(s0), workload="same workload, ob200_set_option('tcgen05', 0): "
                                                "tcg_stiefel_kernel (A p on the fp64 tensor cores)")
        finally:
            ctx.set_option("tcgen05", 1)
        del s0
        torch.cuda.empty_cache()
        pg = P.make_stiefel_critical(n, 32, generic_bf16=True)
        s1 = SingleStiefel(ctx, pg)
        out["stiefel_generic_bf16_A"] = dict(rate(s1), workload="make_stiefel_critical(generic_bf16=True): off-diagonals "
                                             "= bf16-rounded bell-shaped samples (not block-fixed-point); default options")
    except Exception as e:  # never let the secondary measurement break the headline line
        out["error"] = repr(e)
    return out


# ----------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from optimization_b200 import problems as P
    from optimization_b200.device import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = Context(local)
    n = args.n
    prob = P.make_stiefel_critical(n, 32)
    N = n * 32
    if world > 1:
        from optimization_b200.sharded import ShardedStiefel
        solver = ShardedStiefel(ctx, prob, rank, world)
    else:
        from optimization_b200.sharded import SingleStiefel
        solver = SingleStiefel(ctx, prob)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) -------------------------------------------
    for _ in range(args.warmup):
        solver.solve_device(**SOLVE)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:   # the sampler must be running before the timed region starts; keep the GPU under load while it starts
        t0 = time.time()
        while sampler.p is not None and sampler.samples() == 0 and time.time() - t0 < 5.0:
            if world == 1:
                solver.solve_device(**SOLVE)
            else:
                time.sleep(0.02)     # (a sharded solve is collective: the other ranks are not in this loop)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.kernel_launches
    barrier()
    ev0.record()
    iters = 0
    kernel_ms = 0.0
    for _ in range(args.steps):
        out = solver.solve_device(**SOLVE)
        iters += out.num_iterations
        kernel_ms += out.solve_kernel_ms
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches - launches0
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        kt = torch.tensor([kernel_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
        kernel_ms = float(kt.item())
    value = iters / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers (e2e) -----------------------
    for _ in range(2):
        solver.solve_host(**SOLVE)
    barrier()
    ev0.record()
    it_e2e = 0
    for _ in range(args.steps):
        it_e2e += solver.solve_host(**SOLVE).num_iterations
    ev1.record()
    barrier()
    ms_e2e = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    # ---- stand-alone Hessian-vector product (the metric's "HVP GB/s") ------------------
    hvp = None
    if world == 1:
        for _ in range(3):
            solver.hvp_device()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(10):
            solver.hvp_device()
        ev1.record()
        torch.cuda.synchronize()
        t_hvp = ev0.elapsed_time(ev1) / 10
        hb = solver.H.hvp_bytes()
        hvp = {"value": hb / t_hvp / 1e6, "unit": "GB/s", "algorithmic_bytes": hb, "ms": t_hvp,
               "what": "ob200_hvp (stand-alone Hess f(Y)[V]): one prologue launch (exact <V,V>, content checksum of A, clears) + ONE "
                       "persistent launch of the fused kernel in HVP mode (contraction + Gram | projection), no host round trip "
                       "before the final status read; "
                       "inside the fused tCG step the HVP never runs stand-alone"}
    if sampler:   # a short default run may end between two 50 ms samples: stay under the same load until three are in
        t0 = time.time()
        while world == 1 and sampler.p is not None and sampler.samples() < 3 and time.time() - t0 < 3.0:
            solver.solve_device(**SOLVE)
    clocks = sampler.stop() if sampler else None
    # ---- parity inside the bench: the solve just timed against the reference's own STPCG on the host, and
    #      (N > 1) bit identity of the row-sharded solve with a one-GPU solve of the whole problem -----------------
    out = solver.solve_device(**SOLVE)
    s_loc = out.s.cpu().numpy()
    parity = None
    cpu_rows = None
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (solver.lo, s_loc))
        s_full = np.concatenate([p_[1] for p_ in sorted(parts, key=lambda t: t[0])], axis=0) if rank == 0 else None
    else:
        s_full = s_loc
    if rank == 0:
        parity = {"num_iterations": int(out.num_iterations), "exit_reason": out.exit_reason}
        if world > 1:
            c1 = Context(local)                          # independent one-GPU context, whole problem
            from optimization_b200.sharded import SingleStiefel
            o1 = SingleStiefel(c1, prob).solve_device(**SOLVE)
            s1 = o1.s.cpu().numpy()
            parity["vs_one_gpu"] = {"bit_identical": bool(np.array_equal(s1, s_full)
                                                          and o1.num_iterations == out.num_iterations
                                                          and o1.exit_reason == out.exit_reason
                                                          and o1.update_step_M_norm == out.update_step_M_norm),
                                    "kernel_paths": [ctx.last_path, c1.last_path],
                                    "max_abs_diff": float(np.abs(s1 - s_full).max()),
                                    "num_iterations_one_gpu": int(o1.num_iterations)}
            c1.close()
        if not args.no_cpu_baseline:
            cpu_rows, ref = cpu_baseline_rows(prob, full=True)
            s_ref = ref["s"]
            parity["vs_reference"] = {
                "oracle": f"oracle/_ref (the reference's own IterativeSolvers.h STPCG, {ref['cores']} threads)"
                          if ref["kind"] == "reference" else "oracle/stpcg_port.c (C restatement)",
                "num_iterations_reference": int(ref["iterations"]),
                "iterations_equal": bool(int(ref["iterations"]) == int(out.num_iterations)),
                "rel_err_s": float(np.linalg.norm(s_full - s_ref) / np.linalg.norm(s_ref)),
                "rel_err_update_step_M_norm": float(abs(out.update_step_M_norm - ref["update_step_M_norm"])
                                                    / abs(ref["update_step_M_norm"])),
                "tolerance": 1e-10}
            parity["vs_reference"]["ok"] = bool(parity["vs_reference"]["iterations_equal"]
                                                and parity["vs_reference"]["rel_err_s"] < 1e-10)
    c2 = sphere_c2(ctx) if (world == 1 and not args.no_c2) else None
    c4 = lobpcg_c4(ctx) if (world == 1 and not args.no_c2) else None
    fb = stiefel_fallbacks(ctx, prob, n) if (world == 1 and not args.no_c2) else None
    c5 = so3_c5(ctx) if (world == 1 and not args.no_c2) else None

    if rank == 0:
        peak, peak_src = measured_peak()
        step_bytes = solver.step_bytes_total()           # algorithmic bytes of one CG step, whole problem
        per_launch_bytes = step_bytes * (iters / args.steps) / world     # per GPU, per persistent-kernel launch
        per_launch_s = (kernel_ms / args.steps) * 1e-3
        achieved = per_launch_bytes / per_launch_s / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n), "cg_iterations_per_step": iters / args.steps,
                           "l2": "working set 7 x 25.6 MB vectors + 25.6 MB A per solve exceeds the 126 MB L2; "
                                 "no explicit flush", "parallelism": f"row-shard x{world}"},
                "roofline": {"bound": "hbm", "kernel": solver.kernel_name, "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                             "traffic": solver.ncu_traffic_per_launch(iters / args.steps),
                             "algorithmic_bytes_per_cg_step": step_bytes,
                             "kernel_ms_per_launch": kernel_ms / args.steps,
                             "kernel_share_of_step": kernel_ms / ms},
                "e2e": {"value": it_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * N,
                        "d2h_bytes_per_step": 8 * N, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks}
        if hvp:
            line["hvp"] = hvp
        if c2:
            line["other_workloads"] = {"sphere_c2": c2}
            if c4:
                line["other_workloads"]["lobpcg_c4"] = c4
            if fb:
                line["other_workloads"].update(fb)
            if c5:
                line["other_workloads"]["so3_c5"] = c5
        if parity:
            line["parity"] = parity
        if cpu_rows:
            line["cpu_baseline"] = cpu_rows
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
