"""Python host-side handle over the C ABI: contexts, operators, tCG solves.

torch owns device memory and the stream (plumbing only); every arithmetic
statement of the hot path executes inside liboptimization_b200.so.  Names and
argument meaning follow the reference (`STPCG(g, H, inner_product, ...,
Delta, max_iterations, kappa_fgr, theta, P, ..., epsilon)`,
IterativeSolvers.h:166-179); bad parameters raise ValueError where the
reference throws std::invalid_argument (IterativeSolvers.h:183-205).
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np
import torch

from . import capi


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous()
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        assert t.flags.c_contiguous
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(int(t))


@dataclasses.dataclass
class StpcgOutput:
    s: object
    update_step_M_norm: float
    num_iterations: int
    exit_reason: str
    r0_norm: float
    final_rv: float
    kernel_launches: int
    solve_kernel_ms: float = 0.0


class Context:
    """One per GPU / process.  Work is issued on torch's current stream of `device`."""

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("optimization_b200 needs a CUDA device (no CPU fallback)")
        self.lib = capi.lib()
        self.device = int(device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.current_stream(self.device)
        # torch's default stream is the legacy default stream, whose handle is 0; the C ABI reads NULL as "create a
        # private non-blocking stream" (no ordering with torch's work), so stream 0 is passed as cudaStreamLegacy (0x1).
        handle = self.stream.cuda_stream or 0x1
        h = C.c_void_p()
        rc = self.lib.ob200_create(self.device, C.c_void_p(handle), C.byref(h))
        if rc != capi.OK:
            raise capi.Ob200Error(rc, "ob200_create failed (no usable GPU?)")
        self.h = h
        self._keep = []
        import os
        if os.environ.get("OB200_TCGEN05"):      # developer override: which Stiefel kernel generation to run (see set_option)
            self.set_option("tcgen05", int(os.environ["OB200_TCGEN05"]))

    def close(self):
        if getattr(self, "h", None):
            self.lib.ob200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- multi-GPU ---------------------------------------------------------------
    def connect(self, rank: int, world: int, all_gather_bytes=None):
        """Join the NVLink peer-memory exchange of a `world`-rank job (one process
        per GPU).  `all_gather_bytes(b) -> [bytes per rank]` exchanges the 64-byte
        IPC handles; default: torch.distributed.all_gather_object."""
        buf = (C.c_ubyte * capi.COMM_HANDLE_BYTES)()
        self._check(self.lib.ob200_comm_export(self.h, buf))
        mine = bytes(buf)
        if all_gather_bytes is None:
            import torch.distributed as dist

            def all_gather_bytes(b):
                out = [None] * world
                dist.all_gather_object(out, b)
                return out
        handles = all_gather_bytes(mine)
        assert len(handles) == world and handles[rank] == mine
        blob = b"".join(handles)
        self._check(self.lib.ob200_comm_connect(self.h, int(rank), int(world), blob))
        self.rank, self.world = int(rank), int(world)

    # -- helpers ---------------------------------------------------------------
    def _check(self, rc):
        if rc == capi.OK:
            return
        msg = self.lib.ob200_last_error(self.h).decode()
        if rc == capi.INVALID_ARGUMENT:
            raise ValueError(msg)          # reference: std::invalid_argument
        raise capi.Ob200Error(rc, msg)

    def to_device(self, a: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{self.device}")

    @property
    def sm_count(self):
        return self.lib.ob200_sm_count(self.h)

    @property
    def kernel_launches(self):
        return int(self.lib.ob200_kernel_launches(self.h))

    def set_option(self, name: str, value: int):
        self._check(self.lib.ob200_set_option(self.h, name.encode(), int(value)))

    @property
    def last_path(self):
        return {1: "tcgen05", 2: "tcgen05_v4", 0: "dmma", 3: "generic"}.get(self.lib.ob200_last_path(self.h), "?")

    def synchronize(self):
        self._check(self.lib.ob200_synchronize(self.h))

    # -- operators ---------------------------------------------------------------
    def diag_operator(self, d: torch.Tensor) -> "OperatorHandle":
        op = capi.Operator()
        op.kind = capi.OP_DIAG
        op.n = d.numel()
        op.p = 1
        op.diag_dev = d.data_ptr()
        return OperatorHandle(self, op, [d])

    def callback_operator(self, n: int, p: int, fn) -> "OperatorHandle":
        """OB200_OP_HOST_CALLBACK: `fn(v, out)` receives torch views (n x p, float64, this device) of the library's
        device buffers and must fill `out` with H v (unfused fallback: ob200_stpcg runs the reference loop on the host
        over the device level-1 kernels)."""
        cb, keep = self._callback(n, p, fn)
        op = capi.Operator()
        op.kind = capi.OP_HOST_CALLBACK
        op.n, op.p = n, p
        op.apply = C.cast(cb, C.c_void_p).value
        op.apply_user = None
        return OperatorHandle(self, op, [cb, keep])

    def callback_precon(self, n: int, p: int, fn):
        """OB200_PRECON_HOST_CALLBACK: v = fn(r, out) (to be used with a callback operator)."""
        cb, keep = self._callback(n, p, fn)
        pc = capi.Precon()
        pc.kind = capi.PRECON_HOST_CALLBACK
        pc.minv_dev = None
        pc.apply = C.cast(cb, C.c_void_p).value
        pc.apply_user = None
        pc._keep = (cb, keep)
        return pc

    def _callback(self, n, p, fn):
        dev = self.device

        class _View:          # a device pointer as a CUDA array (zero-copy torch view)
            def __init__(self, ptr, ro):
                self.__cuda_array_interface__ = {"shape": (n, p), "typestr": "<f8", "data": (ptr, ro), "version": 2}

        def trampoline(user, in_ptr, out_ptr):
            try:
                with torch.cuda.device(dev):
                    v = torch.as_tensor(_View(in_ptr, False), device=f"cuda:{dev}")   # (torch has no read-only views)
                    out = torch.as_tensor(_View(out_ptr, False), device=f"cuda:{dev}")
                    fn(v, out)
                    torch.cuda.synchronize(dev)      # contract: complete (or stream-ordered) on return
                return 0
            except Exception:   # never let a Python exception cross the C boundary
                import traceback
                traceback.print_exc()
                return 1

        cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)(trampoline)
        return cb, trampoline

    def stiefel_operator(self, A_bf16: torch.Tensor, Y: torch.Tensor) -> "OperatorHandle":
        """Hess f(Y)[V] = P_Y(A V - V sym(Y^T A Y)); computes S on the device."""
        n, p = Y.shape
        S = np.zeros((p, p))
        f = C.c_double(0)
        bound = C.c_double(0)
        self._check(self.lib.ob200_stiefel_model(self.h, n, p, _ptr(A_bf16), _ptr(Y), _ptr(S),
                                                 C.byref(f), None, C.byref(bound)))
        op = capi.Operator()
        op.kind = capi.OP_STIEFEL_BLOCKDIAG
        op.n, op.p = n, p
        op.A_bf16_dev = A_bf16.data_ptr()
        op.Y_dev = Y.data_ptr()
        op.S_host = S.ctypes.data
        op.op_norm_bound = bound.value
        h = OperatorHandle(self, op, [A_bf16, Y, S])
        h.S, h.f = S, f.value
        return h

    def sphere_operator(self, d: torch.Tensor, U: torch.Tensor, sigma: np.ndarray, x: torch.Tensor,
                        Ut: torch.Tensor | None = None) -> "OperatorHandle":
        """Hess f(x)[v] = 2 P_x(A v) - 2 (x^T A x) v for f(x) = x^T A x on the sphere,
        A = diag(d) + U diag(sigma) U^T (U: n x k).  The device keeps U transposed (k x n)."""
        n = d.numel()
        k = int(sigma.size)
        if Ut is None:                       # k rows of ldu = n rounded up to even (16-byte aligned rows)
            Ut = torch.zeros((k, n + (n & 1)), dtype=torch.float64, device=d.device)
            if k:
                Ut[:, :n] = U.t()
        sigma = np.ascontiguousarray(sigma, dtype=np.float64)
        Ax, f, _ = self.sphere_model(d, Ut, sigma, x, want_grad=False)
        op = capi.Operator()
        op.kind = capi.OP_SPHERE_LOWRANK
        op.n, op.p, op.k = n, 1, k
        op.diag_dev = d.data_ptr()
        op.U_dev = Ut.data_ptr() if k else None
        op.sigma_host = sigma.ctypes.data if k else None
        op.ldu = Ut.shape[1] if k else 0
        op.x_dev = x.data_ptr()
        op.Ax_dev = Ax.data_ptr()
        op.xAx = f
        h = OperatorHandle(self, op, [d, Ut, sigma, x, Ax])
        h.f, h.Ut, h.Ax = f, Ut, Ax
        return h

    def csr3_model(self, rowptr, colidx, blocks, X, want_grad=True):
        """(Lambda [N x 9], f = tr(X^T Q X), Riemannian gradient) of the rotation-synchronisation cost on the device."""
        N, r = X.shape[0] // 3, X.shape[1]
        lam = torch.empty((N, 9), dtype=torch.float64, device=X.device)
        grad = torch.empty_like(X) if want_grad else None
        f = C.c_double(0)
        self._check(self.lib.ob200_csr3_model(self.h, N, r, _ptr(rowptr), _ptr(colidx), _ptr(blocks), _ptr(X), _ptr(lam),
                                              C.byref(f), _ptr(grad)))
        return lam, f.value, grad

    def csr3_retract(self, X, V, out=None):
        if out is None:
            out = torch.empty_like(X)
        self._check(self.lib.ob200_csr3_retract(self.h, X.shape[0] // 3, X.shape[1], _ptr(X), _ptr(V), _ptr(out)))
        return out

    def csr3_operator(self, rowptr, colidx, blocks, X) -> "OperatorHandle":
        """Hess f(X)[V] = Proj_X(2 Q V - Lambda V) on St(3, r)^N (BASELINE config C5).  rowptr: int64/uint64 [N + 1],
        colidx: int32/uint32 [nnz], blocks: float64 [nnz, 9], X: float64 [3N, r] (all on the device)."""
        lam, f, _ = self.csr3_model(rowptr, colidx, blocks, X, want_grad=False)
        op = capi.Operator()
        op.kind = capi.OP_BLOCK_CSR3
        op.n, op.p = X.shape
        op.Y_dev = X.data_ptr()
        op.csr_rowptr_dev, op.csr_colidx_dev = rowptr.data_ptr(), colidx.data_ptr()
        op.csr_blocks_dev, op.csr_lambda_dev = blocks.data_ptr(), lam.data_ptr()
        op.csr_nnz = int(colidx.numel())
        h = OperatorHandle(self, op, [rowptr, colidx, blocks, X, lam])
        h.f, h.Lambda, h.nnz = f, lam, int(colidx.numel())
        return h

    def stencil7_operator(self, gx: int, gy: int, gz: int, p: int) -> "OperatorHandle":
        """H V = 7-point Dirichlet Laplacian on the gx x gy x gz grid applied to the p columns of V (n = gx gy gz)."""
        op = capi.Operator()
        op.kind = capi.OP_STENCIL7
        op.n, op.p = gx * gy * gz, p
        op.gx, op.gy, op.gz = gx, gy, gz
        return OperatorHandle(self, op, [])

    def sphere_model(self, d, Ut, sigma, x, want_grad=True):
        """(A x, f = x^T A x, grad = 2 (A x - f x)) on the device."""
        n = d.numel()
        k = int(sigma.size)
        Ax = torch.empty_like(x)
        grad = torch.empty_like(x) if want_grad else None
        f = C.c_double(0)
        self._check(self.lib.ob200_sphere_model(self.h, n, k, _ptr(d), _ptr(Ut) if k else None,
                                                Ut.shape[1] if k else 0, _ptr(sigma) if k else None,
                                                _ptr(x), _ptr(Ax), C.byref(f), _ptr(grad)))
        return Ax, f.value, grad

    def sphere_retract(self, x, v, out=None):
        if out is None:
            out = torch.empty_like(x)
        self._check(self.lib.ob200_sphere_retract(self.h, x.numel(), _ptr(x), _ptr(v), _ptr(out)))
        return out

    # -- LOBPCG (reference LinearAlgebra/LOBPCG.h:131-337) -----------------------------------------
    @staticmethod
    def block_diag(d: torch.Tensor):
        op = capi.BlockOperator()
        op.kind, op.diag_dev = capi.BLK_DIAG, d.data_ptr()
        op._keep = d
        return op

    @staticmethod
    def block_scalar(alpha: float):
        op = capi.BlockOperator()
        op.kind, op.alpha = capi.BLK_SCALAR, float(alpha)
        return op

    @staticmethod
    def block_laplacian3d(gx: int, gy: int, gz: int):
        op = capi.BlockOperator()
        op.kind, op.gx, op.gy, op.gz = capi.BLK_STENCIL7, gx, gy, gz
        return op

    def block_apply(self, op, X: torch.Tensor):
        out = torch.empty_like(X)
        m, k = X.shape
        self._check(self.lib.ob200_block_apply(self.h, C.byref(op), m, k, _ptr(X), k, _ptr(out), k))
        return out

    def lobpcg(self, A, B, T, X0: torch.Tensor, nev: int, max_iters: int, tau: float = 1e-6, Omega=None):
        """Smallest nev eigenpairs of A x = lambda B x.  A, B, T: block operator descriptors (B, T may be None);
        X0: m x nx.  Returns (theta[nev], X[m x nev], num_iters, num_converged); ValueError where the reference
        throws std::invalid_argument (LOBPCG.h:148-155)."""
        X = X0.clone().contiguous()
        m, nx = X.shape
        theta = np.zeros(nev)
        it, nc = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.ob200_lobpcg(self.h, C.byref(A), C.byref(B) if B is not None else None,
                                   C.byref(T) if T is not None else None, m, nx, _ptr(X), nev, max_iters, float(tau),
                                   _ptr(Omega), theta.ctypes.data_as(C.POINTER(C.c_double)), C.byref(it), C.byref(nc))
        self._check(rc)
        return theta, X[:, :nev], int(it.value), int(nc.value)

    def jacobi(self, minv: torch.Tensor | None):
        pc = capi.Precon()
        if minv is None:
            pc.kind = capi.PRECON_NONE
            pc.minv_dev = None
        else:
            pc.kind = capi.PRECON_JACOBI
            pc.minv_dev = minv.data_ptr()
        return pc

    def projected_jacobi(self, minv: torch.Tensor):
        """OB200_PRECON_STIEFEL_PROJECTED_JACOBI: v = P_Y(minv o r) for the Stiefel operator (Y is the operator's point):
        the tangent-space preserving form of the Jacobi scaling; selects the unfused loop with the one-launch HVP."""
        pc = capi.Precon()
        pc.kind = capi.PRECON_STIEFEL_PROJECTED_JACOBI
        pc.minv_dev = minv.data_ptr()
        pc._keep = minv
        return pc

    # -- the hot path --------------------------------------------------------------
    def stpcg(self, g, H: "OperatorHandle", Delta, max_iterations=1000, kappa_fgr=0.1, theta=0.5,
              minv=None, epsilon=1e-8, s_out=None, host=False) -> StpcgOutput:
        """Steihaug-Toint truncated preconditioned CG (IterativeSolvers.h:166-426).

        g: torch cuda tensor (device entry) or, with host=True, a numpy array /
        pinned torch CPU tensor (copies are part of the call)."""
        prm = capi.StpcgParams(float(Delta), int(max_iterations), float(kappa_fgr), float(theta),
                               float(epsilon))
        res = capi.StpcgResult()
        pc = minv if isinstance(minv, capi.Precon) else self.jacobi(minv)
        if host:
            if s_out is None:
                s_out = np.empty_like(g) if isinstance(g, np.ndarray) else torch.empty_like(g)
            rc = self.lib.ob200_stpcg_host(self.h, C.byref(H.op), C.byref(pc), _ptr(g),
                                           C.byref(prm), _ptr(s_out), C.byref(res))
        else:
            if s_out is None:
                s_out = torch.empty_like(g)
            rc = self.lib.ob200_stpcg(self.h, C.byref(H.op), C.byref(pc), _ptr(g), C.byref(prm),
                                      _ptr(s_out), C.byref(res))
        self._check(rc)
        return StpcgOutput(s_out, res.update_step_M_norm, int(res.num_iterations),
                           capi.EXIT_NAMES.get(res.exit_reason, str(res.exit_reason)),
                           res.r0_norm, res.final_rv, int(res.kernel_launches),
                           float(res.solve_kernel_ms))

    def hvp(self, H: "OperatorHandle", v: torch.Tensor, out=None):
        if out is None:
            out = torch.empty_like(v)
        self._check(self.lib.ob200_hvp(self.h, C.byref(H.op), _ptr(v), _ptr(out)))
        return out

    # -- level 1 -------------------------------------------------------------------
    def dot(self, a: torch.Tensor, b: torch.Tensor) -> float:
        r = C.c_double(0)
        self._check(self.lib.ob200_dot(self.h, a.numel(), _ptr(a), _ptr(b), C.byref(r)))
        return r.value

    def axpby(self, alpha, x, beta, y, out=None):
        if out is None:
            out = torch.empty_like(x)
        self._check(self.lib.ob200_axpby(self.h, x.numel(), float(alpha), _ptr(x), float(beta),
                                         _ptr(y), _ptr(out)))
        return out

    def stiefel_model(self, A_bf16, Y, want_grad=True):
        n, p = Y.shape
        S = np.zeros((p, p))
        f = C.c_double(0)
        bound = C.c_double(0)
        grad = torch.empty_like(Y) if want_grad else None
        self._check(self.lib.ob200_stiefel_model(self.h, n, p, _ptr(A_bf16), _ptr(Y), _ptr(S),
                                                 C.byref(f), _ptr(grad), C.byref(bound)))
        return S, f.value, grad, bound.value

    def stiefel_retract(self, Y, V, out=None):
        n, p = Y.shape
        if out is None:
            out = torch.empty_like(Y)
        self._check(self.lib.ob200_stiefel_retract(self.h, n, p, _ptr(Y), _ptr(V), _ptr(out)))
        return out


class OperatorHandle:
    def __init__(self, ctx, op, keep):
        self.ctx, self.op, self._keep = ctx, op, keep

    def step_bytes(self, precon=False):
        pc = capi.Precon()
        pc.kind = capi.PRECON_JACOBI if precon else capi.PRECON_NONE
        return int(self.ctx.lib.ob200_stpcg_step_bytes(C.byref(self.op), C.byref(pc)))

    def hvp_bytes(self):
        return int(self.ctx.lib.ob200_hvp_bytes(C.byref(self.op)))
