"""ctypes binding of liboptimization_b200.so (include/optimization_b200.h).

This is the Python face of the drop-in boundary: every call goes through the C
ABI exactly as the reference-side C++ host templates do.  There is NO fallback:
if the shared library (built in-tree by `make -C optimization_b200/csrc` /
`__graft_entry__.build()`) is missing, importing this module raises, and if no
GPU is usable `Context()` raises.  torch is used only to own device memory and
the CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OB200_LIB: developer override (profiling builds of the same library, tools/ only)
LIB_PATH = os.environ.get("OB200_LIB") or os.path.join(_HERE, "liboptimization_b200.so")

OK, INVALID_ARGUMENT, CUDA_ERROR, UNSUPPORTED, NUMERIC_RANGE, ABORTED = range(6)
EXIT_RESIDUAL, EXIT_MAX_ITERATIONS, EXIT_KERNEL, EXIT_BOUNDARY = range(4)
EXIT_NAMES = {0: "residual", 1: "max_iterations", 2: "kernel", 3: "boundary"}
OP_DIAG, OP_STIEFEL_BLOCKDIAG, OP_SPHERE_LOWRANK, OP_BLOCK_CSR3, OP_STENCIL7, OP_HOST_CALLBACK = 1, 2, 3, 4, 5, 6
PRECON_NONE, PRECON_JACOBI, PRECON_HOST_CALLBACK, PRECON_STIEFEL_PROJECTED_JACOBI = 0, 1, 2, 3

# every symbol include/optimization_b200.h declares (checked by the CPU test-suite)
EXPORTS = [
    "ob200_create", "ob200_destroy", "ob200_last_error", "ob200_version", "ob200_sm_count",
    "ob200_synchronize", "ob200_kernel_launches", "ob200_stpcg", "ob200_stpcg_host", "ob200_hvp",
    "ob200_dot", "ob200_dots", "ob200_axpby", "ob200_hadamard", "ob200_stiefel_model",
    "ob200_stiefel_retract", "ob200_stiefel_project", "ob200_sphere_model", "ob200_sphere_retract", "ob200_lobpcg", "ob200_block_apply", "ob200_malloc",
    "ob200_free", "ob200_memcpy_h2d", "ob200_memcpy_d2h", "ob200_malloc_host", "ob200_free_host",
    "ob200_comm_export", "ob200_comm_connect", "ob200_comm_rank", "ob200_comm_world",
    "ob200_stpcg_step_bytes", "ob200_hvp_bytes", "ob200_debug_phase_times",
    "ob200_debug_block_apply", "ob200_debug_sym_eig32", "ob200_set_option", "ob200_last_path", "ob200_div", "ob200_csr3_model", "ob200_csr3_retract", "ob200_halo_create", "ob200_halo_connect",
]


class Operator(C.Structure):
    _fields_ = [("kind", C.c_int), ("n", C.c_uint64), ("p", C.c_uint64),
                ("diag_dev", C.c_void_p), ("A_bf16_dev", C.c_void_p), ("Y_dev", C.c_void_p),
                ("S_host", C.c_void_p), ("op_norm_bound", C.c_double), ("x_dev", C.c_void_p),
                ("U_dev", C.c_void_p), ("sigma_host", C.c_void_p), ("k", C.c_uint64),
                ("xAx", C.c_double), ("Ax_dev", C.c_void_p), ("ldu", C.c_uint64),
                ("csr_rowptr_dev", C.c_void_p), ("csr_colidx_dev", C.c_void_p), ("csr_blocks_dev", C.c_void_p),
                ("csr_lambda_dev", C.c_void_p), ("csr_nnz", C.c_uint64), ("gx", C.c_uint32), ("gy", C.c_uint32),
                ("gz", C.c_uint32), ("csr_n_halo", C.c_uint64), ("halo_send_idx_dev", C.c_void_p),
                ("halo_send_ptr", C.c_uint64 * 9), ("halo_dst_off", C.c_uint64 * 8),
                ("apply", C.c_void_p), ("apply_user", C.c_void_p)]


class BlockOperator(C.Structure):
    _fields_ = [("kind", C.c_int), ("diag_dev", C.c_void_p), ("alpha", C.c_double), ("gx", C.c_uint32),
                ("gy", C.c_uint32), ("gz", C.c_uint32)]


BLK_DIAG, BLK_STENCIL7, BLK_SCALAR = 1, 2, 3


class Precon(C.Structure):
    _fields_ = [("kind", C.c_int), ("minv_dev", C.c_void_p), ("apply", C.c_void_p), ("apply_user", C.c_void_p)]


class StpcgParams(C.Structure):
    _fields_ = [("Delta", C.c_double), ("max_iterations", C.c_uint64), ("kappa_fgr", C.c_double),
                ("theta", C.c_double), ("epsilon", C.c_double)]


class StpcgResult(C.Structure):
    _fields_ = [("update_step_M_norm", C.c_double), ("num_iterations", C.c_uint64),
                ("exit_reason", C.c_int), ("r0_norm", C.c_double), ("final_rv", C.c_double),
                ("kernel_launches", C.c_uint64), ("solve_kernel_ms", C.c_float)]


class TntParams(C.Structure):
    """Field for field TNTParams (reference TNT.h:76-130 + Concepts.h:42-60,116-131)."""
    _fields_ = [("max_iterations", C.c_uint64), ("max_computation_time", C.c_double),
                ("gradient_tolerance", C.c_double), ("relative_decrease_tolerance", C.c_double),
                ("stepsize_tolerance", C.c_double),
                ("preconditioned_gradient_tolerance", C.c_double),
                ("Delta_tolerance", C.c_double), ("Delta0", C.c_double), ("eta1", C.c_double),
                ("eta2", C.c_double), ("alpha1", C.c_double), ("alpha2", C.c_double),
                ("max_TPCG_iterations", C.c_uint64), ("kappa_fgr", C.c_double),
                ("theta", C.c_double)]


class TntResult(C.Structure):
    _fields_ = [("status", C.c_int), ("f", C.c_double), ("gradfx_norm", C.c_double),
                ("preconditioned_grad_f_x_norm", C.c_double), ("elapsed_time", C.c_double),
                ("n_outer", C.c_uint64), ("n_trace", C.c_uint64), ("cap", C.c_uint64),
                ("inner_iterations", C.POINTER(C.c_uint64)), ("gain_ratios", C.POINTER(C.c_double)),
                ("update_step_norms", C.POINTER(C.c_double)),
                ("update_step_M_norms", C.POINTER(C.c_double)),
                ("trust_region_radius", C.POINTER(C.c_double)),
                ("objective_values", C.POINTER(C.c_double)),
                ("gradient_norms", C.POINTER(C.c_double)), ("kernel_launches", C.c_uint64)]


COMM_HANDLE_BYTES = 64


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `make -C optimization_b200/csrc` "
            "(or __graft_entry__.build()).  optimization_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp, u64, dbl, i = C.c_void_p, C.c_uint64, C.c_double, C.c_int
    lib.ob200_create.argtypes = [i, vp, C.POINTER(vp)]
    lib.ob200_destroy.argtypes = [vp]
    lib.ob200_last_error.argtypes = [vp]
    lib.ob200_last_error.restype = C.c_char_p
    lib.ob200_sm_count.argtypes = [vp]
    lib.ob200_synchronize.argtypes = [vp]
    lib.ob200_kernel_launches.argtypes = [vp]
    lib.ob200_kernel_launches.restype = u64
    lib.ob200_stpcg.argtypes = [vp, C.POINTER(Operator), C.POINTER(Precon), vp,
                                C.POINTER(StpcgParams), vp, C.POINTER(StpcgResult)]
    lib.ob200_stpcg_host.argtypes = lib.ob200_stpcg.argtypes
    lib.ob200_hvp.argtypes = [vp, C.POINTER(Operator), vp, vp]
    lib.ob200_dot.argtypes = [vp, u64, vp, vp, C.POINTER(dbl)]
    lib.ob200_dots.argtypes = [vp, u64, i, C.POINTER(vp), C.POINTER(vp), C.POINTER(dbl)]
    lib.ob200_axpby.argtypes = [vp, u64, dbl, vp, dbl, vp, vp]
    lib.ob200_hadamard.argtypes = [vp, u64, vp, vp, vp]
    lib.ob200_div.argtypes = [vp, u64, vp, dbl, vp]
    lib.ob200_stiefel_model.argtypes = [vp, u64, u64, vp, vp, vp, C.POINTER(dbl), vp,
                                        C.POINTER(dbl)]
    lib.ob200_stiefel_retract.argtypes = [vp, u64, u64, vp, vp, vp]
    lib.ob200_stiefel_project.argtypes = [vp, u64, u64, vp, vp, vp]
    lib.ob200_sphere_model.argtypes = [vp, u64, u64, vp, vp, u64, vp, vp, vp, C.POINTER(dbl), vp]
    lib.ob200_sphere_retract.argtypes = [vp, u64, vp, vp, vp]
    lib.ob200_csr3_model.argtypes = [vp, u64, u64, vp, vp, vp, vp, vp, C.POINTER(dbl), vp]
    lib.ob200_csr3_retract.argtypes = [vp, u64, u64, vp, vp, vp]
    bo = C.POINTER(BlockOperator)
    lib.ob200_lobpcg.argtypes = [vp, bo, bo, bo, u64, u64, vp, u64, u64, dbl, vp, C.POINTER(dbl), C.POINTER(u64),
                                 C.POINTER(u64)]
    lib.ob200_block_apply.argtypes = [vp, bo, u64, u64, vp, u64, vp, u64]
    lib.ob200_malloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.ob200_free.argtypes = [vp, vp]
    lib.ob200_memcpy_h2d.argtypes = [vp, vp, vp, C.c_size_t]
    lib.ob200_memcpy_d2h.argtypes = [vp, vp, vp, C.c_size_t]
    lib.ob200_malloc_host.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    lib.ob200_free_host.argtypes = [vp, vp]
    lib.ob200_halo_create.argtypes = [vp, u64, vp]
    lib.ob200_halo_connect.argtypes = [vp, vp]
    lib.ob200_comm_export.argtypes = [vp, vp]
    lib.ob200_comm_connect.argtypes = [vp, i, i, vp]
    lib.ob200_comm_rank.argtypes = [vp]
    lib.ob200_comm_world.argtypes = [vp]
    lib.ob200_stpcg_step_bytes.argtypes = [C.POINTER(Operator), C.POINTER(Precon)]
    lib.ob200_stpcg_step_bytes.restype = u64
    lib.ob200_hvp_bytes.argtypes = [C.POINTER(Operator)]
    lib.ob200_hvp_bytes.restype = u64
    lib.ob200_set_option.argtypes = [vp, C.c_char_p, i]
    lib.ob200_last_path.argtypes = [vp]
    lib.ob200_debug_block_apply.argtypes = [vp, u64, vp, vp, vp, i]
    lib.ob200_debug_sym_eig32.argtypes = [vp, vp, vp]
    lib.ob200_debug_phase_times.argtypes = [vp, i, C.POINTER(u64), C.POINTER(u64)]
    return lib


_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = load_library()
    return _LIB


class Ob200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ob200 status {code}: {msg}")
        self.code = code
