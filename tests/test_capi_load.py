"""CPU: the C-ABI library loads and exports every symbol include/optimization_b200.h
declares; header and ctypes mirror agree; no compute calls (no GPU here)."""
import os
import re

from optimization_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "optimization_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ob200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(capi.EXPORTS) == syms
    assert lib.ob200_version() >= 100


def test_create_fails_loudly_without_gpu():
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        return
    lib = capi.load_library()
    h = C.c_void_p()
    assert lib.ob200_create(0, None, C.byref(h)) == capi.CUDA_ERROR
    assert not h.value


def test_byte_model():
    import ctypes as C
    lib = capi.load_library()
    op = capi.Operator()
    op.kind, op.n, op.p = capi.OP_STIEFEL_BLOCKDIAG, 100000, 32
    pc = capi.Precon()
    N = 100000 * 32
    A = 782 * 128 * 128 * 2
    assert lib.ob200_stpcg_step_bytes(C.byref(op), C.byref(pc)) == 12 * 8 * N + A
    assert lib.ob200_hvp_bytes(C.byref(op)) == 4 * 8 * N + A
    op.kind, op.n, op.p = capi.OP_DIAG, 1000, 1
    assert lib.ob200_stpcg_step_bytes(C.byref(op), C.byref(pc)) == 11 * 8 * 1000
    pc.kind = capi.PRECON_JACOBI
    assert lib.ob200_stpcg_step_bytes(C.byref(op), C.byref(pc)) == 13 * 8 * 1000
    # sphere, k = 16: (14 + 2k) n e per step, (6 + 2k) n e per stand-alone HVP
    pc.kind = capi.PRECON_NONE
    op.kind, op.n, op.p, op.k = capi.OP_SPHERE_LOWRANK, 1 << 24, 1, 16
    assert lib.ob200_stpcg_step_bytes(C.byref(op), C.byref(pc)) == 46 * 8 * (1 << 24)
    assert lib.ob200_hvp_bytes(C.byref(op)) == 38 * 8 * (1 << 24)


def test_sym_eig32_host_decomposition():
    """The eigen-decomposition S = Q diag(lambda) Q^T behind the rotated Stiefel solve (host code of ob200_stpcg):
    reconstruction, orthogonality and eigenvalues against numpy, on well- and ill-separated spectra."""
    import numpy as np
    lib = capi.load_library()
    rng = np.random.default_rng(7)
    for trial in range(6):
        B = rng.standard_normal((32, 32))
        S = 0.5 * (B + B.T)
        if trial == 1:
            S += np.diag(10.0 + np.arange(32))
        elif trial == 2:
            S = 1e-3 * S + 5.0 * np.eye(32)                      # clustered spectrum
        elif trial == 3:
            S = np.diag(np.linspace(-3.0, 4.0, 32))              # already diagonal
        elif trial == 4:
            S = np.zeros((32, 32))                               # zero matrix
        elif trial == 5:
            S *= 1e150                                           # large scale
        S = np.ascontiguousarray(S)
        Q = np.zeros((32, 32))
        lam = np.zeros(32)
        assert lib.ob200_debug_sym_eig32(S.ctypes.data, Q.ctypes.data, lam.ctypes.data) == 0
        scale = max(np.abs(S).max(), 1e-300)
        assert np.abs(Q @ np.diag(lam) @ Q.T - S).max() <= 1e-13 * scale * 32
        assert np.abs(Q.T @ Q - np.eye(32)).max() <= 1e-13
        assert np.allclose(np.sort(lam), np.linalg.eigvalsh(S), rtol=0, atol=1e-13 * scale * 32)


def test_product_never_imports_oracle():
    pk = os.path.join(ROOT, "optimization_b200")
    for dp, _, fs in os.walk(pk):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle[./](refapi|_ref|stpcg_port|liboracle)",
                                     txt, flags=re.M), f


def test_ctypes_mirror_matches_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct of include/optimization_b200.h as the C compiler lays it out == the ctypes
    mirror in optimization_b200/capi.py (an ABI drift would corrupt descriptors silently)."""
    import ctypes as C
    import subprocess
    structs = {"ob200_operator": capi.Operator, "ob200_precon": capi.Precon, "ob200_stpcg_params": capi.StpcgParams,
               "ob200_stpcg_result": capi.StpcgResult, "ob200_block_operator": capi.BlockOperator}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "optimization_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.splitlines():
        parts = line.split()
        cls = structs[parts[0]]
        assert int(parts[1]) == C.sizeof(cls), parts[0]
        offs = [getattr(cls, f).offset for f, _ in cls._fields_]
        assert [int(x) for x in parts[2:]] == offs, parts[0]


def test_python_constants_equal_the_header_defines():
    """Every OB200_OP_* / OB200_PRECON_* / OB200_EXIT_* / status value of include/optimization_b200.h has the same value
    in the ctypes mirror (a drift would select the wrong kernel or mis-report an exit silently)."""
    import re
    text = open(os.path.join(ROOT, "include", "optimization_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"^#define\s+OB200_([A-Z0-9_]+)\s+(-?\d+)\b", text, re.M)}
    assert defs["PRECON_STIEFEL_PROJECTED_JACOBI"] == 3 and defs["OP_HOST_CALLBACK"] == 6
    checked = 0
    for name, value in defs.items():
        if hasattr(capi, name):
            assert getattr(capi, name) == value, name
            checked += 1
    assert checked >= 12, checked
    for name in ("PRECON_NONE", "PRECON_JACOBI", "PRECON_HOST_CALLBACK", "PRECON_STIEFEL_PROJECTED_JACOBI", "OP_HOST_CALLBACK"):
        assert getattr(capi, name) == defs[name]
