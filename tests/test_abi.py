"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/pyimcom_b200.h declares, the ctypes prototypes cover exactly those symbols with the right arity, and the
Python seams expose the reference's names.  No compute call is made (there is no GPU here)."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pyimcom_b200.h")


def declared():
    """{name: number of parameters} parsed from the public header."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long long|size_t|const char\*)\s+(b200_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


@pytest.fixture(scope="module")
def lib():
    from pyimcom_b200 import _lib

    return _lib


def test_library_exports_every_declared_symbol(lib):
    decl = declared()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(lib.lib, name), f"{name} declared in include/pyimcom_b200.h but not exported"
    assert lib.version() == 100
    assert lib.launch_count() == 0  # nothing launched on import


def test_ctypes_prototypes_match_header(lib):
    decl = declared()
    special = {"b200_last_error", "b200_version", "b200_launch_count", "b200_ozaki_gemm_work_bytes",
               "b200_chol_work_bytes", "b200_ozaki_slices", "b200_eigh_fallback_count"}  # (non-int return types: bound by hand in _lib.py)
    assert set(lib.PROTOTYPES) | special == set(decl)
    for name, argtypes in lib.PROTOTYPES.items():
        assert len(argtypes) == decl[name], f"{name}: ctypes arity {len(argtypes)} != header {decl[name]}"


def test_struct_layouts(lib):
    # sizes a C compiler gives the three structs that cross the boundary (LP64)
    assert ctypes.sizeof(lib.TableRef) == 24
    assert ctypes.sizeof(lib.SolveSys) == 72
    assert ctypes.sizeof(lib.FinalizeArgs) == 184
    assert ctypes.sizeof(lib.PairDesc) == 40
    assert ctypes.sizeof(lib.EighProblem) == 40
    assert ctypes.sizeof(lib.AsmDesc) == 1056  # 648 + 324 + 40 + 36 + 4, rounded up to the 8-byte alignment
    assert ctypes.sizeof(lib.PartCell) == 24 and lib.PART_MAXSLOT == 256


def test_seams_expose_reference_names():
    from pyimcom_b200 import pyimcom_croutines as pc

    for name in ("iD5512C", "iD5512C_sym", "gridD5512C", "iD5512C_getw", "lakernel1", "lsolve_sps",
                 "build_reduced_T_wrap"):  # routine.py:29-588
        assert callable(getattr(pc, name))
    from pyimcom_b200 import lakernel as lk

    for name in ("CholKernel", "EigenKernel", "IterKernel", "EmpirKernel"):  # coadd.py:839-844
        cls = getattr(lk, name)
        assert callable(cls) and hasattr(cls, "__call__")


def test_product_never_imports_oracle():
    """The product path must not route through the CPU oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "pyimcom_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "liboracle" not in txt


def test_kernel_raises_without_gpu():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cases
    from pyimcom_b200 import lakernel as lk

    outst = cases.la_outst([1e-2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lk.CholKernel(outst)()
    assert not hasattr(outst, "T") or not isinstance(getattr(outst, "T", None), np.ndarray)


@pytest.mark.parametrize("modname", ["pyimcom_croutines", "furry_parakeet.pyimcom_croutines"])
def test_reference_import_chain_binds_the_b200_callables(modname):
    """The function seam as the reference really resolves it: psfutil.py:37-49 and lakernel.py:41-47 try
    furry_parakeet.pyimcom_croutines, then a top-level pyimcom_croutines, then their own .routine (tests/pyimcom/
    test_missing.py:11-24 exercises the chain).  With pyimcom_b200.pyimcom_croutines registered under either name the
    reference's own modules, imported verbatim from /root/reference, must hold the b200 callables.  (Child process: the
    import chain runs once per interpreter.  Build container only: /root/reference does not exist on the GPU box.)"""
    import subprocess
    import sys

    from oracle import refhost

    if not refhost.available():
        pytest.skip("needs /root/reference")
    code = f"""
import sys, types
sys.path.insert(0, {ROOT!r})
import pyimcom_b200.pyimcom_croutines as B
name = {modname!r}
if "." in name:
    pkg = types.ModuleType(name.split(".")[0]); pkg.__path__ = []
    setattr(pkg, name.split(".")[1], B)
    sys.modules[name.split(".")[0]] = pkg
sys.modules[name] = B
from oracle import refhost
ref = refhost.load()
for fn in ("iD5512C", "iD5512C_sym", "gridD5512C"):
    assert getattr(ref.psfutil, fn) is getattr(B, fn), fn
for fn in ("lakernel1", "build_reduced_T_wrap"):
    assert getattr(ref.lakernel, fn) is getattr(B, fn), fn
# the process-global interpolator selector the reference calls through (psfutil.py:52-76) points at them as well
P = ref.psfutil.PSFInterpolator
assert P.gridC is B.gridD5512C and P.iC is B.iD5512C and P.iC_sym is B.iD5512C_sym
print("bound")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "bound" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


def test_workspace_sizes_of_the_sliced_path(lib):
    """Host-side contract of the INT8 path (no compute call): 8 digit planes per operand; the per-system scratch of
    b200_dev_chol_solve holds the digit planes of W and X (8 bytes per matrix element -- exactly the size of the float64
    matrices they describe) plus the row scales, and the GEMM scratch the planes of both operands."""
    assert lib.lib.b200_ozaki_slices() == 8
    npad, mpad = 6272, 1536
    wb = lib.lib.b200_chol_work_bytes(npad, mpad)
    planes = 8 * npad * (npad + mpad)
    scales = 8 * (2 * npad + (npad // 128) * mpad)
    assert planes + scales <= wb <= planes + scales + 4096
    assert lib.lib.b200_chol_work_bytes(128, 0) < lib.lib.b200_chol_work_bytes(256, 0)
    M, N, K = 256, 128, 320
    gb = lib.lib.b200_ozaki_gemm_work_bytes(M, N, K)
    assert 8 * (M + N) * K + 8 * (M + N) <= gb <= 8 * (M + N) * K + 8 * (M + N) + 8192
    assert ctypes.sizeof(lib.SolveSys) == 72 and lib.SolveSys.work.offset == 56
