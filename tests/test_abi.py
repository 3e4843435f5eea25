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
               "b200_chol_work_bytes", "b200_ozaki_slices"}  # (non-int return types: bound by hand in _lib.py)
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
