"""ctypes binding of libpyimcom_b200.so (C ABI: include/pyimcom_b200.h).

The product has NO CPU fallback: importing this module raises if the CUDA library has not been built
(``python -m pyimcom_b200.build``), and every wrapper raises ``B200Error`` on a non-zero return code.
"""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(HERE, "lib", "libpyimcom_b200.so")  # B200_LIB: A/B builds


class B200Error(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m pyimcom_b200.build` (nvcc, sm_100a). "
        "pyimcom_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

NB = 128
MAXB = 16

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_long
f64 = C.c_double
szt = C.c_size_t


class TableRef(C.Structure):
    """b200_table_ref"""

    _fields_ = [("offset", C.c_longlong), ("flip", C.c_int), ("pad_", C.c_int), ("penalty_sub", C.c_double)]


class SolveSys(C.Structure):
    """b200_solve_sys"""

    _fields_ = [("W", vp), ("X", vp), ("Dinv", vp), ("info", vp), ("npad", i32), ("mpad", i32), ("ldw", i32),
                ("ldx", i32), ("mrows", i32), ("pad_", i32), ("work", vp), ("work_bytes", szt)]


class PairDesc(C.Structure):
    """b200_pair_desc"""

    _fields_ = [("offA", i32), ("nA", i32), ("offB", i32), ("nB", i32), ("out", C.c_longlong), ("ld", i32),
                ("lut", i32), ("same", i32), ("pad_", i32)]


PART_MAXSLOT = 256  # B200_PART_MAXSLOT


class PartCell(C.Structure):
    _fields_ = [("bottom", C.c_int), ("left", C.c_int), ("h", C.c_int), ("w", C.c_int), ("off", C.c_longlong)]


class AsmDesc(C.Structure):
    """b200_asm_desc"""

    _fields_ = [("blk", C.c_longlong * 81), ("ld", i32 * 81), ("seg_start", i32 * 10), ("inst_off", i32 * 9),
                ("pad_", i32)]


class EighProblem(C.Structure):
    """b200_eigh_problem"""

    _fields_ = [("A", vp), ("Vt", vp), ("lam", vp), ("lda", i32), ("ldv", i32), ("n", i32), ("pad_", i32)]


class FinalizeArgs(C.Structure):
    """b200_finalize_args"""

    _fields_ = [("Tpi", vp), ("strideT", szt), ("ldt", i32), ("w", vp), ("nv", i32), ("mB", vp), ("ldb", i32),
                ("m", i32), ("n", i32), ("n2f", i32), ("fade", i32), ("fade_w", vp), ("indata", vp), ("ldi", i32),
                ("n_inframe", i32), ("seg_end", vp), ("seg_img", vp), ("nseg", i32), ("n_img", i32), ("T32", vp),
                ("ldt32", i32), ("Ti64", vp), ("ldt64", i32), ("D", vp), ("N", vp), ("outimage", vp),
                ("Tsum_image", vp)]


# name -> argtypes, exactly the declarations of include/pyimcom_b200.h (checked by tests/test_abi.py)
PROTOTYPES = {
    "b200_release_scratch": [],
    "b200_profile": [i32],
    "b200_profile_read": [i32, C.POINTER(f64), C.POINTER(f64), C.POINTER(C.c_longlong)],
    "b200_iD5512C": [vp, i32, i32, i32, vp, vp, i64, vp],
    "b200_iD5512C_sym": [vp, i32, i32, i32, vp, vp, i64, vp],
    "b200_gridD5512C": [vp, i32, i32, vp, vp, i64, i32, i32, vp],
    "b200_iD5512C_getw": [vp, f64],
    "b200_lakernel1": [vp, vp, i64, i64, f64, f64, f64, f64, i32, vp, vp, vp, vp, f64],
    "b200_lsolve_sps": [i32, vp, vp, vp],
    "b200_build_reduced_T_wrap": [vp, vp, vp, vp, i32, i64, f64, f64, vp, vp, vp, vp, vp, vp],
    "b200_dev_iD5512C": [vp, i32, i32, i32, vp, vp, i64, vp, vp],
    "b200_dev_iD5512C_sym": [vp, i32, i32, i32, vp, vp, i64, vp, vp],
    "b200_dev_gridD5512C": [vp, i32, i32, vp, vp, i64, i32, i32, vp, vp],
    "b200_dev_layout_tables": [vp, i32, i32, i32, i32, i32, vp, vp],
    "b200_dev_cmul_conj": [vp, vp, vp, vp, C.c_longlong, C.c_longlong, C.c_longlong, f64, vp, vp, vp],
    "b200_dev_gather_stamp": [vp, i32, i32, vp, vp, vp, vp, i64, i32, vp, vp, vp, vp, i32, vp],
    "b200_dev_build_A": [vp, vp, vp, i32, i32, vp, vp, i32, i32, i32, f64, f64, f64, vp, i32, f64, i32, vp],
    "b200_dev_pair_blocks": [vp, vp, vp, vp, vp, i32, i32, vp, vp, i32, i32, f64, f64, f64, i32, vp, f64, vp],
    "b200_dev_assemble_A": [C.POINTER(AsmDesc), vp, i32, i32, vp, vp, i32, f64, vp],
    "b200_dev_build_B": [vp, vp, vp, i32, i32, vp, vp, i32, i32, f64, f64, i32, i32, f64, f64, vp, i32, szt, vp],
    "b200_dev_chol_solve": [C.POINTER(SolveSys), i32, i32, i32, vp],
    "b200_dev_pad_system": [vp, i32, i32, i32, vp, i32, C.POINTER(f64), i32, vp],
    "b200_dev_gemm_nt": [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp],
    "b200_dev_ozaki_gemm_nt": [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, szt, vp],
    "b200_dev_transpose": [vp, i32, vp, i32, i32, i32, vp],
    "b200_dev_eigh_batch": [C.POINTER(EighProblem), i32, i32, C.POINTER(i32), vp],
    "b200_dev_eigh": [vp, i32, i32, vp, i32, vp, i32, C.POINTER(i32), vp],
    "b200_dev_tridiag": [vp, i32, i32, vp, vp, vp, vp],
    "b200_dev_lakernel1": [vp, vp, i32, i32, i32, f64, f64, f64, f64, i32, vp, vp, vp, vp, i32, f64, vp],
    "b200_dev_eigen_single": [vp, vp, i32, i32, i32, f64, f64, vp, vp, vp, i32, vp],
    "b200_dev_lsolve_sps": [i32, vp, vp, vp, vp, vp],
    "b200_dev_build_reduced_T": [vp, vp, vp, vp, i32, i32, f64, f64, vp, vp, vp, vp, vp, vp, vp],
    "b200_dev_node_stats": [vp, i32, vp, i32, szt, i32, i32, i32, C.POINTER(f64), f64, vp, vp, vp, vp, vp, vp, vp],
    "b200_dev_rowdot": [vp, i32, vp, i32, i32, i32, vp, i32, vp],
    "b200_dev_single_kappa_maps": [vp, vp, vp, i32, f64, f64, vp, vp, vp, vp],
    "b200_dev_empir_T": [vp, vp, vp, vp, i32, i32, i32, i32, f64, vp, i32, vp],
    "b200_dev_scale": [vp, f64, i32, vp, vp],
    "b200_dev_iter_cg": [vp, i32, f64, vp, i32, i32, i32, vp, vp, vp, vp, f64, f64, i32, vp, i32, vp, vp, vp],
    "b200_dev_finalize": [C.POINTER(FinalizeArgs), vp],
    "b200_dev_stamp_maps": [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp],
    "b200_dev_accumulate": [vp, i32, i32, i32, vp, i32, i32, i32, vp],
    "b200_dev_accumulate_stamp": [vp, i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, i32,
                                  vp],
    "b200_dev_unfade_crop": [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp],
    "b200_dev_compress_map": [vp, C.c_long, i32, i32, vp, vp],
    "b200_dev_partition": [vp, i32, vp, vp, vp, i32, vp, i32, i32, f64, f64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                           vp, vp, vp],
    "b200_dev_extract_layers": [vp, i32, i32, vp, vp, vp, i32, i32, i32, vp, vp],
    "b200_dev_assemble_instamps": [vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, C.c_longlong, vp, vp, vp, vp, vp],
}

lib.b200_last_error.restype = C.c_char_p
lib.b200_last_error.argtypes = []
lib.b200_version.restype = C.c_int
lib.b200_version.argtypes = []
lib.b200_ozaki_gemm_work_bytes.restype = C.c_size_t
lib.b200_ozaki_gemm_work_bytes.argtypes = [i32, i32, i32]
lib.b200_ozaki_slices.restype = C.c_int
lib.b200_ozaki_slices.argtypes = []
lib.b200_chol_work_bytes.restype = C.c_size_t
lib.b200_chol_work_bytes.argtypes = [i32, i32]
lib.b200_launch_count.restype = C.c_longlong
lib.b200_launch_count.argtypes = []
lib.b200_eigh_fallback_count.restype = C.c_longlong
lib.b200_eigh_fallback_count.argtypes = []


def _wrap(name, argtypes):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = C.c_int

    def call(*args):
        rc = fn(*args)
        if rc != 0:
            raise B200Error(f"{name} -> {rc}: {lib.b200_last_error().decode(errors='replace')}")

    call.__name__ = name
    return call


for _n, _a in PROTOTYPES.items():
    globals()[_n[len("b200_"):]] = _wrap(_n, _a)
profile_read_raw = globals()["profile_read"]


PROF_KINDS = ("chol_super_update", "potrf_diag", "chol_panel", "chol_inner_update", "back_super_update", "back_diag",
              "back_inner_update", "build_A", "build_B", "finalize", "gemm_nt", "iter_cg", "lakernel1", "eigh",
              "assemble_A", "oz_slice", "oz_gemm")


def profile_read():
    """{kind: (total ms, total algorithmic work, launches)} for every kind with at least one recorded launch."""
    out = {}
    for k, name in enumerate(PROF_KINDS):
        ms, work, cnt = f64(0), f64(0), C.c_longlong(0)
        globals()["profile_read_raw"](k, C.byref(ms), C.byref(work), C.byref(cnt))
        if cnt.value:
            out[name] = (ms.value, work.value, cnt.value)
    return out


def launch_count() -> int:
    return int(lib.b200_launch_count())


def eigh_fallback_count() -> int:
    """How many eigenproblems the tridiagonalisation-based solver handed to the Jacobi solver so far."""
    return int(lib.b200_eigh_fallback_count())


def version() -> int:
    return int(lib.b200_version())
