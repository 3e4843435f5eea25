"""ctypes front-end of oracle/croutines.c (TEST INFRASTRUCTURE, see oracle/__init__.py).

Function names, argument order and in-place semantics follow the reference's
``pyimcom.routine`` (routine.py:125-588) so parity tests read like the reference's own tests.
"""

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = C.CDLL(_build.build())
_dp = C.POINTER(C.c_double)
_NT = int(os.environ.get("ORACLE_THREADS", "1"))


def set_threads(n: int) -> None:
    """Number of OpenMP threads used by the interpolation / bisection / CG loops."""
    global _NT
    _NT = max(1, int(n))


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous, "oracle wants C-contiguous float64"
    return a.ctypes.data_as(_dp)


def iD5512C_getw(w, fh):
    """routine.py:29-122."""
    _lib.orc_iD5512C_getw(_d(w), C.c_double(fh))


def iD5512C(infunc, xpos, ypos, fhatout):
    """routine.py:125-181."""
    nlayer, ngy, ngx = infunc.shape
    _lib.orc_iD5512C(_d(infunc), C.c_long(nlayer), C.c_long(ngy), C.c_long(ngx), _d(xpos), _d(ypos),
                     C.c_long(xpos.size), _d(fhatout), C.c_int(_NT))


def iD5512C_sym(infunc, xpos, ypos, fhatout):
    """routine.py:184-253."""
    nlayer, ngy, ngx = infunc.shape
    _lib.orc_iD5512C_sym(_d(infunc), C.c_long(nlayer), C.c_long(ngy), C.c_long(ngx), _d(xpos),
                         _d(ypos), C.c_long(xpos.size), _d(fhatout), C.c_int(_NT))


def gridD5512C(infunc, xpos, ypos, fhatout):
    """routine.py:256-338."""
    ngy, ngx = infunc.shape
    npi, nxo = xpos.shape[:2]
    nyo = ypos.shape[1]
    _lib.orc_gridD5512C(_d(infunc), C.c_long(ngy), C.c_long(ngx), _d(xpos), _d(ypos), C.c_long(npi),
                        C.c_long(nxo), C.c_long(nyo), _d(fhatout), C.c_int(_NT))


def lakernel1(lam, Q, mPhalf, C_, targetleak, kCmin, kCmax, nbis, kappa, Sigma, UC, T, smax):
    """routine.py:341-430.  kappa/Sigma/UC may be float32 views (lakernel.py:216-218)."""
    m, n = mPhalf.shape
    k64, s64, u64 = (np.zeros(m) for _ in range(3))
    _lib.orc_lakernel1(_d(lam), _d(mPhalf), C.c_long(m), C.c_long(n), C.c_double(C_),
                       C.c_double(targetleak), C.c_double(kCmin), C.c_double(kCmax), C.c_long(nbis),
                       _d(k64), _d(s64), _d(u64), _d(T), C.c_double(smax), C.c_int(_NT))
    kappa[:] = k64
    Sigma[:] = s64
    UC[:] = u64


def lsolve_sps(N, A, x, b):
    """routine.py:433-484 (destroys A)."""
    _lib.orc_lsolve_sps(C.c_long(N), _d(A), _d(x), _d(np.ascontiguousarray(b)))


def build_reduced_T_wrap(Nflat, Dflat, Eflat, kappa, ucmin, smax, out_kappa, out_Sigma, out_UC, out_w,
                         out_iv=None, out_branch=None):
    """routine.py:487-588; optional int32 out_iv/out_branch record the discrete decisions."""
    nv = kappa.size
    m = out_kappa.size
    ip = C.POINTER(C.c_int32)
    _lib.orc_build_reduced_T(_d(Nflat), _d(Dflat), _d(Eflat), _d(np.ascontiguousarray(kappa, dtype=np.float64)),
                             C.c_long(nv), C.c_long(m), C.c_double(ucmin), C.c_double(smax),
                             _d(out_kappa), _d(out_Sigma), _d(out_UC), _d(out_w),
                             out_iv.ctypes.data_as(ip) if out_iv is not None else None,
                             out_branch.ctypes.data_as(ip) if out_branch is not None else None)


def iterative_wrapper(AA, mBhalf, relevant, rtol, maxiter, niter=None):
    """lakernel.py:548-590: per-pixel gathered CG; returns float32 Ti (m, n)."""
    m, n = mBhalf.shape
    Ti = np.zeros((m, n), dtype=np.float32)
    rel = np.ascontiguousarray(relevant, dtype=np.uint8)
    _lib.orc_iterative_wrapper(_d(AA), _d(mBhalf), rel.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_long(m),
                               C.c_long(n), C.c_double(rtol), C.c_long(maxiter),
                               Ti.ctypes.data_as(C.POINTER(C.c_float)),
                               niter.ctypes.data_as(C.POINTER(C.c_int32)) if niter is not None else None,
                               C.c_int(_NT))
    return Ti
