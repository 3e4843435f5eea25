"""Drop-in for ``furry_parakeet.pyimcom_croutines`` / ``pyimcom.routine`` on a B200.

Same function names, argument order and in-place NumPy semantics as the reference
(routine.py:125-588; import chains psfutil.py:37-49, lakernel.py:41-47): the caller allocates every
array, the callee writes in place and returns ``None``.  Each call copies its operands to the GPU, runs
the hand-written sm_100a kernel through the C ABI (include/pyimcom_b200.h, section 1) and copies the
result back.  There is no CPU fallback.

To make the unmodified reference pick this module up, expose it as a top-level ``pyimcom_croutines``
(see INTEGRATION.md): ``sys.modules["pyimcom_croutines"] = pyimcom_b200.pyimcom_croutines``.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _in(a):
    """C-contiguous float64 view/copy of an input array."""
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _Out:
    """In-place output: a C-contiguous float64 staging array written back to ``dst`` on exit.

    The reference passes float32 views for kappa/Sigma/UC (lakernel.py:216-218) and float64 elsewhere."""

    def __init__(self, dst, preload=False):
        self.dst = dst
        if dst.dtype == np.float64 and dst.flags.c_contiguous:
            self.buf = dst
        else:
            self.buf = np.ascontiguousarray(dst, dtype=np.float64) if preload else np.zeros(dst.shape, dtype=np.float64)

    def done(self):
        if self.buf is not self.dst:
            self.dst[...] = self.buf


def iD5512C_getw(w, fh):
    """routine.py:29-122."""
    o = _Out(w)
    _lib.iD5512C_getw(_p(o.buf), float(fh))
    o.done()


def iD5512C(infunc, xpos, ypos, fhatout):
    """routine.py:125-181: infunc (nlayer, ngy, ngx), xpos/ypos (nout,), fhatout (nlayer, nout) in place."""
    infunc, xpos, ypos = _in(infunc), _in(xpos), _in(ypos)
    nlayer, ngy, ngx = infunc.shape
    o = _Out(fhatout, preload=True)
    _lib.iD5512C(_p(infunc), nlayer, ngy, ngx, _p(xpos), _p(ypos), xpos.size, _p(o.buf))
    o.done()


def iD5512C_sym(infunc, xpos, ypos, fhatout):
    """routine.py:184-253."""
    infunc, xpos, ypos = _in(infunc), _in(xpos), _in(ypos)
    nlayer, ngy, ngx = infunc.shape
    o = _Out(fhatout, preload=True)
    _lib.iD5512C_sym(_p(infunc), nlayer, ngy, ngx, _p(xpos), _p(ypos), xpos.size, _p(o.buf))
    o.done()


def gridD5512C(infunc, xpos, ypos, fhatout):
    """routine.py:256-338: infunc (ngy, ngx), xpos (npi, nxo), ypos (npi, nyo), fhatout (npi, nyo*nxo)."""
    infunc, xpos, ypos = _in(infunc), _in(xpos), _in(ypos)
    ngy, ngx = infunc.shape
    npi, nxo = xpos.shape
    nyo = ypos.shape[1]
    o = _Out(fhatout)
    _lib.gridD5512C(_p(infunc), ngy, ngx, _p(xpos), _p(ypos), npi, nxo, nyo, _p(o.buf))
    o.done()


def lakernel1(lam, Q, mPhalf, C_, targetleak, kCmin, kCmax, nbis, kappa, Sigma, UC, T, smax):
    """routine.py:341-430 (``Q`` is unused, as in the reference)."""
    lam, mPhalf = _in(lam), _in(mPhalf)
    m, n = mPhalf.shape
    ok, oS, oU, oT = _Out(kappa), _Out(Sigma), _Out(UC), _Out(T)
    _lib.lakernel1(_p(lam), _p(mPhalf), m, n, float(C_), float(targetleak), float(kCmin), float(kCmax), int(nbis),
                   _p(ok.buf), _p(oS.buf), _p(oU.buf), _p(oT.buf), float(smax))
    for o in (ok, oS, oU, oT):
        o.done()


def lsolve_sps(N, A, x, b):
    """routine.py:433-484 (destroys A: it holds the Cholesky factor afterwards)."""
    oA = _Out(A, preload=True)
    ox = _Out(x)
    b = _in(b)
    _lib.lsolve_sps(int(N), _p(oA.buf), _p(ox.buf), _p(b))
    oA.done()
    ox.done()


def build_reduced_T_wrap(Nflat, Dflat, Eflat, kappa, ucmin, smax, out_kappa, out_Sigma, out_UC, out_w, out_iv=None,
                         out_branch=None):
    """routine.py:487-588; ``out_iv``/``out_branch`` (int32, optional) expose the discrete decisions."""
    Nflat, Dflat, Eflat, kappa = _in(Nflat), _in(Dflat), _in(Eflat), _in(kappa)
    nv = kappa.size
    m = out_kappa.size
    ok, oS, oU, ow = _Out(out_kappa), _Out(out_Sigma), _Out(out_UC), _Out(out_w)
    for a in (out_iv, out_branch):
        assert a is None or (a.dtype == np.int32 and a.flags.c_contiguous)
    _lib.build_reduced_T_wrap(_p(Nflat), _p(Dflat), _p(Eflat), _p(kappa), nv, m, float(ucmin), float(smax), _p(ok.buf),
                              _p(oS.buf), _p(oU.buf), _p(ow.buf), _p(out_iv) if out_iv is not None else None,
                              _p(out_branch) if out_branch is not None else None)
    for o in (ok, oS, oU, ow):
        o.done()
