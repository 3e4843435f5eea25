"""CPU oracle of the linear-algebra kernels (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates ``pyimcom.lakernel`` (lakernel.py:50-744) on NumPy/SciPy LAPACK with the reference's
``Kernel(outst)(); outst.T/UC/Sigma/kappa`` contract.  Besides the float32 products the reference
emits, each kernel keeps the float64 intermediates (``self.f64``) that the P-f64 parity tests
compare: pre-cast ``Ti`` / ``Tpi``, ``D, N, E``, ``out_w``, bracket/branch words, CG
iteration counts.
"""

import warnings

import numpy as np
from scipy.linalg import LinAlgError, cho_solve, cholesky

from . import routines as R

ARCSEC = np.pi / 648000.0  # config.py:87


def rho_acc(cfg):
    """Acceptance radius in output pixels, lakernel.py:618 / coadd.py:923."""
    return (cfg.instamp_pad / ARCSEC) / (cfg.dtheta * 3600.0)


class _Kernel:
    """lakernel.py:50-138."""

    def __init__(self, outst):
        self.outst = outst
        cfg = outst.blk.cfg
        self.n_out = cfg.n_out
        self.n2f = cfg.n2f
        self.m = cfg.n2f**2
        self.n = int(outst.inpix_cumsum[-1])
        self.kappaC_arr = np.asarray(cfg.kappaC_arr, dtype=np.float64)
        self.nv = self.kappaC_arr.size
        self.ucmin = cfg.uctarget
        self.smax = cfg.sigmamax
        self.f64 = {}

    def __call__(self):
        o = self.outst
        shape = (self.n_out, self.n2f, self.n2f)
        if self.n == 0:  # lakernel.py:110-119
            o.T = np.zeros((self.n_out, self.m, 0), dtype=np.float32)
            o.UC = np.ones(shape, dtype=np.float32)
            o.Sigma = np.zeros(shape, dtype=np.float32)
            o.kappa = np.ones(shape, dtype=np.float32)
            return
        o.T = np.zeros((self.n_out, self.m, self.n), dtype=np.float32)
        self.UC_ = np.zeros((self.n_out, self.m), dtype=np.float32)
        self.Sigma_ = np.zeros((self.n_out, self.m), dtype=np.float32)
        self.kappa_ = np.zeros((self.n_out, self.m), dtype=np.float32)
        if self.nv == 1:
            self._single()
        else:
            self._multi()
        o.UC = self.UC_.reshape(shape)
        o.Sigma = self.Sigma_.reshape(shape)
        o.kappa = self.kappa_.reshape(shape)

    # shared tail of the multi-kappa Cholesky / Iterative kernels (lakernel.py:361-393, 703-741)
    def _reduce_nodes(self, j_out, Tpi, kappa_arr, Cj, Epq=None):
        o = self.outst
        mB = o.mhalfb[j_out]
        nv, m = self.nv, self.m
        Dp = np.einsum("ai,pai->ap", mB, Tpi)
        Npq = np.einsum("pai,qai->apq", Tpi, Tpi)
        if Epq is None:
            Epq = np.zeros((m, nv, nv))
            for p in range(nv):
                for q in range(p + 1):
                    Epq[:, q, p] = Epq[:, p, q] = Dp[:, q] - kappa_arr[p] * Npq[:, p, q]
        ok, oS, oU = np.zeros(m), np.zeros(m), np.zeros(m)
        ow = np.zeros(m * nv)
        iv = np.zeros(m, dtype=np.int32)
        br = np.zeros(m, dtype=np.int32)
        R.build_reduced_T_wrap(Npq.ravel().copy(), (Dp.ravel() / Cj).copy(), (Epq.ravel() / Cj).copy(),
                               self.kappaC_arr, self.ucmin, self.smax, ok, oS, oU, ow, iv, br)
        self.kappa_[j_out] = ok * Cj
        self.Sigma_[j_out] = oS
        self.UC_[j_out] = oU
        Ti = np.einsum("pai,ap->ai", Tpi, ow.reshape(m, nv))
        o.T[j_out] = Ti
        self.f64[j_out] = dict(Tpi=Tpi.copy(), Dp=Dp, Npq=Npq, Epq=Epq, out_w=ow.reshape(m, nv), iv=iv,
                               branch=br, Ti=Ti, kappa=ok * Cj, Sigma=oS, UC=oU)


class CholKernel(_Kernel):
    """lakernel.py:226-394."""

    repaired = 0

    @staticmethod
    def _cholesky_wrapper(AA, di, A):
        """lakernel.py:241-279: LAPACK potrf; on failure shift by |w0|+1e-16 and retry."""
        try:
            L = cholesky(AA, lower=True, check_finite=False)
        except LinAlgError:
            w = np.linalg.eigvalsh(A)
            shift = np.abs(w[0]) + 1e-16
            AA[di] += shift
            warnings.warn(f"oracle CholKernel: repaired negative eigenvalue {w[0]:19.12e}", stacklevel=2)
            L = cholesky(AA, lower=True, check_finite=False)
            AA[di] -= shift
            CholKernel.repaired += 1
        return L

    def _single(self):
        o = self.outst
        A, mB, Cv = o.sysmata, o.mhalfb, o.outovlc
        n = self.n
        di = np.diag_indices(n)
        for j in range(self.n_out):
            kap = self.kappaC_arr[0] * Cv[j]
            AA = A.copy()
            if kap:
                AA[di] += kap
            L = self._cholesky_wrapper(AA, di, A)
            Ti = cho_solve((L, True), mB[j].T, check_finite=False).T
            D = np.einsum("ai,ai->a", mB[j], Ti)
            N = np.einsum("ai,ai->a", Ti, Ti)
            self.kappa_[j] = kap
            self.Sigma_[j] = N
            self.UC_[j] = 1.0 - (kap * N + D) / Cv[j]
            o.T[j] = Ti
            self.f64[j] = dict(Ti=Ti, D=D, N=N, kappa=np.full(self.m, kap), Sigma=N,
                               UC=1.0 - (kap * N + D) / Cv[j])

    def _multi(self):
        o = self.outst
        A, mB, Cv = o.sysmata, o.mhalfb, o.outovlc
        nv, m, n = self.nv, self.m, self.n
        di = np.diag_indices(n)
        for j in range(self.n_out):
            AA = A.copy()
            kappa_arr = self.kappaC_arr * Cv[j]
            Tpi = np.zeros((nv, m, n))
            for p in range(nv):  # cumulative diagonal increments, lakernel.py:356
                AA[di] += kappa_arr[p] - (kappa_arr[p - 1] if p > 0 else 0)
                L = self._cholesky_wrapper(AA, di, A)
                Tpi[p] = cho_solve((L, True), mB[j].T, check_finite=False).T
            self._reduce_nodes(j, Tpi, kappa_arr, Cv[j])


class EigenKernel(_Kernel):
    """lakernel.py:141-223."""

    def _single(self):
        o = self.outst
        A, mB, Cv = o.sysmata, o.mhalfb, o.outovlc
        lam, Q = np.linalg.eigh(A)
        for k in range(self.n_out):
            P = mB[k] @ Q
            kap = self.kappaC_arr[0] * Cv[k]
            tt = P / (lam + kap)
            Sig = np.sum(tt**2, axis=1)
            UC = 1 - (lam + 2 * kap) / (lam + kap) ** 2 @ P.T**2 / Cv[k]
            Ti = tt @ Q.T
            self.kappa_[k] = kap
            self.Sigma_[k] = Sig
            self.UC_[k] = UC
            o.T[k] = Ti
            self.f64[k] = dict(Ti=Ti, kappa=np.full(self.m, kap), Sigma=Sig, UC=UC, lam=lam)

    def _multi(self, nbis=13):
        o = self.outst
        A, mB, Cv = o.sysmata, o.mhalfb, o.outovlc
        kCmin, kCmax = self.kappaC_arr[0], self.kappaC_arr[-1]
        lam, Q = np.linalg.eigh(A)
        m, n = self.m, self.n
        for k in range(self.n_out):
            tt = np.zeros((m, n))
            k64, s64, u64 = np.zeros(m), np.zeros(m), np.zeros(m)
            P = np.ascontiguousarray(mB[k] @ Q)
            R.lakernel1(lam, Q, P, Cv[k], self.ucmin, kCmin * Cv[k], kCmax * Cv[k], nbis, k64, s64, u64, tt,
                        self.smax)
            self.kappa_[k] = k64
            self.Sigma_[k] = s64
            self.UC_[k] = u64
            self.kappa_[k] *= Cv[k]  # second multiplication: reference quirk, lakernel.py:222
            Ti = tt @ Q.T
            o.T[k] = Ti
            self.f64[k] = dict(Ti=Ti, kappa=k64, Sigma=s64, UC=u64, lam=lam)


class IterKernel(_Kernel):
    """lakernel.py:533-744."""

    def _relevant(self):
        o = self.outst
        dy = o.yx_val[0].ravel()[:, None] - o.iny_val[None, :]
        dx = o.yx_val[1].ravel()[:, None] - o.inx_val[None, :]
        return np.hypot(dy, dx) < rho_acc(o.blk.cfg)

    def _single(self, exact_UC=False):
        o = self.outst
        cfg = o.blk.cfg
        A, mB, Cv = o.sysmata, o.mhalfb, o.outovlc
        rel = self._relevant()
        di = np.diag_indices(self.n)
        for j in range(self.n_out):
            AA = A.copy()
            kap = self.kappaC_arr[0] * Cv[j]
            if kap:
                AA[di] += kap
            nit = np.zeros(self.m, dtype=np.int32)
            Ti = R.iterative_wrapper(AA, np.ascontiguousarray(mB[j]), rel, cfg.iter_rtol, cfg.iter_max, nit)
            D = np.einsum("ai,ai->a", mB[j], Ti)
            N = np.einsum("ai,ai->a", Ti, Ti)
            self.kappa_[j] = kap
            self.Sigma_[j] = N
            if exact_UC:
                E = np.einsum("ij,ai,aj->a", A, Ti, Ti)
                UC = 1.0 + (E - 2 * D) / Cv[j]
            else:
                UC = 1.0 - (kap * N + D) / Cv[j]
            self.UC_[j] = UC
            o.T[j] = Ti
            self.f64[j] = dict(Ti=Ti, D=D, N=N, niter=nit, UC=UC, Sigma=N, kappa=np.full(self.m, kap),
                               relevant=rel)

    def _multi(self, exact_UC=True):
        o = self.outst
        cfg = o.blk.cfg
        A, mB, Cv = o.sysmata, o.mhalfb, o.outovlc
        nv, m, n = self.nv, self.m, self.n
        rel = self._relevant()
        di = np.diag_indices(n)
        for j in range(self.n_out):
            AA = A.copy()
            kappa_arr = self.kappaC_arr * Cv[j]
            Tpi = np.zeros((nv, m, n))
            nits = np.zeros((nv, m), dtype=np.int32)
            for p in range(nv):
                AA[di] += kappa_arr[p] - (kappa_arr[p - 1] if p > 0 else 0)
                Tpi[p] = R.iterative_wrapper(AA, np.ascontiguousarray(mB[j]), rel, cfg.iter_rtol,
                                             cfg.iter_max, nits[p])
            Epq = None
            if exact_UC:  # lakernel.py:709-714
                Epq = np.zeros((m, nv, nv))
                for p in range(nv):
                    ATp = Tpi[p] @ A
                    for q in range(p + 1):
                        Epq[:, q, p] = Epq[:, p, q] = np.einsum("ai,ai->a", ATp, Tpi[q])
            self._reduce_nodes(j, Tpi, kappa_arr, Cv[j], Epq)
            self.f64[j]["niter"] = nits
            self.f64[j]["relevant"] = rel


class EmpirKernel(_Kernel):
    """lakernel.py:747-805: no linear system -- T from the empirical weight max(rho_acc - distance, 0), row-normalised;
    U/C from the exact quadratic form E = T A T^T (skipped with no_qlt_ctrl)."""

    def _single(self):
        o = self.outst
        cfg = o.blk.cfg
        dy = o.yx_val[0].ravel()[:, None] - o.iny_val[None, :]
        dx = o.yx_val[1].ravel()[:, None] - o.inx_val[None, :]
        Ti = np.maximum(rho_acc(cfg) - np.hypot(dy, dx), 0)
        Ti /= np.sum(Ti, axis=-1)[:, None]
        if getattr(o, "no_qlt_ctrl", False):
            o.T[:, :, :] = Ti
            for j in range(self.n_out):
                self.f64[j] = dict(Ti=Ti)
            return
        A, mB, Cv = o.sysmata, o.mhalfb, o.outovlc
        for j in range(self.n_out):
            kap = self.kappaC_arr[0] * Cv[j]
            D = np.einsum("ai,ai->a", mB[j], Ti)
            N = np.einsum("ai,ai->a", Ti, Ti)
            E = np.einsum("ij,ai,aj->a", A, Ti, Ti)
            self.kappa_[j] = kap
            self.Sigma_[j] = N
            self.UC_[j] = 1.0 + (E - 2 * D) / Cv[j]
            o.T[j] = Ti
            self.f64[j] = dict(Ti=Ti, D=D, N=N, E=E, kappa=np.full(self.m, kap), Sigma=N, UC=1.0 + (E - 2 * D) / Cv[j])

    _multi = _single  # lakernel.py:801-805
