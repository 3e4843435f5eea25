"""Linear-algebra kernels of the coaddition hot path on a B200: ``CholKernel``, ``EigenKernel``, ``IterKernel``.

Drop-in for the kernel-class seam of the reference (``OutStamp.LAKERNEL``, coadd.py:839-844; classes at
lakernel.py:141-744): ``Kernel(outst)()`` reads ``outst.sysmata (n,n) f64``, ``outst.mhalfb (n_out,m,n) f64``,
``outst.outovlc (n_out,)`` and ``outst.blk.cfg.{n_out,n2f,kappaC_arr,uctarget,sigmamax}`` (Iterative also
``instamp_pad, dtheta, iter_rtol, iter_max`` and ``outst.yx_val, iny_val, inx_val``) and writes
``outst.T (n_out,m,n) float32`` and ``outst.UC / Sigma / kappa (n_out,n2f,n2f) float32``.

All arithmetic runs in hand-written sm_100a kernels behind the C ABI (include/pyimcom_b200.h); torch is
used only to own device memory and the stream.  The same device pipeline (``solve_stamp``) serves the
device-resident block driver in ``pyimcom_b200.coadd``, which never round-trips through host memory.

There is no CPU fallback: without the CUDA library or a GPU these classes raise.
"""

from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib

NB = _lib.NB
ARCSEC = math.pi / 648000.0  # config.py:87


def rup(x: int, q: int = NB) -> int:
    return max(q, (int(x) + q - 1) // q * q)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream_handle():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_gpu():
    if not torch.cuda.is_available():
        raise RuntimeError("pyimcom_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def rho_acc(cfg) -> float:
    """Acceptance radius in output pixels (lakernel.py:618, coadd.py:923)."""
    return (cfg.instamp_pad / ARCSEC) / (cfg.dtheta * 3600.0)


@dataclass
class DeviceSystem:
    """Device-resident system matrices of one output stamp (padded to multiples of 128).

    A  (npad, npad) f64: in-in overlap, identity in the padding, NO kappa on the diagonal
    mB (n_out, mpad, npad) f64: -B/2, zero padded
    C  (n_out,) host floats
    px, py (n,) f64 input pixel positions and the output grid origin (IterKernel only)
    """

    n: int
    m: int
    n2f: int
    A: torch.Tensor
    mB: torch.Tensor
    C: np.ndarray
    px: torch.Tensor | None = None
    py: torch.Tensor | None = None
    outx: torch.Tensor | None = None  # (m,) f64 output pixel positions (IterKernel only)
    outy: torch.Tensor | None = None
    # Block driver only: A may be left unmaterialised (None); assemble(diag_add) then cuts A + diag_add*I straight
    # out of the cached InStamp-pair blocks, which saves the copy A -> W of the single-kappa Cholesky path.
    assemble: object = None

    @property
    def npad(self):
        return self.A.shape[0] if self.A is not None else self.mB.shape[2]

    def matrix(self) -> torch.Tensor:
        """A, materialised on first use."""
        if self.A is None:
            self.A = self.assemble(0.0)
        return self.A

    @property
    def mpad(self):
        return self.mB.shape[1]

    @property
    def n_out(self):
        return self.mB.shape[0]


@dataclass
class KernelOutput:
    """What a kernel hands to the T-apply stage for one output PSF.

    Tpi (nv, mpad, npad) f64 node solutions and w (m, nv) f64 node weights (None when nv == 1: Tpi[0] is T);
    kappa, Sigma, UC (m,) f64 or None when they are to be formed from finalize's D and N (single kappa:
    ``kappa_scalar`` set, ``E`` optionally the exact T A T^T)."""

    Tpi: torch.Tensor
    w: torch.Tensor | None
    kappa: torch.Tensor | None = None
    Sigma: torch.Tensor | None = None
    UC: torch.Tensor | None = None
    kappa_scalar: float | None = None
    E: torch.Tensor | None = None
    extras: dict = field(default_factory=dict)


def _f64(*shape):
    return torch.empty(shape, dtype=torch.float64, device="cuda")


def _zeros64(*shape):
    return torch.zeros(shape, dtype=torch.float64, device="cuda")


def _incs(vals):
    arr = (C.c_double * max(1, len(vals)))(*vals)
    return arr, len(vals)


# ------------------------------------------------------------------------------------------------------
# Cholesky: factor + solve of up to 16 padded systems per call
# ------------------------------------------------------------------------------------------------------
def chol_solve_batch(Ws, Xs, factor=True, solve=True, mrows=None):
    """In place: W_k -> L_k (and L_k^T above the diagonal blocks), X_k -> X_k (L_k L_k^T)^-1.

    mrows[k] (optional): number of real rows of X_k; the rows beyond it must be zero (they stay zero) and a last
    128-row tile with at most 64 real rows is then solved at half cost.
    Returns the device int32 tensor of LAPACK-style info codes (0 = success)."""
    nsys = len(Ws)
    assert 0 < nsys <= _lib.MAXB
    info = torch.zeros(nsys, dtype=torch.int32, device="cuda")
    # sliced INT8 (tcgen05) path for the long-K panel updates: needs a per-system scratch for the digit planes; only
    # worth it when the batch has more than one super-panel of block columns
    use_oz = OZAKI and factor and max(W.shape[0] for W in Ws) > OZAKI_MIN_N
    sysarr = (_lib.SolveSys * nsys)()
    keep = []
    for k in range(nsys):
        W, X = Ws[k], Xs[k] if Xs is not None else None
        npad = W.shape[0]
        Dinv = _f64(2 * (npad // NB), NB, NB)
        keep.append(Dinv)
        s = sysarr[k]
        s.W, s.Dinv = W.data_ptr(), Dinv.data_ptr()
        s.X = X.data_ptr() if X is not None else None
        s.info = info.data_ptr() + 4 * k
        s.npad, s.ldw = npad, W.stride(0)
        s.mpad = X.shape[0] if X is not None else 0
        s.ldx = X.stride(0) if X is not None else npad
        s.mrows = int(mrows[k]) if (mrows is not None and X is not None) else 0
        if use_oz:
            nbytes = int(_lib.lib.b200_chol_work_bytes(npad, s.mpad))
            work = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            keep.append(work)
            s.work, s.work_bytes = work.data_ptr(), nbytes
    _lib.dev_chol_solve(sysarr, nsys, int(factor), int(solve and Xs is not None), stream_handle())
    return info, keep


def eigh_device_batch(items):
    """np.linalg.eigh replacement for several matrices at once: items = [(A_pad, n), ...] (A_pad destroyed, identity
    padded).  Returns [(lam (npad,), Vt (npad, npad) with eigenvectors as rows)], and the sweep count."""
    out, sweeps_max = [], 0
    for c0 in range(0, len(items), _lib.MAXB):
        chunk = items[c0:c0 + _lib.MAXB]
        pr = (_lib.EighProblem * len(chunk))()
        res = []
        for k, (A_pad, n) in enumerate(chunk):
            npad = A_pad.shape[0]
            Vt = _zeros64(npad, npad)
            lam = _zeros64(npad)
            pr[k].A, pr[k].Vt, pr[k].lam = A_pad.data_ptr(), Vt.data_ptr(), lam.data_ptr()
            pr[k].lda, pr[k].ldv, pr[k].n = A_pad.stride(0), Vt.stride(0), int(n)
            res.append((lam, Vt))
        sweeps = C.c_int(0)
        _lib.dev_eigh_batch(pr, len(chunk), 80, C.byref(sweeps), stream_handle())
        sweeps_max = max(sweeps_max, sweeps.value)
        out += res
    return out, sweeps_max


def eigh_device(A_pad, n):
    """np.linalg.eigh replacement: returns (lam (n,), Vt (npad, npad) with eigenvectors as rows). A_pad is destroyed."""
    ((lam, Vt),), sweeps = eigh_device_batch([(A_pad, n)])
    return lam, Vt, sweeps


def _padded_system(ds: DeviceSystem, incs):
    if ds.A is None and ds.assemble is not None and len(incs) <= 1:
        return ds.assemble(float(incs[0]) if incs else 0.0)  # same single rounding A_ii + inc as k_pad_system
    ds.matrix()
    W = _f64(ds.npad, ds.npad)
    arr, k = _incs(incs)
    _lib.dev_pad_system(ptr(W), W.stride(0), ds.n, ds.npad, ptr(ds.A), ds.A.stride(0), arr, k, stream_handle())
    return W


OZAKI = os.environ.get("B200_OZAKI", "1") != "0"  # long-K panel updates on the INT8 tcgen05 tensor cores (csrc/ozaki.cu)
OZAKI_MIN_N = 512  # (the library itself only switches over when there is more than one super-panel)
SOLVE_STREAMS = int(os.environ.get("B200_SOLVE_STREAMS", "2"))  # concurrent groups of systems in the batched factorisation (1 = everything on the caller's stream)
# Equal priorities: the groups' kernels interleave CTA by CTA.  (1: staggered priorities, the first group running as if
# alone and the others filling what it leaves -- better with the all-DMMA kernels of round 1, worse now: 64-stamp block
# 275.3 vs 266.4 ms, 256 stamps 1076.6 vs 1056.5 ms)
STREAM_PRIORITIES = os.environ.get("B200_STREAM_PRIORITIES", "0") != "0"
_SIDE = {}


def _side_streams(n):
    dev = torch.cuda.current_device()
    have = _SIDE.setdefault(dev, [])
    while len(have) < n:
        # (B200_STREAM_PRIORITIES=1: staggered priorities, lower number = higher priority)
        prio = -len(have) if STREAM_PRIORITIES else 0
        have.append(torch.cuda.Stream(priority=prio))
    return have[:n]


def side_streams_in_use():
    """The solve streams created so far on the current device."""
    return list(_SIDE.get(torch.cuda.current_device(), []))


def _chol_launch(items):
    """Enqueue factor + solve of A_k + sum(incs_k) for every item (ds, incs, j_out): the systems go through the
    batched factorisation together (up to MAXB per launch sequence, independent groups on separate streams).
    Nothing is synchronised; returns the handle _chol_finish completes."""
    n_items = len(items)
    Xs = [None] * n_items
    infos, events = [], []
    cur = torch.cuda.current_stream()
    nstream = min(SOLVE_STREAMS, n_items) if n_items > 1 else 1
    if nstream <= 1:
        chunks = [(c0, min(c0 + _lib.MAXB, n_items), cur) for c0 in range(0, n_items, _lib.MAXB)]
        pool = []
    else:
        # independent groups of systems on separate streams: the serial diagonal-block factorisations and the partial
        # last waves of one group's tile kernels are filled by the other group's work
        per = min(_lib.MAXB, -(-n_items // nstream))
        pool = _side_streams(nstream)
        chunks = [(c0, min(c0 + per, n_items), pool[q % nstream]) for q, c0 in enumerate(range(0, n_items, per))]
    # right-hand sides are cloned on the caller's stream (whose allocator pool they return to after the T-apply stage);
    # the solve streams only own what they allocate and free themselves (W, the inverted diagonal blocks)
    # (consecutive items of one stamp and output PSF -- its kappa nodes -- share one (nv, mpad, npad) buffer, which is
    # later handed to the node reduction as it is: no torch.stack copy of nv x 72 MB)
    bufs = {}
    k = 0
    while k < n_items:
        ds, _, j = items[k]
        k1 = k + 1
        while k1 < n_items and items[k1][0] is ds and items[k1][2] == j:
            k1 += 1
        buf = ds.mB[j].unsqueeze(0).repeat(k1 - k, 1, 1)
        bufs[k] = buf
        for p in range(k, k1):
            Xs[p] = buf[p - k]
        k = k1
    for st in pool:
        st.wait_stream(cur)
    for c0, c1, st in chunks:
        with torch.cuda.stream(st):
            Ws = [_padded_system(ds, incs) for ds, incs, _ in items[c0:c1]]
            info, _keep = chol_solve_batch(Ws, Xs[c0:c1], mrows=[it[0].m for it in items[c0:c1]])
            if st is not cur:
                info.record_stream(cur)
            infos.append(info)
            del Ws
    for st in pool:  # the caller's stream will wait for exactly this work, not for whatever is enqueued later
        ev = torch.cuda.Event()
        ev.record(st)
        events.append(ev)
    return dict(items=items, Xs=Xs, infos=infos, events=events, bufs=bufs)


def _chol_finish(h):
    """Join the solve streams, read the LAPACK-style info codes back once, and send failing items through the
    repair branch of CholKernel._cholesky_wrapper (lakernel.py:262-279).  Returns the solutions X_k (mpad, npad)."""
    items, Xs = h["items"], h["Xs"]
    cur = torch.cuda.current_stream()
    for ev in h["events"]:
        cur.wait_event(ev)
    bad = torch.cat(h["infos"]).cpu().numpy()
    if bad.any():
        shifts = {}
        for k in np.nonzero(bad)[0]:
            ds, incs, j = items[k]
            if id(ds) not in shifts:  # one eigh per stamp serves all of its failing nodes
                lam, _, _ = eigh_device(ds.matrix().clone(), ds.n)
                shifts[id(ds)] = float(lam[: ds.n].min().item())
            w0 = shifts[id(ds)]
            warnings.warn(f"CholKernel: repaired negative eigenvalue {w0:19.12e}", stacklevel=3)
            W = _padded_system(ds, list(incs) + [abs(w0) + 1e-16])
            Xs[k].copy_(ds.mB[j])  # (in place: Xs[k] may be a slice of the stamp's node buffer)
            info2, _keep2 = chol_solve_batch([W], [Xs[k]], mrows=[ds.m])
            if int(info2.item()) != 0:
                raise np.linalg.LinAlgError("Cholesky failed after the eigenvalue repair")
    return Xs


def _chol_solve_items(items):
    """Factor + solve for every item (ds, incs, j_out), with the repair branch; returns the list of solutions."""
    return _chol_finish(_chol_launch(items))


def _chol_with_repair(ds: DeviceSystem, inc_lists, j_out):
    """Single-stamp form of _chol_solve_items: the systems A + sum(incs) for every increment list."""
    return _chol_solve_items([(ds, incs, j_out) for incs in inc_lists])


def _node_reduce(ds, j_out, Tpi, kappa_arr, kappaC_arr, ucmin, smax, Epq_in=None):
    """Shared tail of the multi-kappa kernels (lakernel.py:361-393, 703-741)."""
    nv, m = Tpi.shape[0], ds.m
    Cj = float(ds.C[j_out])
    Dp, DpC = _f64(m, nv), _f64(m, nv)
    Npq, Epq, EpqC = _f64(m, nv, nv), _f64(m, nv, nv), _f64(m, nv, nv)
    kn = (C.c_double * 16)(*([float(k) for k in kappa_arr] + [0.0] * (16 - nv)))
    st = stream_handle()
    mB = ds.mB[j_out]
    _lib.dev_node_stats(ptr(mB), mB.stride(0), ptr(Tpi), Tpi.stride(1), Tpi.stride(0), nv, m, ds.n, kn, Cj, ptr(Dp),
                        ptr(Npq), ptr(Epq), ptr(DpC), ptr(EpqC), ptr(Epq_in), st)
    kap = torch.tensor(np.asarray(kappaC_arr, dtype=np.float64), device="cuda")
    ok, oS, oU, ow = _f64(m), _f64(m), _f64(m), _f64(m, nv)
    iv = torch.empty(m, dtype=torch.int32, device="cuda")
    br = torch.empty(m, dtype=torch.int32, device="cuda")
    _lib.dev_build_reduced_T(ptr(Npq), ptr(DpC), ptr(EpqC), ptr(kap), nv, m, float(ucmin), float(smax), ptr(ok),
                             ptr(oS), ptr(oU), ptr(ow), ptr(iv), ptr(br), st)
    kappa = _f64(m)
    _lib.dev_scale(ptr(ok), Cj, m, ptr(kappa), st)  # kappa = out_kappa * C (lakernel.py:390)
    return KernelOutput(Tpi=Tpi, w=ow, kappa=kappa, Sigma=oS, UC=oU,
                        extras=dict(Dp=Dp, Npq=Npq, Epq=Epq, out_w=ow, iv=iv, branch=br))


def solve_chol_launch(dss, cfg, j_out: int):
    """CholKernel for one output PSF of several output stamps at once (lakernel.py:281-394): enqueue the nv systems of
    every stamp into one batched factorisation.  Returns the handle solve_chol_finish turns into KernelOutputs."""
    kappaC = np.asarray(cfg.kappaC_arr, dtype=np.float64)
    nv = kappaC.size
    items = []
    for ds in dss:
        kappa_arr = kappaC * float(ds.C[j_out])
        if nv == 1:
            items.append((ds, [float(kappa_arr[0])] if kappa_arr[0] else [], j_out))
            continue
        run = []
        for p in range(nv):  # cumulative diagonal increments (lakernel.py:356)
            run.append(float(kappa_arr[p] - (kappa_arr[p - 1] if p > 0 else 0)))
            items.append((ds, list(run), j_out))
    return dict(h=_chol_launch(items), dss=list(dss), cfg=cfg, j_out=j_out) if items else None


def solve_chol_finish(handle):
    if handle is None:
        return []
    dss, cfg, j_out = handle["dss"], handle["cfg"], handle["j_out"]
    kappaC = np.asarray(cfg.kappaC_arr, dtype=np.float64)
    nv = kappaC.size
    Xs = _chol_finish(handle["h"])
    bufs = handle["h"]["bufs"]
    outs = []
    for q, ds in enumerate(dss):
        kappa_arr = kappaC * float(ds.C[j_out])
        if nv == 1:
            outs.append(KernelOutput(Tpi=Xs[q].unsqueeze(0), w=None, kappa_scalar=float(kappa_arr[0])))
        else:
            Tpi = bufs[q * nv]  # the nv node solutions of this stamp, already contiguous
            assert Tpi.shape[0] == nv
            outs.append(_node_reduce(ds, j_out, Tpi, kappa_arr, kappaC, cfg.uctarget, cfg.sigmamax))
    return outs


def solve_chol_batch(dss, cfg, j_out: int):
    """solve_chol_launch + solve_chol_finish.  Returns one KernelOutput per stamp."""
    return solve_chol_finish(solve_chol_launch(dss, cfg, j_out))


def solve_chol(ds: DeviceSystem, cfg, j_out: int) -> KernelOutput:
    """CholKernel for one output PSF (lakernel.py:281-394)."""
    return solve_chol_batch([ds], cfg, j_out)[0]


def eigen_decompose_batch(dss):
    """eigh(A) once per stamp (lakernel.py:162, 201), all stamps of a batch together: returns per stamp
    (lam (npad,), Q^T (rows = eigenvectors), Q, sweeps)."""
    res, sweeps = eigh_device_batch([(ds.matrix().clone(), ds.n) for ds in dss])
    out = []
    for ds, (lam, Vt) in zip(dss, res):
        Q = _f64(ds.npad, ds.npad)
        _lib.dev_transpose(ptr(Vt), Vt.stride(0), ptr(Q), Q.stride(0), ds.npad, ds.npad, stream_handle())
        out.append((lam, Vt, Q, sweeps))
    return out


def eigen_decompose(ds: DeviceSystem):
    """Single-stamp form of eigen_decompose_batch."""
    return eigen_decompose_batch([ds])[0]


def solve_eigen(ds: DeviceSystem, cfg, j_out: int, eig=None, nbis: int = 13) -> KernelOutput:
    """EigenKernel for one output PSF (lakernel.py:154-223)."""
    if eig is None:
        eig = eigen_decompose(ds)
    lam, Vt, Q, sweeps = eig
    st = stream_handle()
    m, n, mpad, npad = ds.m, ds.n, ds.mpad, ds.npad
    kappaC = np.asarray(cfg.kappaC_arr, dtype=np.float64)
    Cj = float(ds.C[j_out])
    mB = ds.mB[j_out]
    # mPhalf = mBhalf @ Q : P[a,k] = sum_i mB[a,i] Vt[k,i]
    P = _f64(mpad, npad)
    _lib.dev_gemm_nt(ptr(mB), mB.stride(0), ptr(Vt), Vt.stride(0), ptr(P), P.stride(0), mpad, npad, npad, 0, st)
    tt = _zeros64(mpad, npad)
    kappa, Sigma, UC = _f64(m), _f64(m), _f64(m)
    if kappaC.size == 1:
        kap = float(kappaC[0] * Cj)
        _lib.dev_eigen_single(ptr(lam), ptr(P), P.stride(0), m, n, Cj, kap, ptr(Sigma), ptr(UC), ptr(tt), tt.stride(0),
                              st)
        kappa.fill_(kap)
    else:
        k0 = _f64(m)
        _lib.dev_lakernel1(ptr(lam), ptr(P), P.stride(0), m, n, Cj, float(cfg.uctarget), float(kappaC[0] * Cj),
                           float(kappaC[-1] * Cj), nbis, ptr(k0), ptr(Sigma), ptr(UC), ptr(tt), tt.stride(0),
                           float(cfg.sigmamax), st)
        # the reference stores kappa into a float32 array and multiplies by C once more (lakernel.py:216, 222)
        kappa = k0.float().double() * Cj
    # T = tt @ Q^T : T[a,i] = sum_k tt[a,k] Q[i,k]
    T = _f64(mpad, npad)
    _lib.dev_gemm_nt(ptr(tt), tt.stride(0), ptr(Q), Q.stride(0), ptr(T), T.stride(0), mpad, npad, npad, 0, st)
    return KernelOutput(Tpi=T.unsqueeze(0), w=None, kappa=kappa, Sigma=Sigma, UC=UC,
                        extras=dict(lam=lam, sweeps=sweeps))


ITER_NCAP = 6100  # accepted input pixels per output pixel the CG kernel holds (csrc/iter.cu)


def solve_iter(ds: DeviceSystem, cfg, j_out: int, exact_UC=None) -> KernelOutput:
    """IterKernel for one output PSF (lakernel.py:592-744)."""
    assert ds.px is not None and ds.py is not None, "IterKernel needs the input pixel positions"
    st = stream_handle()
    m, n, mpad, npad = ds.m, ds.n, ds.mpad, ds.npad
    kappaC = np.asarray(cfg.kappaC_arr, dtype=np.float64)
    nv = kappaC.size
    Cj = float(ds.C[j_out])
    kappa_arr = kappaC * Cj
    mB = ds.mB[j_out]
    rho = rho_acc(cfg)
    Tpi = _zeros64(nv, mpad, npad)
    niter = torch.zeros((nv, m), dtype=torch.int32, device="cuda")
    nsel = torch.zeros((nv, m), dtype=torch.int32, device="cuda")
    run = []
    for p in range(nv):
        inc = float(kappa_arr[p] - (kappa_arr[p - 1] if p > 0 else 0))
        if nv > 1 or inc:  # single kappa: "if my_kappa: AA[di] += my_kappa" (lakernel.py:628-629)
            run.append(inc)
        AA = _padded_system(ds, run)
        _lib.dev_iter_cg(ptr(AA), AA.stride(0), 0.0, ptr(mB), mB.stride(0), m, n, ptr(ds.px), ptr(ds.py), ptr(ds.outx),
                         ptr(ds.outy), float(rho), float(cfg.iter_rtol), int(cfg.iter_max),
                         ptr(Tpi[p]), Tpi.stride(1), ptr(niter[p]), ptr(nsel[p]), st)
    if n > ITER_NCAP and int(niter.min().item()) < 0:
        raise _lib.B200Error(f"IterKernel: an output pixel accepts more than {ITER_NCAP} input pixels within rho_acc = "
                             f"{rho:.3f} (n = {n}); the conjugate-gradient kernel keeps its vectors in shared memory")
    if exact_UC is None:
        exact_UC = nv > 1  # defaults of the reference (lakernel.py:592, 656)

    def exact_E(p, q, out, ostride):
        ATp = _f64(mpad, npad)  # (Tpi[p] @ A): A is symmetric, so the NT product with A's rows is the same
        _lib.dev_gemm_nt(ptr(Tpi[p]), Tpi.stride(1), ptr(ds.matrix()), ds.matrix().stride(0), ptr(ATp), ATp.stride(0), mpad, npad, npad,
                         0, st)
        return ATp

    if nv == 1:
        E = None
        if exact_UC:
            E = _f64(m)
            ATp = exact_E(0, 0, None, 0)
            _lib.dev_rowdot(ptr(ATp), ATp.stride(0), ptr(Tpi[0]), Tpi.stride(1), m, n, ptr(E), 1, st)
        return KernelOutput(Tpi=Tpi, w=None, kappa_scalar=float(kappa_arr[0]), E=E,
                            extras=dict(niter=niter, nsel=nsel))
    Epq_in = None
    if exact_UC:  # lakernel.py:709-716
        Epq_in = _f64(m, nv, nv)
        for p in range(nv):
            ATp = exact_E(p, p, None, 0)
            for q in range(p + 1):
                _lib.dev_rowdot(ptr(ATp), ATp.stride(0), ptr(Tpi[q]), Tpi.stride(1), m, n,
                                C.c_void_p(Epq_in.data_ptr() + 8 * (p * nv + q)), nv * nv, st)
                if q != p:
                    _lib.dev_rowdot(ptr(ATp), ATp.stride(0), ptr(Tpi[q]), Tpi.stride(1), m, n,
                                    C.c_void_p(Epq_in.data_ptr() + 8 * (q * nv + p)), nv * nv, st)
    out = _node_reduce(ds, j_out, Tpi, kappa_arr, kappaC, cfg.uctarget, cfg.sigmamax, Epq_in)
    out.extras.update(niter=niter, nsel=nsel)
    return out


def solve_empir(ds: DeviceSystem, cfg, j_out: int, no_qlt_ctrl: bool = False) -> KernelOutput:
    """EmpirKernel for one output PSF (lakernel.py:747-805): no linear system; U/C from the exact E = T A T^T."""
    assert ds.px is not None and ds.outx is not None, "EmpirKernel needs the pixel positions"
    st = stream_handle()
    m, n, mpad, npad = ds.m, ds.n, ds.mpad, ds.npad
    T = _f64(mpad, npad)
    _lib.dev_empir_T(ptr(ds.px), ptr(ds.py), ptr(ds.outx), ptr(ds.outy), m, mpad, n, npad, float(rho_acc(cfg)), ptr(T),
                     T.stride(0), st)
    if no_qlt_ctrl:  # the reference leaves kappa, Sigma, U/C at their zero initialisation (lakernel.py:770-774)
        z = _zeros64(m)
        return KernelOutput(Tpi=T.unsqueeze(0), w=None, kappa=z, Sigma=z.clone(), UC=z.clone())
    A = ds.matrix()
    AT = _f64(mpad, npad)  # T @ A (A symmetric: the NT product with A's rows)
    _lib.dev_gemm_nt(ptr(T), T.stride(0), ptr(A), A.stride(0), ptr(AT), AT.stride(0), mpad, npad, npad, 0, st)
    E = _f64(m)
    _lib.dev_rowdot(ptr(AT), AT.stride(0), ptr(T), T.stride(0), m, n, ptr(E), 1, st)
    kap = float(np.asarray(cfg.kappaC_arr, dtype=np.float64)[0] * float(ds.C[j_out]))
    return KernelOutput(Tpi=T.unsqueeze(0), w=None, kappa_scalar=kap, E=E)


SOLVERS = {"Cholesky": solve_chol, "Eigen": solve_eigen, "Iterative": solve_iter, "Empirical": solve_empir}


# ------------------------------------------------------------------------------------------------------
# T-apply stage shared by the kernel classes and the block driver
# ------------------------------------------------------------------------------------------------------
def trapezoid_weights(fade_kernel: int) -> np.ndarray:
    """coadd.py:1269-1271."""
    fk2 = 2 * fade_kernel
    s = np.arange(1, fk2 + 1, dtype=np.float64) / (fk2 + 1)
    s -= np.sin(2 * np.pi * s) / (2 * np.pi)
    return s


@dataclass
class ApplySpec:
    """Inputs of the T-apply stage that do not depend on the kernel."""

    fade: int = 0
    fade_w: torch.Tensor | None = None  # (2*fade,) f64 device
    indata: torch.Tensor | None = None  # (n_inframe, npad) f32 device
    seg_end: torch.Tensor | None = None  # (nseg,) int32 device
    seg_img: torch.Tensor | None = None
    n_img: int = 0
    n2: int = 1
    clamp_iter: bool = False
    want_T32: bool = True
    want_Ti64: bool = False


FINALIZE_MAX_LAYERS = 16  # input layers per k_finalize launch (two DMMA passes of 8)


def apply_T(ds: DeviceSystem, ko: KernelOutput, j_out: int, spec: ApplySpec):
    """finalize + maps for one output PSF: returns a dict of device tensors.

    T32 (m, npad) f32 faded; outimage (n_inframe, m) f32; kappa/Sigma/UC (m,) f32 faded;
    Tsum_stamp (n_img,), Tsum_inpix (m,), Neff (m,) f64; D, N (m,) f64."""
    st = stream_handle()
    m, n, npad = ds.m, ds.n, ds.npad
    nfr = spec.indata.shape[0] if spec.indata is not None else 0
    nseg = spec.seg_end.numel() if spec.seg_end is not None else 0
    out = {}
    # (k_finalize writes columns < n only: the padding columns n..npad-1 of the kept copies are zero by construction)
    T32 = torch.zeros((m, npad), dtype=torch.float32, device="cuda") if spec.want_T32 else None
    Ti64 = _zeros64(m, npad) if spec.want_Ti64 else None
    D, N = _f64(m), _f64(m)
    outimage = torch.zeros((max(nfr, 1), m), dtype=torch.float32, device="cuda")
    Tsum_image = _zeros64(m, max(spec.n_img, 1))
    mB = ds.mB[j_out]
    a = _lib.FinalizeArgs()
    a.Tpi, a.strideT, a.ldt = ko.Tpi.data_ptr(), ko.Tpi.stride(0), ko.Tpi.stride(1)
    a.w = ko.w.data_ptr() if ko.w is not None else None
    a.nv = ko.Tpi.shape[0] if ko.w is not None else 1
    a.mB, a.ldb = mB.data_ptr(), mB.stride(0)
    a.m, a.n, a.n2f, a.fade = m, n, ds.n2f, spec.fade
    a.fade_w = spec.fade_w.data_ptr() if spec.fade_w is not None else None
    a.indata = spec.indata.data_ptr() if spec.indata is not None else None
    a.ldi = spec.indata.stride(0) if spec.indata is not None else 0
    a.n_inframe = nfr
    a.seg_end = spec.seg_end.data_ptr() if nseg else None
    a.seg_img = spec.seg_img.data_ptr() if nseg else None
    a.nseg, a.n_img = nseg, spec.n_img
    a.T32, a.ldt32 = (T32.data_ptr(), T32.stride(0)) if T32 is not None else (None, 0)
    a.Ti64, a.ldt64 = (Ti64.data_ptr(), Ti64.stride(0)) if Ti64 is not None else (None, 0)
    a.D, a.N = D.data_ptr(), N.data_ptr()
    a.outimage = outimage.data_ptr()
    a.Tsum_image = Tsum_image.data_ptr()
    a.n_inframe = min(nfr, FINALIZE_MAX_LAYERS)
    _lib.dev_finalize(C.byref(a), st)
    # more input layers than one launch takes (16): the remaining ones in further passes over the node solutions, which
    # write the same T / D / N / Tsum again and the coadded images of their own layers
    for f0 in range(FINALIZE_MAX_LAYERS, nfr, FINALIZE_MAX_LAYERS):
        a.T32, a.ldt32, a.Ti64, a.ldt64 = None, 0, None, 0
        a.indata = spec.indata.data_ptr() + 4 * f0 * spec.indata.stride(0)
        a.outimage = outimage.data_ptr() + 4 * f0 * m
        a.n_inframe = min(nfr - f0, FINALIZE_MAX_LAYERS)
        _lib.dev_finalize(C.byref(a), st)
    kappa, Sigma, UC = ko.kappa, ko.Sigma, ko.UC
    if kappa is None:  # single kappa: maps from D and N (lakernel.py:312-316, 643-648)
        kappa, Sigma, UC = _f64(m), _f64(m), _f64(m)
        _lib.dev_single_kappa_maps(ptr(D), ptr(N), ptr(ko.E), m, float(ko.kappa_scalar), float(ds.C[j_out]), ptr(kappa),
                                   ptr(Sigma), ptr(UC), st)
    k32, S32, U32 = (torch.empty(m, dtype=torch.float32, device="cuda") for _ in range(3))
    Tsum_stamp, Tsum_inpix, Neff = _f64(max(spec.n_img, 1)), _f64(m), _f64(m)
    _lib.dev_stamp_maps(ptr(kappa), ptr(Sigma), ptr(UC), m, ds.n2f, spec.fade, int(spec.clamp_iter), ptr(spec.fade_w),
                        ptr(k32), ptr(S32), ptr(U32), ptr(Tsum_image) if spec.n_img else None, spec.n_img, spec.n2,
                        ptr(Tsum_stamp), ptr(Tsum_inpix), ptr(Neff), st)
    out.update(T32=T32, Ti64=Ti64, D=D, N=N, outimage=outimage[:nfr], kappa=k32, Sigma=S32, UC=U32,
               kappa64=kappa, Sigma64=Sigma, UC64=UC, Tsum_stamp=Tsum_stamp, Tsum_inpix=Tsum_inpix, Neff=Neff,
               Tsum_image=Tsum_image)
    return out


# ------------------------------------------------------------------------------------------------------
# Kernel-class seam (host NumPy in / out), lakernel.py:50-138
# ------------------------------------------------------------------------------------------------------
def upload_system(A, mBhalf, Cvec, n2f, px=None, py=None, outx=None, outy=None) -> DeviceSystem:
    """Host (n,n), (n_out,m,n) -> padded device system."""
    _need_gpu()
    n = A.shape[0]
    n_out, m, _ = mBhalf.shape
    npad, mpad = rup(n), rup(m)
    Ad = torch.eye(npad, dtype=torch.float64, device="cuda")
    Ad[:n, :n] = torch.from_numpy(np.ascontiguousarray(A, dtype=np.float64)).cuda()
    Bd = _zeros64(n_out, mpad, npad)
    Bd[:, :m, :n] = torch.from_numpy(np.ascontiguousarray(mBhalf, dtype=np.float64)).cuda()
    ds = DeviceSystem(n=n, m=m, n2f=n2f, A=Ad, mB=Bd, C=np.asarray(Cvec, dtype=np.float64))
    if px is not None:
        up = lambda v: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).ravel()).cuda()  # noqa: E731
        ds.px, ds.py, ds.outx, ds.outy = up(px), up(py), up(outx), up(outy)
    return ds


class _LAKernel:
    """Abstract base (lakernel.py:50-138)."""

    KIND = None

    def __init__(self, outst) -> None:
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
        self.f64 = {}  # per output PSF: float64 intermediates (device copies brought to host) for parity tests

    def _device_system(self) -> DeviceSystem:
        o = self.outst
        return upload_system(o.sysmata, o.mhalfb, o.outovlc, self.n2f)

    def _solve(self, ds, j_out):
        return SOLVERS[self.KIND](ds, self.outst.blk.cfg, j_out)

    def __call__(self, keep_f64: bool = False) -> None:
        _need_gpu()
        o = self.outst
        shape = (self.n_out, self.n2f, self.n2f)
        if self.n == 0:  # lakernel.py:110-119
            o.T = np.zeros((self.n_out, self.m, 0), dtype=np.float32)
            o.UC = np.ones(shape, dtype=np.float32)
            o.Sigma = np.zeros(shape, dtype=np.float32)
            o.kappa = np.ones(shape, dtype=np.float32)
            return
        ds = self._device_system()
        o.T = np.zeros((self.n_out, self.m, self.n), dtype=np.float32)
        UC_ = np.zeros((self.n_out, self.m), dtype=np.float32)
        Sigma_ = np.zeros((self.n_out, self.m), dtype=np.float32)
        kappa_ = np.zeros((self.n_out, self.m), dtype=np.float32)
        spec = ApplySpec(fade=0, want_T32=True, want_Ti64=keep_f64)
        for j in range(self.n_out):
            ko = self._solve(ds, j)
            res = apply_T(ds, ko, j, spec)
            o.T[j] = res["T32"][:, : self.n].cpu().numpy()
            UC_[j] = res["UC"].cpu().numpy()
            Sigma_[j] = res["Sigma"].cpu().numpy()
            kappa_[j] = res["kappa"].cpu().numpy()
            if keep_f64:
                d = dict(Ti=res["Ti64"][:, : self.n].cpu().numpy(), D=res["D"].cpu().numpy(), N=res["N"].cpu().numpy(),
                         kappa=res["kappa64"].cpu().numpy(), Sigma=res["Sigma64"].cpu().numpy(),
                         UC=res["UC64"].cpu().numpy())
                for k, v in ko.extras.items():
                    d[k] = v.cpu().numpy() if torch.is_tensor(v) else v
                self.f64[j] = d
        o.UC = UC_.reshape(shape)
        o.Sigma = Sigma_.reshape(shape)
        o.kappa = kappa_.reshape(shape)


class CholKernel(_LAKernel):
    """lakernel.py:226-394: batched FP64 Cholesky + triangular solves on the DMMA tensor pipe."""

    KIND = "Cholesky"


class EigenKernel(_LAKernel):
    """lakernel.py:141-223: Jacobi eigendecomposition + per-output-pixel kappa bisection."""

    KIND = "Eigen"

    def __call__(self, keep_f64: bool = False) -> None:
        self._eig = None
        super().__call__(keep_f64)

    def _solve(self, ds, j_out):
        if self._eig is None:  # one decomposition serves every output PSF (lakernel.py:162, 201)
            self._eig = eigen_decompose(ds)
        return solve_eigen(ds, self.outst.blk.cfg, j_out, eig=self._eig)


class EmpirKernel(_LAKernel):
    """lakernel.py:747-805: empirical weights instead of a solve (fast approximation)."""

    KIND = "Empirical"

    def _device_system(self) -> DeviceSystem:
        o = self.outst
        if getattr(o, "no_qlt_ctrl", False):  # no system matrices exist in this mode (coadd.py:1020-1025)
            n, m = self.n, self.m
            A = np.eye(n)
            mB = np.zeros((self.n_out, m, n))
            Cv = np.ones(self.n_out)
        else:
            A, mB, Cv = o.sysmata, o.mhalfb, o.outovlc
        return upload_system(A, mB, Cv, self.n2f, px=o.inx_val, py=o.iny_val, outx=np.asarray(o.yx_val[1]),
                             outy=np.asarray(o.yx_val[0]))

    def _solve(self, ds, j_out):
        return solve_empir(ds, self.outst.blk.cfg, j_out, no_qlt_ctrl=bool(getattr(self.outst, "no_qlt_ctrl", False)))


class IterKernel(_LAKernel):
    """lakernel.py:533-744: per-output-pixel conjugate gradient on the accepted sub-system."""

    KIND = "Iterative"

    def _device_system(self) -> DeviceSystem:
        o = self.outst  # lakernel.py:615-617: output positions yx_val, input positions iny_val/inx_val
        return upload_system(o.sysmata, o.mhalfb, o.outovlc, self.n2f, px=o.inx_val, py=o.iny_val,
                             outx=np.asarray(o.yx_val[1]), outy=np.asarray(o.yx_val[0]))
