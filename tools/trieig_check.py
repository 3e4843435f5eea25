#!/usr/bin/env python
"""Stage-by-stage check of the tridiagonalisation-based eigensolver (csrc/trieig.cu) against NumPy / SciPy."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pyimcom_b200 import _lib  # noqa: E402
from pyimcom_b200 import lakernel as GL  # noqa: E402

ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731


def test_matrix(n, kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "random":
        G = rng.standard_normal((n, n))
        return (G + G.T) / 2
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.zeros(n // 5), np.logspace(-11, -1, n - n // 5)])
    A = (Q * lam) @ Q.T
    return (A + A.T) / 2


def check(n, kind):
    A = test_matrix(n, kind)
    npad = (n + 127) // 128 * 128
    Ad = torch.eye(npad, dtype=torch.float64, device="cuda")
    Ad[:n, :n] = torch.from_numpy(A).cuda()
    # 1. tridiagonalisation: eigenvalues of T == eigenvalues of A
    W = Ad.clone()
    d, e, tau = (torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(3))
    _lib.dev_tridiag(ptr(W), npad, n, ptr(d), ptr(e), ptr(tau), st())
    torch.cuda.synchronize()
    import scipy.linalg as sl

    lt = sl.eigvalsh_tridiagonal(d.cpu().numpy(), e.cpu().numpy()[: n - 1])
    la = np.linalg.eigvalsh(A)
    print(f"n={n} {kind}: tridiag eigenvalue error {np.abs(lt - la).max():.2e} (|A| = {np.abs(la).max():.2e})", flush=True)
    # 2. the whole solver
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lam, Vt, _ = GL.eigh_device(Ad.clone(), n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    lam, V = lam[:n].cpu().numpy(), Vt[:n, :n].cpu().numpy().T
    print(f"   eigh: {dt * 1e3:.1f} ms, |dlam| {np.abs(np.sort(lam) - la).max():.2e}, orth {np.abs(V.T @ V - np.eye(n)).max():.2e}, "
          f"resid {np.abs(A @ V - V * lam).max():.2e}, padding ok {bool(torch.equal(Vt[n:, n:], torch.eye(npad - n, dtype=torch.float64, device='cuda')))}",
          flush=True)


if __name__ == "__main__":
    torch.cuda.set_device(0)
    sizes = [(int(a), "graded") for a in sys.argv[1:]] or [(5, "random"), (130, "random"), (300, "graded"),
                                                           (1000, "graded"), (1532, "graded")]
    for n, kind in sizes:
        check(n, kind)
