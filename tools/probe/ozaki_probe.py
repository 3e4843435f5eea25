"""Feasibility of an integer-sliced (Ozaki-style) FP64 update for the two long-K kernels -- CPU experiment, no GPU.

Question for the next round: the Cholesky path now runs at 95 % of the DMMA rate (36 TFLOP/s); the only larger lever
left on a B200 is arithmetic -- the tcgen05 INT8 path is nominally 4.5 Pop/s dense, > 100x the FP64 tensor rate.  In
the Ozaki scheme each operand row is scaled by a power of two and cut into `s` signed slices of `b` bits; every slice
product is an exact INT8 x INT8 -> INT32 GEMM (b = 6: K * 63^2 < 2^31 up to K = 540 k), and
the s(s+1)/2 leading slice pairs are recombined in FP64.  What it costs in accuracy on OUR matrices decides `s`, and
`s` decides whether it pays: the break-even against DMMA is roughly s(s+1)/2 < (INT8 rate) / (FP64 rate) ~ 60-80.

This script factors A + kappa I of a real (synthetic-block) stamp with the same left-looking super-panel schedule as
csrc/linalg.cu, but with the super-panel update  W[i][j] -= W[i][0:c0] W[j][0:c0]^T  and the corresponding updates of
the right-hand-side rows evaluated by the sliced product (emulated exactly in float64: slice values are small
integers, so a float64 GEMM of two slices is exact), everything else in plain float64, and reports the error of
T = mBhalf (A + kappa I)^-1 against the all-float64 solution for s = 3 .. 9 -- to be read against the P-f64 bar (1e-9).

    python tools/probe/ozaki_probe.py            # config-1 sized stamp (n ~ 1.5 k), a few seconds per s
"""
import os
import sys
import time

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import routines as R  # noqa: E402
from oracle.sysmat import OracleOutStamp  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402
from pyimcom_b200.synth import StampConfig, SynthBlock  # noqa: E402

BITS = int(os.environ.get("OZAKI_BITS", "6"))  # bits per slice: |slice values| <= 2^(BITS-1) + 1, INT8 holds BITS <= 7


def slices_of(M, s):
    """Row-scaled slices: M[r, :] ~ 2^e[r] * sum_k S_k[r, :] 2^(-BITS (k+1)), S_k integer-valued with |S_k| <= 2^BITS."""
    amax = np.abs(M).max(axis=1)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)
    rem = M / np.exp2(e)[:, None]  # |rem| <= 1/2
    out = []
    for k in range(s):
        q = np.rint(rem * 2.0 ** (BITS * (k + 1)))  # integer, |q| <= 2^(BITS-1) + carry
        out.append(q)
        rem = rem - q / 2.0 ** (BITS * (k + 1))
    return e, out


def sliced_product(Am, Bm, s):
    """Am Bm^T from s slices per operand, keeping the slice pairs (k, l) with k + l < s (exact integer GEMMs)."""
    ea, Sa = slices_of(Am, s)
    eb, Sb = slices_of(Bm, s)
    acc = np.zeros((Am.shape[0], Bm.shape[0]))
    for d in range(s - 1, -1, -1):  # smallest terms first
        part = np.zeros_like(acc)
        for k in range(d + 1):
            part += Sa[k] @ Sb[d - k].T  # exact: |entries| <= K * 2^(2 BITS) < 2^53
        acc += part * 2.0 ** (-BITS * (d + 2))
    return acc * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]


def solve_sliced(W, X, s, nb=128, sp=4):
    """Left-looking super-panel Cholesky + forward solve of the rows X (as csrc/linalg.cu), sliced long-K updates."""
    n = W.shape[0]
    L = np.tril(W).copy()
    Z = X.copy()
    for c0 in range(0, n, nb * sp):
        c1 = min(c0 + nb * sp, n)
        if c0 > 0:  # super-panel update with everything to the left (the long-K kernel)
            P = sliced_product(L[c0:, :c0], L[c0:c1, :c0], s) if s else L[c0:, :c0] @ L[c0:c1, :c0].T
            L[c0:, c0:c1] -= P
            Pz = sliced_product(Z[:, :c0], L[c0:c1, :c0], s) if s else Z[:, :c0] @ L[c0:c1, :c0].T
            Z[:, c0:c1] -= Pz
        # inside the super-panel: plain float64 (the K = 128 kernels stay on DMMA)
        D = L[c0:c1, c0:c1]
        D = np.tril(D) + np.tril(D, -1).T
        Ld = cholesky(D, lower=True)
        L[c0:c1, c0:c1] = Ld
        if c1 < n:
            L[c1:, c0:c1] = solve_triangular(Ld, L[c1:, c0:c1].T, lower=True).T
        Z[:, c0:c1] = solve_triangular(Ld, Z[:, c0:c1].T, lower=True).T
    L = np.tril(L)
    # backward solve Ti = Z L^-1 with the same split: long-K updates sliced
    T = Z.copy()
    for e1 in range(n, 0, -nb * sp):
        e0 = max(e1 - nb * sp, 0)
        if e1 < n:
            P = sliced_product(T[:, e1:], L[e1:, e0:e1].T, s) if s else T[:, e1:] @ L[e1:, e0:e1]
            T[:, e0:e1] -= P
        T[:, e0:e1] = solve_triangular(L[e0:e1, e0:e1], T[:, e0:e1].T, lower=True, trans="T").T
    return T


def main():
    R.set_threads(os.cpu_count() or 1)
    cfg = StampConfig(n1=4, n2=25, dtheta_arcsec=0.04, fade_kernel=1, postage_pad=0, npixpsf=42, oversamp=6,
                      instamp_pad_arcsec=0.8, n_inframe=5, kappaC_arr=np.array([5e-4]))
    blk = SynthBlock(cfg, n_image=3, seed=12345, psf_sigmas=(0.85, 0.95, 1.05), star=True)
    o = OracleOutStamp(blk, PSFTables(blk, R.iD5512C, R.gridD5512C, dedup=True), 2, 2)
    o.build_system_matrices()
    A, B = np.asarray(o.sysmata), np.asarray(o.mhalfb[0])
    kap = float(cfg.kappaC_arr[0]) * float(o.outovlc[0])
    n, m = A.shape[0], B.shape[0]
    W = A + kap * np.eye(n)
    lam = np.linalg.eigvalsh(W)
    ref = cho_solve((cholesky(W, lower=True), True), B.T).T
    print(f"stamp n = {n}, m = {m}, kappa = {kap:.3e}, cond(A + kappa I) = {lam[-1] / lam[0]:.2e}, "
          f"{BITS}-bit slices, pairs kept: k + l < s")
    base = solve_sliced(W, B, 0)
    print(f"  float64 schedule (no slicing): max|dT|/max|T| = {np.abs(base - ref).max() / np.abs(ref).max():.2e}")
    for s in range(3, 10):
        t0 = time.perf_counter()
        T = solve_sliced(W, B, s)
        err = np.abs(T - ref).max() / np.abs(ref).max()
        print(f"  s = {s}: {s * (s + 1) // 2:2d} INT8 GEMMs per product, max|dT|/max|T| = {err:.2e}"
              f"   ({time.perf_counter() - t0:.1f} s)", flush=True)


if __name__ == "__main__":
    main()
