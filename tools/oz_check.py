#!/usr/bin/env python
"""Sliced INT8 (Ozaki) GEMM on tcgen05 vs the DMMA GEMM and cuBLAS: accuracy and rate.  Run under `timeout`."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyimcom_b200 import _lib  # noqa: E402


def ptr(t):
    return C.c_void_p(t.data_ptr())


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(M, N, K, reps=3, graded=True, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((M, K), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((N, K), dtype=torch.float64, device="cuda", generator=g)
    if graded:  # rows of very different magnitude, entries spread over 6 decades inside a row
        A *= 10.0 ** (-6 * torch.rand((M, 1), dtype=torch.float64, device="cuda", generator=g))
        B *= 10.0 ** (-6 * torch.rand((N, 1), dtype=torch.float64, device="cuda", generator=g))
        A *= 10.0 ** (-6 * torch.rand((M, K), dtype=torch.float64, device="cuda", generator=g))
    C0 = torch.randn((M, N), dtype=torch.float64, device="cuda", generator=g) * 1e-3
    wb = int(_lib.lib.b200_ozaki_gemm_work_bytes(M, N, K))
    work = torch.empty(wb, dtype=torch.uint8, device="cuda")
    Coz = C0.clone()
    _lib.dev_ozaki_gemm_nt(ptr(A), K, ptr(B), K, ptr(Coz), N, M, N, K, ptr(work), wb, st())
    torch.cuda.synchronize()
    ref = C0 - A @ B.T
    bound = A.abs() @ B.abs().T + C0.abs()
    err = ((Coz - ref).abs() / bound).max().item()
    # normwise bound of the scheme: 2^-56-ish of rowmax(A) rowmax(B) K
    rm = A.abs().amax(dim=1, keepdim=True) * B.abs().amax(dim=1, keepdim=True).T * K
    err_n = ((Coz - ref).abs() / rm).max().item()
    Cd = C0.clone()
    if M % 128 == 0 and N % 128 == 0:
        _lib.dev_gemm_nt(ptr(A), K, ptr(B), K, ptr(Cd), N, M, N, K, -1, st())
        torch.cuda.synchronize()
        err_d = ((Cd - ref).abs() / bound).max().item()
    else:
        err_d = float("nan")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = {}
    for name, fn in (("ozaki_total", lambda: _lib.dev_ozaki_gemm_nt(ptr(A), K, ptr(B), K, ptr(Coz), N, M, N, K, ptr(work), wb, st())),
                     ("dmma", lambda: _lib.dev_gemm_nt(ptr(A), K, ptr(B), K, ptr(Cd), N, M, N, K, -1, st()))):
        if name == "dmma" and not (M % 128 == 0 and N % 128 == 0):
            continue
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t[name] = e0.elapsed_time(e1) / reps
    _lib.profile(1)
    _lib.dev_ozaki_gemm_nt(ptr(A), K, ptr(B), K, ptr(Coz), N, M, N, K, ptr(work), wb, st())
    torch.cuda.synchronize()
    pr = _lib.profile_read()
    _lib.profile(0)
    fl = 2.0 * M * N * K
    kern_ms = pr["gemm_nt"][0] / pr["gemm_nt"][2]
    print(f"M={M} N={N} K={K} graded={graded}: err/(|A||B|+|C|) ozaki {err:.2e} dmma {err_d:.2e}; normwise {err_n:.2e}; "
          f"ozaki kernel {kern_ms:.3f} ms = {fl / kern_ms / 1e9:.1f} TFLOP/s-equivalent (with slicing {t['ozaki_total']:.3f} ms)"
          + (f"; dmma {t['dmma']:.3f} ms = {fl / t['dmma'] / 1e9:.1f} TFLOP/s" if "dmma" in t else ""), flush=True)
    return err


if __name__ == "__main__":
    torch.cuda.set_device(0)
    ok = True
    for (M, N, K, gr) in ((128, 64, 64, False), (128, 64, 128, False), (256, 128, 512, True), (1024, 512, 512, True),
                          (4096, 512, 4096, True), (6272, 512, 6272, False), (8192, 8192, 2048, False)):
        ok &= run(M, N, K, graded=gr) < 1e-12
    print("OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
