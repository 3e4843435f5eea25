#!/usr/bin/env python
"""Rate of the sliced INT8 GEMM kernel alone (no accuracy check): B200_OZ_DBG=1 (TMA only) / 2 (MMA only) experiments."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyimcom_b200 import _lib  # noqa: E402

ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
torch.cuda.set_device(0)
for (M, N, K) in ((6272, 512, 6272), (8192, 2048, 2048), (9472, 512, 512)):
    A = torch.randn((M, K), dtype=torch.float64, device="cuda")
    B = torch.randn((N, K), dtype=torch.float64, device="cuda")
    Cm = torch.zeros((M, N), dtype=torch.float64, device="cuda")
    wb = int(_lib.lib.b200_ozaki_gemm_work_bytes(M, N, K))
    work = torch.empty(wb, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        _lib.dev_ozaki_gemm_nt(ptr(A), K, ptr(B), K, ptr(Cm), N, M, N, K, ptr(work), wb, st())
    torch.cuda.synchronize()
    _lib.profile(1)
    for _ in range(3):
        _lib.dev_ozaki_gemm_nt(ptr(A), K, ptr(B), K, ptr(Cm), N, M, N, K, ptr(work), wb, st())
    torch.cuda.synchronize()
    pr = _lib.profile_read()
    _lib.profile(0)
    ms = pr["gemm_nt"][0] / pr["gemm_nt"][2]
    tiles = (M // 128) * (N // 64)
    waves = -(-tiles // 148)
    print(f"dbg={os.environ.get('B200_OZ_DBG', '0')} M={M} N={N} K={K}: {ms:.3f} ms, {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s-eq, "
          f"{tiles} tiles ({waves} waves), {ms * 1e3 / waves / (K // 64):.2f} us per K chunk per tile", flush=True)
