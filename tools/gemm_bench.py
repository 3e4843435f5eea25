"""Micro-benchmarks run on the GPU box: FP64 GEMM (ours vs cuBLAS through torch), batched Cholesky solve."""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyimcom_b200 import _lib
from pyimcom_b200 import lakernel as GL


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


out = {}
for N in (2048, 4096, 8192):
    A = torch.randn(N, N, dtype=torch.float64, device="cuda")
    B = torch.randn(N, N, dtype=torch.float64, device="cuda")
    Cm = torch.empty(N, N, dtype=torch.float64, device="cuda")
    t_cublas = timeit(lambda: torch.matmul(A, B.T, out=Cm))
    st = GL.stream_handle()
    t_ours = timeit(lambda: _lib.dev_gemm_nt(GL.ptr(A), N, GL.ptr(B), N, GL.ptr(Cm), N, N, N, N, 0, st))
    out[f"dgemm_{N}"] = dict(cublas_tflops=2 * N**3 / t_cublas / 1e12, ours_tflops=2 * N**3 / t_ours / 1e12)
    print(N, out[f"dgemm_{N}"], flush=True)

for n, m, nsys in ((1536, 768, 1), (1536, 768, 8), (5632, 1536, 1), (5632, 1536, 4)):
    rng = np.random.default_rng(0)
    G = torch.randn(n, n, dtype=torch.float64, device="cuda")
    A0 = G @ G.T / n + torch.eye(n, dtype=torch.float64, device="cuda")
    B0 = torch.randn(m, n, dtype=torch.float64, device="cuda")

    def run():
        Ws = [A0.clone() for _ in range(nsys)]
        Xs = [B0.clone() for _ in range(nsys)]
        GL.chol_solve_batch(Ws, Xs)

    def clone_only():
        Ws = [A0.clone() for _ in range(nsys)]
        Xs = [B0.clone() for _ in range(nsys)]

    t = timeit(run, reps=3, warm=1) - timeit(clone_only, reps=3, warm=1)
    fl = nsys * (n**3 / 3 + 2 * n * n * m)
    out[f"chol_{n}_{m}_x{nsys}"] = dict(seconds=t, tflops=fl / t / 1e12)
    print(n, m, nsys, out[f"chol_{n}_{m}_x{nsys}"], flush=True)
    # torch/cuSOLVER reference timing for context
    def ref():
        L = torch.linalg.cholesky(A0)
        torch.cholesky_solve(B0.T.contiguous(), L)
    t = timeit(ref, reps=2, warm=1)
    out[f"chol_{n}_{m}_cusolver"] = dict(seconds=t, tflops=(n**3 / 3 + 2 * n * n * m) / t / 1e12)
    print("cusolver", out[f"chol_{n}_{m}_cusolver"], flush=True)
json.dump(out, open("gpurun_out/gemm_bench.json", "w"), indent=1)
