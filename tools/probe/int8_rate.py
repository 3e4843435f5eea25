"""INT8 tensor-core rate available on this GPU through the library (torch._int_mm -> cuBLASLt), for the Ozaki-slicing
estimate of DESIGN.md section 9: an FP64 product from s slices per operand costs s(s+1)/2 INT8 GEMMs, so the
break-even against the 36 TFLOP/s DMMA rate at s = 8 is 1.3 Pop/s.  Shapes: a square GEMM and the shape of one
super-panel update (all row tiles of 6 systems x 512 columns, K = 3072)."""
import torch


def timeit(fn, reps=10, warm=3):
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


for M, N, K in ((8192, 8192, 8192), (16384, 16384, 4096), (28000, 512, 3072), (4736, 512, 3072)):
    a = torch.randint(-64, 64, (M, K), dtype=torch.int8, device="cuda")
    b = torch.randint(-64, 64, (K, N), dtype=torch.int8, device="cuda")
    t = timeit(lambda: torch._int_mm(a, b))
    af, bf = a.to(torch.bfloat16), b.to(torch.bfloat16)
    t16 = timeit(lambda: af @ bf)
    print(f"M={M} N={N} K={K}: int8 {2 * M * N * K / t / 1e12:.0f} Top/s ({1e3 * t:.3f} ms), "
          f"bf16 {2 * M * N * K / t16 / 1e12:.0f} TFLOP/s", flush=True)
