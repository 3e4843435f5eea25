import sys, torch
sys.path.insert(0, ".")
from pyimcom_b200 import _lib
from pyimcom_b200 import lakernel as GL
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
for N, K in ((4096, 4096), (8192, 8192), (6144, 128), (6144, 3072)):
    A = torch.randn(N, K, dtype=torch.float64, device="cuda")
    B = torch.randn(N, K, dtype=torch.float64, device="cuda")
    Cm = torch.zeros(N, N, dtype=torch.float64, device="cuda")
    st = GL.stream_handle()
    t_c = timeit(lambda: torch.matmul(A, B.T, out=Cm))
    ref = Cm.clone()
    t_o = timeit(lambda: _lib.dev_gemm_nt(GL.ptr(A), K, GL.ptr(B), K, GL.ptr(Cm), N, N, N, K, 0, st))
    err = float((Cm - ref).abs().max())
    t_s = timeit(lambda: _lib.dev_gemm_nt(GL.ptr(A), K, GL.ptr(B), K, GL.ptr(Cm), N, N, N, K, -1, st))
    print(N, K, "cublas %.2f ours %.2f ours(sub) %.2f TF/s  maxerr %.2e" % (2*N*N*K/t_c/1e12, 2*N*N*K/t_o/1e12, 2*N*N*K/t_s/1e12, err), flush=True)
