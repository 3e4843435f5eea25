"""What does the library DGEMM that sets the FP64 roofline look like?  (run under ncu: launch geometry, registers,
shared memory and pipe utilisation of the cuBLAS kernel behind torch.matmul for the 6144^3 peak measurement)"""
import torch

n = 6144
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    c = a @ b.T
torch.cuda.synchronize()
print(float(c[0, 0]))
