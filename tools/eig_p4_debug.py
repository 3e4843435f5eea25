#!/usr/bin/env python
"""Where do non-finite values appear when the eigensolver runs on the paper-4 stamp's system matrix (n = 6248)?"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pyimcom_b200 import _lib  # noqa: E402
from pyimcom_b200 import lakernel as GL  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402

ptr = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731

name = sys.argv[1] if len(sys.argv) > 1 else "p4"
spec = cases.FULL_CASES[name]
blk = cases.make_full_block(name)
tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
gb = GpuBlock(blk, tab).prepare(stamps=[spec["stamp"]])
k = gb.order.index(spec["stamp"]) if hasattr(gb, "order") else 0
ds, indata = gb.build_system(k, need_A=True)
A = ds.matrix()
n, npad = ds.n, ds.npad
print("n", n, "npad", npad, "A finite", bool(torch.isfinite(A).all()), "|A|", float(A.abs().max()))
W = A.clone()
d, e, tau = (torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(3))
_lib.dev_tridiag(ptr(W), W.stride(0), n, ptr(d), ptr(e), ptr(tau), st())
torch.cuda.synchronize()
print("tridiag finite", bool(torch.isfinite(d).all()), bool(torch.isfinite(e).all()), bool(torch.isfinite(tau).all()),
      bool(torch.isfinite(W).all()))
dn, en = d.cpu().numpy(), e.cpu().numpy()
print("  |d| range", np.abs(dn).min(), np.abs(dn).max(), " |e| range", np.abs(en[:n - 1]).min(), np.abs(en).max(),
      " zeros in e:", int((en[:n - 1] == 0).sum()))
import scipy.linalg as sl

lt = sl.eigvalsh_tridiagonal(dn, en[: n - 1])
print("  eig(T): min %.3e max %.3e  #neg %d  #|lam|<1e-15|T| %d" % (lt.min(), lt.max(), (lt < 0).sum(),
                                                                    (np.abs(lt) < 1e-15 * lt.max()).sum()))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", f"tridiag_{name}.npz"), d=dn, e=en, lt=lt)
os.environ["B200_EIGH_TIMING"] = "1"
lam, Vt, _ = GL.eigh_device(A.clone(), n)
torch.cuda.synchronize()
print("lam finite", bool(torch.isfinite(lam).all()), "Vt finite", bool(torch.isfinite(Vt).all()))
bad = (~torch.isfinite(Vt)).any(dim=1).nonzero().ravel()
print("rows with non-finite entries:", bad.numel(), bad[:20].tolist())
lamn = lam[:n].cpu().numpy()
print("|lam - eig(T)| max", np.abs(np.sort(lamn) - lt).max())
if bad.numel() == 0:
    V = Vt[:n, :n]
    print("orth", float((V @ V.T - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max()),
          "resid", float((V @ A[:n, :n] - lam[:n, None] * V).abs().max()))
