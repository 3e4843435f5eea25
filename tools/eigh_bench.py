#!/usr/bin/env python
"""eigh on the device for the config-2 stamp matrix (n = 1532) and a batch of 16 such stamps: sweeps, time, residuals,
for a few settings of the experiment knobs (B200_EIGH_INNER, ...)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pyimcom_b200 import lakernel as GL  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402

blk = cases.make_full_block("cfg2")
gb = GpuBlock(blk, PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)).prepare()
dss = [gb.build_system(k)[0] for k in range(len(gb.order))]
torch.cuda.synchronize()
A0 = dss[5].matrix()[: dss[5].n, : dss[5].n].cpu().numpy()
lam_ref = np.linalg.eigvalsh(A0)
for setting in sys.argv[1:] or ["B200_EIGH_INNER=1"]:
    for kv in setting.split(","):
        k, v = kv.split("=")
        os.environ[k] = v
    for nb in ((16,) if os.environ.get("EIGH_BENCH_ONLY16") else (1, 16)):
        mk = lambda: ([(ds.matrix().clone(), ds.n) for ds in dss[:nb]] if nb > 1  # noqa: E731
                      else [(dss[5].matrix().clone(), dss[5].n)])
        GL.eigh_device_batch(mk())  # warm-up: scratch allocations, kernel attributes
        items = mk()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res, sweeps = GL.eigh_device_batch(items)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if nb == 1:
            lam, Vt = res[0]
            n = dss[5].n
            lam, V = lam[:n].cpu().numpy(), Vt[:n, :n].cpu().numpy().T
            err_l = np.abs(np.sort(lam) - lam_ref).max()
            orth = np.abs(V.T @ V - np.eye(n)).max()
            resid = np.abs(A0 @ V - V * lam).max()
            print(f"{setting}: 1 stamp n={n}: {ms:.1f} ms, {sweeps} sweeps, |dlam| {err_l:.2e}, orth {orth:.2e}, resid {resid:.2e}", flush=True)
        else:
            print(f"{setting}: {nb} stamps: {ms:.1f} ms, {sweeps} sweeps", flush=True)
