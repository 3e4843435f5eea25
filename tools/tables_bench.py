#!/usr/bin/env python
"""Row f1: PSF-overlap table build for one 2x2 InStamp group of the paper-4 shape, host NumPy FFT vs device DFT-GEMM."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.psfovl_device import DeviceTables  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402


def build(cls, blk):
    tab = cls(blk, G.iD5512C, G.gridD5512C)
    Ga, Gb = (2, 2), (2, 4)
    t = [tab.get_self(Ga), tab.get_io(Ga), tab.get_cross(Ga, Gb)]
    torch.cuda.synchronize()
    return tab, t


def main():
    blk = bench.make_block(0, n1=2)
    build(DeviceTables, blk)  # warm-up (DFT matrices, kernels)
    t0 = time.perf_counter()
    dtab, dt = build(DeviceTables, blk)
    t_dev = time.perf_counter() - t0
    # the transform chain alone (spectra of the group already on the device): 63 products + partial inverse DFTs
    Ga, Gb = (2, 2), (2, 4)
    for d in (dtab.self_, dtab.io, dtab.cross, dtab._cache):
        keep = {k: v for k, v in d.items() if isinstance(k, tuple) and k and k[0] == "grp"} if d is dtab._cache else {}
        d.clear()
        d.update(keep)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dtab.get_self(Ga), dtab.get_io(Ga), dtab.get_cross(Ga, Gb)
    torch.cuda.synchronize()
    t_chain = time.perf_counter() - t0
    t0 = time.perf_counter()
    htab, ht = build(PSFTables, blk)
    t_host = time.perf_counter() - t0
    ntab = sum(int(np.prod(x.shape[:-2])) for x in ht)
    err = max(float(np.abs(d.cpu().numpy() - h).max() / np.abs(h).max()) for d, h in zip(dt, ht))
    print(f"paper-4 group: {ntab} tables of {ht[0].shape[-1]}^2 (nfft {blk.cfg.nfft}): host NumPy {1e3*t_host:.0f} ms "
          f"({os.cpu_count()} cores), device {1e3*t_dev:.0f} ms (products + inverse transforms alone {1e3*t_chain:.1f} ms), "
          f"max rel diff {err:.2e}")


if __name__ == "__main__":
    main()
