#!/usr/bin/env python
"""Sustained throughput on a larger paper-4-shaped block (several batches per block): python tools/big_block.py [n1]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyimcom_b200 import _lib  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402
from pyimcom_b200.synth import StampConfig, SynthBlock  # noqa: E402


def main():
    n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    c0 = bench.workload_cfg()
    cfg = StampConfig(n1=n1, n2=c0.n2, dtheta_arcsec=c0.dtheta_arcsec, fade_kernel=c0.fade_kernel, postage_pad=1,
                      npixpsf=c0.npixpsf, oversamp=c0.oversamp, instamp_pad_arcsec=1.24, n_out=1, n_inframe=6,
                      linear_algebra="Cholesky", kappaC_arr=np.array([6e-4]), uctarget=1e-6, sigmamax=0.5)
    blk = SynthBlock(cfg, n_image=6, seed=1000, psf_sigmas=(0.85, 0.9, 0.95, 1.0, 1.05, 1.1), star=True)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
    t0 = time.perf_counter()
    gb = GpuBlock(blk, tab).prepare()
    torch.cuda.synchronize()
    t_prep = time.perf_counter() - t0
    ns = len(gb.order)
    for rep in range(3):
        gb.reset_maps(); gb.reset_cache()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gb.run()
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        tot = sum(gb.plans[ji].n ** 2 / 2 for ji in gb.order)
        print(f"n1P={cfg.n1P} stamps={ns} prepare {1e3*t_prep:.0f} ms run {1e3*t:.1f} ms -> {ns*cfg.n2**2/t:.0f} px/s, "
              f"{1e3*t/ns:.2f} ms/stamp; interpolated {gb.pair_points/tot:.2f} of the n^2/2 entries; "
              f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)


if __name__ == "__main__":
    main()
