#!/usr/bin/env python
"""Dump the CUDA IterKernel result of config 3 (tests/cases.py FULL_CASES['cfg3']) for offline comparison with the
oracle: float32 T and the per-pixel CG iteration counts -> gpurun_out/cfg3_gpu.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock, GpuOutStamp  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
spec = cases.FULL_CASES[name]
blk = cases.make_full_block(name)
gb = GpuBlock(blk, PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)).prepare(stamps=[spec["stamp"]])
s = GpuOutStamp(gb, *spec["stamp"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"{name}_gpu.npz"), T=s.T[0], niter=s.extras[0]["niter"],
                    Ti64=s.Ti64[0].astype(np.float32))
print("wrote", name, s.T.shape)
