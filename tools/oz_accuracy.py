#!/usr/bin/env python
"""Accuracy of the Cholesky solve against the reference-made goldens (tests/golden/full_*.npz) for the library in use
(B200_LIB / B200_OZAKI select the sliced INT8 path with NS = 8 / 7 digit planes or the all-DMMA path)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from pyimcom_b200 import _lib  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock, GpuOutStamp  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402

tag = f"lib={os.path.basename(_lib.LIB_PATH)} NS={_lib.lib.b200_ozaki_slices()} OZAKI={os.environ.get('B200_OZAKI', '1')}"
for name in ("cfg1", "cfg5", "p4"):
    spec = cases.FULL_CASES[name]
    blk = cases.make_full_block(name)
    g = np.load(os.path.join(ROOT, "tests", "golden", f"full_{name}.npz"))
    gb = GpuBlock(blk, PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)).prepare(stamps=[spec["stamp"]])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = GpuOutStamp(gb, *spec["stamp"])
    e = cases.full_errors(s, g)
    # residual of the defining equation in float64 (independent of any reference): T (A + kappa I) = -B/2
    A, mB, T = s.sysmata, s.mhalfb, s.Ti64
    res = 0.0
    for j in range(blk.cfg.n_out):
        kap = float(blk.cfg.kappaC_arr[0]) * float(s.outovlc[j])
        res = max(res, np.abs(T[j] @ A + kap * T[j] - mB[j]).max() / np.abs(mB[j]).max())
    print(f"{tag} {name}: Ti64 vs reference {e['Ti64']:.2e}, T {e['T']:.2e}, outimage {e['outimage']:.2e}, residual {res:.2e}", flush=True)
