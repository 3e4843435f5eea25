#!/usr/bin/env python
"""Per-kind device times of one 16-stamp paper-4-shaped block on one solve stream (the library's per-launch events),
plus a hash of the block's output map: the quick A/B check of a kernel change (B200_LIB selects the other build)."""
import os, sys, torch
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from pyimcom_b200 import _lib
from pyimcom_b200 import pyimcom_croutines as G
from pyimcom_b200.coadd import GpuBlock
from pyimcom_b200.psfovl_host import PSFTables
import pyimcom_b200.lakernel as LK
torch.cuda.set_device(0)
blk = bench.make_block(0, n1=2)
gb = GpuBlock(blk, PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)).prepare()
for _ in range(2):
    gb.reset_maps(); gb.reset_cache(); gb.run()
torch.cuda.synchronize()
import hashlib
h = hashlib.sha256(gb.out_map.cpu().numpy().tobytes()).hexdigest()[:16]
LK.SOLVE_STREAMS = 1
_lib.profile(1)
for _ in range(3):
    gb.reset_maps(); gb.reset_cache(); gb.run()
torch.cuda.synchronize()
pr = _lib.profile_read(); _lib.profile(0)
print("out_map sha", h, {k: (round(v[0] / 3, 3), int(v[2] / 3)) for k, v in pr.items() if v[2]})
