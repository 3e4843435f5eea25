#!/usr/bin/env python
"""One GPU: coadd a block as `world` strips of 2x2 stamp-group rows (the --strong layout of bench.py), sum the cubes,
and compare with the unsharded block: which map rows differ, by how much."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402
from pyimcom_b200.shard import assign_stamp_groups  # noqa: E402

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 6
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
blk = bench.make_block(0, n1=n1)
cfg = blk.cfg
tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
full = GpuBlock(blk, tab).prepare().run()
ref = full.out_map.clone()
again = GpuBlock(blk, tab).prepare().run().out_map
print("unsharded twice identical:", bool(torch.equal(ref, again)))
cube = torch.zeros_like(ref)
for r in range(world):
    mine = assign_stamp_groups(cfg.n1P, world, r)
    gb = GpuBlock(blk, tab).prepare(stamps=mine).run()
    cube += gb.out_map
d = (cube - ref).abs().amax(dim=(0, 1, 3)).cpu().numpy() / float(ref.abs().max())
rows = np.nonzero(d)[0]
print(f"n1P={cfg.n1P} world={world}: rows that differ: {rows.tolist()[:40]}{'...' if rows.size > 40 else ''} ({rows.size} rows), max rel {d.max():.2e}")
