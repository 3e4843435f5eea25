#!/usr/bin/env python
"""Are two ranks (two processes on two GPUs) bit-consistent?  Every rank builds the SAME block and coadds the SAME
OutStamp; hashes of each intermediate (host inputs, PSF-overlap tables, A, -B/2, float64 T, outimage) are compared on
rank 0.  Launch with torchrun --nproc-per-node 2."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock, GpuOutStamp  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
h = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]  # noqa: E731
blk = bench.make_block(0, n1=2)
out = {"x_val": h(blk.inimages[0].x_val), "data": h(blk.inimages[0].data)}
tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
gb = GpuBlock(blk, tab).prepare(stamps=[(2, 2)])
for k, v in list(tab.self_.items())[:1]:
    out["self_table"] = h(v)
for k, v in list(tab.io.items())[:1]:
    out["io_table"] = h(v)
out["outovlc"] = h(np.asarray(tab.outovlc))
s = GpuOutStamp(gb, 2, 2)
out.update(A=h(s.sysmata), mB=h(s.mhalfb), Ti64=h(s.Ti64), T=h(s.T), outimage=h(s.outimage), UC=h(s.UC))
os.environ["B200_OZAKI"] = "0"
from pyimcom_b200 import lakernel as GL  # noqa: E402

GL.OZAKI = False
s2 = GpuOutStamp(GpuBlock(blk, tab).prepare(stamps=[(2, 2)]), 2, 2)
out.update(Ti64_dmma=h(s2.Ti64))
# the same strip of a larger block on every rank, sliced-INT8 path and all-DMMA path
from pyimcom_b200.shard import assign_stamp_groups  # noqa: E402

blk6 = bench.make_block(0, n1=6)
tab6 = PSFTables(blk6, G.iD5512C, G.gridD5512C, dedup=True)
mine = assign_stamp_groups(blk6.cfg.n1P, 2, 1)
for oz in (True, False):
    GL.OZAKI = oz
    g1 = GpuBlock(blk6, tab6).prepare(stamps=mine).run()
    out[f"strip1_oz{int(oz)}"] = h(g1.out_map.cpu().numpy())
    g1.reset_maps(); g1.reset_cache(); g1.run()
    out[f"strip1_oz{int(oz)}_again"] = h(g1.out_map.cpu().numpy())
    out[f"strip1_oz{int(oz)}_batch"] = str(g1.batch_size())
allo = [None] * world
dist.all_gather_object(allo, out)
if rank == 0:
    for k in out:
        vals = [o[k] for o in allo]
        print(f"{k:12s} {'same' if len(set(vals)) == 1 else 'DIFFERENT'} {vals}")
dist.destroy_process_group()
