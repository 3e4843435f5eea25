#!/usr/bin/env python
"""The smallest program that runs bench.py's step (for ncu: every intercepted launch costs milliseconds there, so the
bench's peak measurements, profile repeats and end-to-end passes are left out): one warm-up block, then N timed-shape
blocks.  `--n1` chooses the block size (default: the bench block, n1P = 8)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from pyimcom_b200 import _lib  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n1", type=int, default=6)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
args = ap.parse_args()
torch.cuda.set_device(0)
blk = bench.make_block(0, n1=args.n1)
gb = GpuBlock(blk, PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)).prepare()
for _ in range(args.warmup):
    gb.reset_maps(); gb.reset_cache(); gb.run()
torch.cuda.synchronize()
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    gb.reset_maps(); gb.reset_cache(); gb.run()
e1.record()
torch.cuda.synchronize()
print(f"{len(gb.order)} stamps per block, {(_lib.launch_count() - n0) // args.steps} launches per step, "
      f"{e0.elapsed_time(e1) / args.steps:.1f} ms per step, pool evictions {gb.pool_evictions}", flush=True)
