"""Resident-block step time of the bench workload for the current environment knobs (B200_SP, B200_SOLVE_STREAMS,
B200_MAX_BATCH, B200_TILE64): one line per process, so a shell loop can sweep the knobs (they are read once).

    for sp in 2 3 4 6 8; do B200_SP=$sp python tools/knob_sweep.py; done
"""
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402

blk = bench.make_block(0)
tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
gb = GpuBlock(blk, tab)
gb.prepare()


def step():
    gb.reset_maps()
    gb.reset_cache()
    gb.run()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 3
e0.record()
for _ in range(K):
    step()
e1.record()
torch.cuda.synchronize()
knobs = {k: os.environ[k] for k in ("B200_SP", "B200_SOLVE_STREAMS", "B200_MAX_BATCH", "B200_TILE64") if k in os.environ}
print(knobs, "ms_per_step %.1f" % (e0.elapsed_time(e1) / K), flush=True)
