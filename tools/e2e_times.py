#!/usr/bin/env python
"""Host-side breakdown of the end-to-end step (GpuBlock(blk, tab).prepare().run().download()) over several repeats."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402


def main():
    blk = bench.make_block(0, n1=int(sys.argv[1]) if len(sys.argv) > 1 else 2)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gb = GpuBlock(blk, tab)
        gb.prepare()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        gb.run()
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        maps = gb.download()
        t5 = time.perf_counter()
        print(f"rep {rep}: prepare host {1e3*(t1-t0):.1f} (plan {1e3*gb.host_seconds['plan']:.1f}, upload enqueue "
              f"{1e3*gb.host_seconds['upload_enqueue']:.1f}) +sync {1e3*(t2-t1):.1f} | run enqueue {1e3*(t3-t2):.1f} +sync "
              f"{1e3*(t4-t3):.1f} | download {1e3*(t5-t4):.1f} | total {1e3*(t5-t0):.1f} ms")


if __name__ == "__main__":
    main()
