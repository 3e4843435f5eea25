#!/usr/bin/env python
"""Per-step, per-phase device and host times of the resident bench step over many repeats (noise hunting)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyimcom_b200 import lakernel as GL  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402


def main():
    nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    blk = bench.make_block(0, n1=2)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
    gb = GpuBlock(blk, tab).prepare()
    cfg = gb.cfg
    ks = list(range(len(gb.order)))
    for _ in range(3):
        gb.reset_maps(); gb.reset_cache(); gb.run()
    torch.cuda.synchronize()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    for rep in range(nrep):
        gb.reset_maps(); gb.reset_cache()
        torch.cuda.synchronize()
        e = [ev() for _ in range(4)]
        h = [time.perf_counter()]
        e[0].record()
        gb.ensure_pairs([gb.plans[gb.order[k]] for k in ks])
        live = [(k, gb.plans[gb.order[k]]) + gb.build_system(k) for k in ks]
        e[1].record(); h.append(time.perf_counter())
        kos = GL.solve_chol_batch([t[2] for t in live], cfg, 0)
        e[2].record(); h.append(time.perf_counter())
        for u, (k, p, ds, indata) in enumerate(live):
            spec = gb.apply_spec(k, indata, want_T32=False, want_Ti64=False)
            res = GL.apply_T(ds, kos[u], 0, spec)
            gb._overlap_add(p, 0, res)
        e[3].record(); h.append(time.perf_counter())
        torch.cuda.synchronize()
        h.append(time.perf_counter())
        d = [e[i].elapsed_time(e[i + 1]) for i in range(3)]
        print(f"rep {rep:2d} device build {d[0]:6.1f} solve {d[1]:6.1f} apply {d[2]:5.1f} total {sum(d):6.1f} | host build "
              f"{1e3*(h[1]-h[0]):5.1f} solve {1e3*(h[2]-h[1]):6.1f} apply {1e3*(h[3]-h[2]):5.1f} tail {1e3*(h[4]-h[3]):5.1f}", flush=True)
        del live, kos, res


if __name__ == "__main__":
    main()
