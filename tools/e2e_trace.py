#!/usr/bin/env python
"""Fresh-GpuBlock end-to-end steps, many repeats, host timestamps inside run(): where do slow steps lose their time?"""
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
    nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    blk = bench.make_block(0, n1=2)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
    cfg = blk.cfg
    for rep in range(nrep):
        torch.cuda.synchronize()
        h = [time.perf_counter()]
        gb = GpuBlock(blk, tab)
        gb.prepare()
        h.append(time.perf_counter())
        ks = list(range(len(gb.order)))
        nb = gb.batch_size()
        h.append(time.perf_counter())
        gb.ensure_pairs([gb.plans[gb.order[k]] for k in ks])
        h.append(time.perf_counter())
        live = [(k, gb.plans[gb.order[k]]) + gb.build_system(k) for k in ks]
        h.append(time.perf_counter())
        torch.cuda.synchronize()
        h.append(time.perf_counter())
        kos = GL.solve_chol_batch([t[2] for t in live], cfg, 0)
        h.append(time.perf_counter())
        for u, (k, p, ds, indata) in enumerate(live):
            spec = gb.apply_spec(k, indata, want_T32=False, want_Ti64=False)
            res = GL.apply_T(ds, kos[u], 0, spec)
            gb._overlap_add(p, 0, res)
        h.append(time.perf_counter())
        maps = gb.download()
        h.append(time.perf_counter())
        names = ["prepare", "batch_size", "ensure_pairs", "build_enq", "build_sync", "solve", "apply_enq", "download"]
        d = [1e3 * (h[i + 1] - h[i]) for i in range(len(h) - 1)]
        tot = sum(d)
        from pyimcom_b200 import coadd as _c
        ms = torch.cuda.memory_stats()
        print(f"rep {rep:2d} total {tot:6.1f} | " + " ".join(f"{n} {v:.1f}" for n, v in zip(names, d))
              + f" | segs {ms.get('segment.all.allocated', 0)} reserved {ms.get('reserved_bytes.all.current', 0) / 2**30:.2f} "
              f"pool@{gb._pool.data_ptr() if gb._pool is not None else 0:x} parked {[t.numel() for t in _c._PARKED_POOLS.get(0, [])]}",
              flush=True)
        del live, kos, res, maps, gb


if __name__ == "__main__":
    main()
