#!/usr/bin/env python
"""Where does a step go?  Host enqueue time vs device time of the three phases of one batch of the bench workload."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200 import lakernel as GL  # noqa: E402
from pyimcom_b200.coadd import GpuBlock  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402


def main():
    blk = bench.make_block(0, n1=2)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
    t0 = time.perf_counter()
    gb = GpuBlock(blk, tab)
    gb.prepare()
    torch.cuda.synchronize()
    print(f"prepare (host planning + H2D): {1e3 * (time.perf_counter() - t0):.1f} ms")
    for rep in range(3):
        gb.reset_maps()
        gb.run()
    torch.cuda.synchronize()
    cfg = gb.cfg
    ks = list(range(len(gb.order)))
    for rep in range(2):
        gb.reset_maps()
        torch.cuda.synchronize()
        T = {}
        t0 = time.perf_counter()
        live = [(k, gb.plans[gb.order[k]]) + gb.build_system(k) for k in ks]
        T["build_enqueue"] = time.perf_counter() - t0
        torch.cuda.synchronize()
        T["build_total"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        kos = GL.solve_chol_batch([t[2] for t in live], cfg, 0)
        T["solve_total(incl. info sync)"] = time.perf_counter() - t0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for u, (k, p, ds, indata) in enumerate(live):
            spec = gb.apply_spec(k, indata, want_T32=False, want_Ti64=False)
            res = GL.apply_T(ds, kos[u], 0, spec)
            gb._overlap_add(p, 0, res)
        T["finalize_enqueue"] = time.perf_counter() - t0
        torch.cuda.synchronize()
        T["finalize_total"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        del live, kos, res
        torch.cuda.synchronize()
        T["free"] = time.perf_counter() - t0
        print({k: round(1e3 * v, 2) for k, v in T.items()})
    t0 = time.perf_counter()
    for rep in range(3):
        gb.reset_maps()
        gb.run()
    torch.cuda.synchronize()
    print(f"run(): {1e3 * (time.perf_counter() - t0) / 3:.1f} ms per step")
    print(torch.cuda.memory_summary(abbreviated=True)[:1500])


if __name__ == "__main__":
    main()
