#!/usr/bin/env python
"""Can the stage-(a) gather kernel (k_pair_blocks: LSU / L1 bound) run BESIDE the INT8 tensor kernel (k_oz_gemm: owns
the SM's TMEM and 198 KB of its shared memory) on the same SMs, instead of time-sharing them?

Measures, on one GPU: the pair-block launch of a 16-stamp batch alone, a long sliced INT8 GEMM alone, and both at once
on two streams (gated so that they start together); then a whole 64-stamp step.  Knobs: B200_LIB (a build whose
k_oz_gemm is capped at 128 registers: `python -m pyimcom_b200.build --tag=ozcap -DB200_OZ_LB=512`), B200_PAIR_CARVEOUT
(shared-memory carve-out k_pair_blocks asks for; set by this script).  Run under `timeout`."""
import ctypes as C
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


def ptr(t):
    return C.c_void_p(t.data_ptr())


def main():
    torch.cuda.set_device(0)
    n1 = int(os.environ.get("CORES_N1", "6"))
    blk = bench.make_block(0, n1=n1)
    gb = GpuBlock(blk, PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)).prepare()
    gb.reset_maps(); gb.reset_cache(); gb.run()  # warm-up: pools, planning, attributes
    torch.cuda.synchronize()
    plans = [gb.plans[gb.order[k]] for k in range(16)]
    plans = [p for p in plans if p.n > 0]
    M = N = K = 6272
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((M, K), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((N, K), dtype=torch.float64, device="cuda", generator=g)
    Cm = torch.zeros((M, N), dtype=torch.float64, device="cuda")
    wb = int(_lib.lib.b200_ozaki_gemm_work_bytes(M, N, K))
    work = torch.empty(wb, dtype=torch.uint8, device="cuda")
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()  # (sa is replaced per configuration)
    NP, NO = 3, 6  # pair launches / GEMMs per measurement

    marks = []  # (label, event) recorded on the two streams: the spans of the individual launches

    def mark(label):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        marks.append((label, e))

    def pairs():
        for q in range(NP):
            gb.reset_cache()
            mark(f"p{q}<")
            gb.ensure_pairs(plans)
            mark(f"p{q}>")

    def oz():
        for q in range(NO):
            mark(f"g{q}<")
            _lib.dev_ozaki_gemm_nt(ptr(A), K, ptr(B), K, ptr(Cm), N, M, N, K, ptr(work), wb,
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
            mark(f"g{q}>")

    def timed(do_pairs, do_oz):
        nonlocal sa
        torch.cuda.synchronize()
        gate = torch.cuda.Event()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        cur = torch.cuda.current_stream()
        torch.cuda._sleep(20_000_000)  # ~10 ms: both streams are fully enqueued before either starts
        ev[0].record(cur)
        gate.record(cur)
        with torch.cuda.stream(sa):
            sa.wait_event(gate)
            if do_pairs:
                pairs()
            ev[1].record(sa)
        with torch.cuda.stream(sb):
            sb.wait_event(gate)
            if do_oz:
                oz()
            ev[2].record(sb)
        cur.wait_event(ev[1])
        cur.wait_event(ev[2])
        ev[3].record(cur)
        torch.cuda.synchronize()
        timed.spans = " ".join(f"{lab}{ev[0].elapsed_time(e):.1f}" for lab, e in marks)
        marks.clear()
        return ev[0].elapsed_time(ev[1]), ev[0].elapsed_time(ev[2]), ev[0].elapsed_time(ev[3])

    lib = os.path.basename(_lib.LIB_PATH)
    # reference pool content (one CTA per tile) for the bit-identity check of the resident form
    os.environ.pop("B200_PAIR_RESIDENT", None)
    gb.reset_cache(); gb.ensure_pairs(plans)
    torch.cuda.synchronize()
    ref_pool = gb._pool[: gb._pool_used].clone()
    configs = [tuple(int(v) for v in c.split(":")) for c in
               os.environ.get("CORES_CONFIGS", "0:0,1:0,1:-3,2:-3").split(",")]  # resident CTAs per SM : stream priority
    for resident, prio in configs:
        if resident:
            os.environ["B200_PAIR_RESIDENT"] = str(resident)
        else:
            os.environ.pop("B200_PAIR_RESIDENT", None)
        sa = torch.cuda.Stream(priority=prio)
        gb.reset_cache(); gb.ensure_pairs(plans)
        torch.cuda.synchronize()
        same = bool(torch.equal(ref_pool, gb._pool[: gb._pool_used]))
        timed(True, True)
        tp = min(timed(True, False)[0] for _ in range(2))
        to = min(timed(False, True)[1] for _ in range(2))
        both = [timed(True, True) for _ in range(3)]
        spans = timed.spans
        b = min(both, key=lambda t: t[2])
        tag = f"{lib} resident={resident} prio={prio}"
        print(f"{tag}: blocks identical {same}; pairs alone {tp:.2f} ms ({NP} launches), INT8 GEMM alone {to:.2f} ms "
              f"({NO} x {M}^3 with slicing), together {b[2]:.2f} ms (pairs done at {b[0]:.2f}, GEMM at {b[1]:.2f}); "
              f"sum {tp + to:.2f} -> overlap {(tp + to - b[2]) / min(tp, to):.2f} of the shorter one", flush=True)
        print(f"{tag}: launch spans of the last joint run (ms after the gate; p = pair launch, g = GEMM): {spans}",
              flush=True)
        if os.environ.get("CORES_NO_STEP"):
            continue
        # the whole step, stage (a) on a stream of that priority
        with torch.cuda.stream(sa):
            gb.reset_maps(); gb.reset_cache(); gb.run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                gb.reset_maps(); gb.reset_cache(); gb.run()
            e1.record()
            torch.cuda.synchronize()
        print(f"{tag}: step {e0.elapsed_time(e1) / 3:.1f} ms ({len(gb.order)} stamps)", flush=True)


if __name__ == "__main__":
    main()
