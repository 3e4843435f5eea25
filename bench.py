#!/usr/bin/env python
"""bench.py -- coadd output pixels/sec of the per-postage-stamp coaddition hot path on B200.

    python bench.py --gpus N --steps K --warmup W           (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU path on the host cores, same config)

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): a paper4-shaped synthetic
block -- n2 = 32, FADE = 3 (m = 1444 output px per stamp incl. fade), dtheta = 0.0390625", NPIXPSF = 48,
oversamp = 8 (395^2-entry padded PSF-overlap tables), INPAD = 1.24", KAPPAC = [6e-4], 6 input images,
n_inframe = 6 layers, CholKernel; n ~ 6.3 k selected input pixels per stamp -- reduced to n1P x n1P = 8 x 8
output stamps per block so that a step takes about half a second.  One STEP = one block (64 OutStamps, four batches
of 16): gather -> A / mBhalf assembly -> batched FP64 Cholesky + triangular solves -> T apply -> overlap-add.
Every rank owns its own block (weak scaling, seed = 1000 + rank); the only collective is the final gather of
the output cube to rank 0 (NCCL).

value : stamps * n2^2 / time with the block inputs already resident in HBM (prepare() done before the clock)
e2e   : the same through the public API from HOST buffers: GpuBlock(blk, tables).prepare().run().download()
        (host planning, pinned H2D of pixels/tables/plans, the stamp loop, D2H of the block maps) + gather.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv[1:]:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all host cores (BLAS included), and
    # the thread pools read the variable when numpy / scipy are first imported
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from pyimcom_b200.synth import StampConfig, SynthBlock  # noqa: E402

# DRAM traffic of the dominant kernel from the committed ncu capture (profiles/ncu_r01.txt: k_chol_super_update,
# grid (8,74,6) = super-panel c0=24 of 49 block columns for the 6 systems of one solve stream, 2.48 ms under ncu):
# dram__bytes_read.sum + dram__bytes_write.sum of that launch
NCU_TRAFFIC_BYTES = 844.72192e6 + 102.279936e6
NCU_TRAFFIC_NOTE = ("ncu --set full, one launch of k_chol_super_update (super-panel c0=24 of 49 block columns, 6 systems of one "
                    "solve stream): 0.845 GB read + 0.102 GB written vs 8.8e10 flop => 93 flop/B, far above the FP64 ridge "
                    "(~5.5 flop/B); tensor pipe 93.1 % busy")
METRIC = "coadd_output_pixels_per_sec"
UNIT = "output px/s"
SEED0 = 1000


def workload_cfg(kernel="Cholesky", n1=6):
    return StampConfig(n1=n1, n2=32, dtheta_arcsec=0.0390625, fade_kernel=3, postage_pad=1, npixpsf=48, oversamp=8,
                       instamp_pad_arcsec=1.24, n_out=1, n_inframe=6, linear_algebra=kernel,
                       kappaC_arr=np.array([6e-4]), uctarget=1e-6, sigmamax=0.5)


def make_block(rank, n1=6):
    cfg = workload_cfg(n1=n1)
    return SynthBlock(cfg, n_image=6, seed=SEED0 + rank, psf_sigmas=(0.85, 0.9, 0.95, 1.0, 1.05, 1.1), star=True)


def config_dict(cfg, n_stamps, extra=None):
    d = {"workload": f"paper4-shaped synthetic block (BASELINE.json configs[3]), CholKernel, reduced to n1P={cfg.n1P}",
         "n2": cfg.n2, "fade": cfg.fade_kernel, "n2f": cfg.n2f, "m": cfg.n2f**2, "stamps_per_block": n_stamps,
         "n_images": 6, "n_inframe": cfg.n_inframe, "kappaC": [6e-4], "npixpsf": cfg.npixpsf, "oversamp": cfg.oversamp,
         "inpad_arcsec": 1.24, "dtheta_arcsec": cfg.dtheta_arcsec, "blocks_per_gpu_per_step": 1,
         "l2_policy": "per-stamp working set (A 0.31 GB + mBhalf 0.08 GB, fresh buffers every stamp) exceeds the 126 MB L2"}
    if extra:
        d.update(extra)
    return d


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None
        self.i0 = self.i1 = None

    def mark(self):
        self.i0 = len(self.lines)

    def mark_end(self):
        time.sleep(0.12)  # let the 100 ms sampler print the last line of the timed region
        self.i1 = len(self.lines)

    def __enter__(self):
        if os.environ.get("B200_BENCH_NO_SAMPLER"):
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        lines = self.lines[self.i0:self.i1] if self.i0 is not None and self.lines[self.i0:self.i1] else self.lines
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's algorithm) on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_stamps(blk, n_stamps, threads):
    """Time the oracle's OutStamp path for the first n_stamps of the block; returns seconds."""
    import warnings

    from oracle import lakernel as OL
    from oracle import routines as R
    from oracle.sysmat import OracleOutStamp
    from pyimcom_b200.psfovl_host import PSFTables

    import contextlib

    R.set_threads(os.cpu_count() or 1)
    tab = PSFTables(blk, R.iD5512C, R.gridD5512C, dedup=True)
    order = list(blk.stamp_order())[:n_stamps]
    for (j, i) in order[:1]:  # builds the PSF-overlap tables outside the clock, as on the GPU arm
        OracleOutStamp(blk, tab, j, i).build_system_matrices()
    R.set_threads(threads)
    limit = contextlib.nullcontext()
    if threads == 1:  # production layout of the reference: one core per block (docs/run_README.rst)
        from threadpoolctl import threadpool_limits

        limit = threadpool_limits(limits=1)
    ii_cache = {}  # the reference's SysMatA cache: InStamp-pair blocks are interpolated once per block
    with limit:
        t0 = time.perf_counter()
        for (j, i) in order:
            o = OracleOutStamp(blk, tab, j, i, ii_cache=ii_cache)
            o.build_system_matrices()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                OL.CholKernel(o)()
            o.post_kernel()
            o.perform_coaddition()
        return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    blk = make_block(0)
    cfg = blk.cfg
    sample = 4  # OutStamps per step: a bounded sample of the block (~5 s of CPU work per step on 16 cores)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_stamps(blk, 1, cores)
    times = [cpu_stamps(blk, sample, cores) for _ in range(args.steps)]
    t = float(np.mean(times))
    val = sample * cfg.n2**2 / t
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, sample, {"note": f"each step = a bounded sample of {sample} OutStamps of the same block"}),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} OutStamps per step, oracle/ (C + OpenMP interpolation, SciPy/OpenBLAS Cholesky)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from pyimcom_b200 import _lib
    from pyimcom_b200 import pyimcom_croutines as G
    from pyimcom_b200.coadd import GpuBlock
    from pyimcom_b200.psfovl_host import PSFTables
    from pyimcom_b200.shard import gather_cube

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    blk = make_block(rank)
    cfg = blk.cfg
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)  # PSF-overlap tables (row f1): built once, outside the clock
    gb = GpuBlock(blk, tab)
    gb.prepare()
    n_stamps = len(gb.order)
    px_per_step = n_stamps * cfg.n2**2

    # FP64 peak of this GPU: cuBLAS DGEMM through torch (MEASURED_PEAKS.json carries no FP64 entry)
    N = 6144
    a = torch.randn(N, N, dtype=torch.float64, device="cuda")
    b = torch.randn(N, N, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    best = 1e9
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1) * 1e-3)
    fp64_peak = 2 * N**3 / best / 1e12
    del a, b, c

    def step_resident():
        gb.reset_maps()
        gb.reset_cache()  # every step is a new block: no InStamp-pair block survives from the previous one
        gb.run()
        return gather_cube(gb.out_map, world, rank)

    e2e_host = []

    def step_e2e():
        t0 = time.perf_counter()
        g2 = GpuBlock(blk, tab)
        g2.prepare()
        t1 = time.perf_counter()
        g2.run()
        t2 = time.perf_counter()
        maps = g2.download()
        gather_cube(g2.out_map, world, rank)
        t3 = time.perf_counter()
        ms = torch.cuda.memory_stats()
        e2e_host.append((round(1e3 * (t1 - t0), 1), round(1e3 * (t2 - t1), 1), round(1e3 * (t3 - t2), 1),
                         ms.get("segment.all.allocated", 0), ms.get("segment.all.freed", 0),
                         round(ms.get("reserved_bytes.all.current", 0) / 2**30, 2), ms.get("num_alloc_retries", 0)))
        return g2, maps

    # the clock sampler starts before the warm-up (nvidia-smi takes a moment to produce its first line on an 8-GPU
    # box); only the lines printed during the timed region are kept if there are any
    with ClockSampler(local) as clk:
        for _ in range(max(args.warmup, 3)):
            step_resident()
        # ---- timed: resident inputs ----
        barrier()
        n0 = _lib.launch_count()
        clk.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_resident()
        e1.record()
        barrier()
        clk.mark_end()
    launches = _lib.launch_count() - n0
    t_res = e0.elapsed_time(e1) * 1e-3
    # ---- the same K steps again with a pair of CUDA events around every launch of the library (per-kernel device
    # time for the roofline): identical work and stream layout; the event records themselves cost a few percent, which
    # is why the throughput above is taken without them ----
    barrier()
    _lib.profile(1)
    e0p, e1p = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0p.record()
    for _ in range(args.steps):
        step_resident()
    e1p.record()
    barrier()
    t_prof = e0p.elapsed_time(e1p) * 1e-3
    prof = _lib.profile_read()
    _lib.profile(0)
    # ---- one extra step with the solve groups serialised on one stream: per-launch event intervals of the timed
    # region above include time-sharing between the concurrent solve streams; this pass shows each kernel alone ----
    from pyimcom_b200 import lakernel as GL

    n_streams = GL.SOLVE_STREAMS
    prof_serial = {}
    if n_streams > 1:
        GL.SOLVE_STREAMS = 1
        step_resident()
        barrier()
        _lib.profile(1)
        step_resident()
        barrier()
        prof_serial = _lib.profile_read()
        _lib.profile(0)
        GL.SOLVE_STREAMS = n_streams
    # ---- timed: end to end from host buffers ----
    for _ in range(4):  # warm-up, written exactly like the timed loop (the previous block stays referenced while the next
        g2, maps = step_e2e()  # one is built): the caching allocators must have seen two coexisting blocks before the
        # clock starts -- a late cudaMalloc of a GB-sized segment costs 10-200 ms
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        g2, maps = step_e2e()
    e1.record()
    barrier()
    t_e2e = e0.elapsed_time(e1) * 1e-3
    h2d = int(g2.h2d_bytes)
    d2h = int(sum(v.nbytes for v in maps.values()))
    if world > 1:
        tt = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(tt[0]), float(tt[1])
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        TENSOR = ("chol_super_update", "chol_panel", "chol_inner_update", "back_super_update", "back_diag",
                  "back_inner_update", "gemm_nt", "potrf_diag")

        def stage_table(pr, t_total):
            out = {}
            for name, (ms, work, cnt) in pr.items():
                tensor = name in TENSOR
                rate = work / (ms * 1e-3) / (1e12 if tensor else 1e9) if ms > 0 else 0.0
                out[name] = {"launches": cnt, "ms_total": round(ms, 3), "share_of_step": round(ms * 1e-3 / t_total, 4),
                             "achieved": round(rate, 3), "unit": "TFLOP/s" if tensor else "GB/s",
                             "frac": round(rate / (fp64_peak if tensor else hbm_peak), 4)}
            return out

        def dmma_total(pr):
            ms = sum(pr[k][0] for k in pr if k.startswith(("chol_", "back_")))
            fl = sum(pr[k][1] for k in pr if k.startswith(("chol_", "back_")))
            return ms, fl

        stages = stage_table(prof, t_prof)
        dom = "chol_super_update"
        ms, work, cnt = prof.get(dom, (0.0, 0.0, 0))
        ach = work / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        tensor_ms, tensor_fl = dmma_total(prof)
        roofline = {"bound": "tensor", "kernel": "k_chol_super_update (FP64 DMMA m8n8k4 tile GEMM, long-K)",
                    "achieved": round(ach, 3), "peak": round(fp64_peak, 3), "unit": "TFLOP/s",
                    "frac": round(ach / fp64_peak, 4),
                    "traffic": NCU_TRAFFIC_BYTES,
                    "traffic_note": NCU_TRAFFIC_NOTE,
                    "peak_source": f"cuBLAS DGEMM {N}^3 through torch.matmul, best of 5, measured in this run "
                                   "(MEASURED_PEAKS.json has no FP64 entry)",
                    "launches": cnt, "avg_launch_ms": round(ms / max(cnt, 1), 4),
                    "flops_per_launch": work / max(cnt, 1),
                    "note": (f"per-launch events are recorded on a repeat of the K timed steps (same work, same {n_streams} "
                             "concurrent solve streams, batches pipelined): a launch's event interval includes the time it "
                             "shares the SMs with the other streams' kernels; 'serialized' repeats one step on one stream"),
                    "all_dmma_kernels": {"achieved": round(tensor_fl / (tensor_ms * 1e-3) / 1e12, 3) if tensor_ms else 0,
                                         "share_of_step": round(tensor_ms * 1e-3 / t_prof, 4),
                                         "flops_per_step": tensor_fl / max(args.steps, 1),
                                         "achieved_wall": round(tensor_fl / t_res / 1e12, 3)},
                    "ms_per_step_with_events": round(1e3 * t_prof / args.steps, 3),
                    "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src, "stages": stages}
        if prof_serial:
            t_ser = sum(v[0] for v in prof_serial.values()) * 1e-3
            ms_s, work_s, cnt_s = prof_serial.get(dom, (0.0, 0.0, 0))
            ach_s = work_s / (ms_s * 1e-3) / 1e12 if ms_s > 0 else 0.0
            sm, sf = dmma_total(prof_serial)
            roofline["serialized"] = {"achieved": round(ach_s, 3), "frac": round(ach_s / fp64_peak, 4), "launches": cnt_s,
                                      "avg_launch_ms": round(ms_s / max(cnt_s, 1), 4),
                                      "all_dmma_kernels": {"achieved": round(sf / (sm * 1e-3) / 1e12, 3) if sm else 0},
                                      "stages": stage_table(prof_serial, max(t_ser, 1e-9))}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu:
            ncpu = 8  # ~10 s of CPU work on 16 cores
            tcpu = cpu_stamps(blk, ncpu, cores)
            cpu = {"value": ncpu * cfg.n2**2 / tcpu, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {ncpu} OutStamps of the same block, oracle/ (C + OpenMP interpolation on all cores, "
                             "SciPy/OpenBLAS Cholesky)", "seconds_per_stamp": tcpu / ncpu}
            t1 = cpu_stamps(blk, 1, 1)  # a second figure with one thread (the reference runs one core per block)
            cpu["single_thread"] = {"value": cfg.n2**2 / t1, "unit": UNIT, "cores": 1, "sample": "first OutStamp of the block",
                                    "seconds_per_stamp": t1}
        n_in = int(np.mean([gb.plans[ji].n for ji in gb.order]))
        line = {"metric": METRIC, "value": world * px_per_step * args.steps / t_res, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_res / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(cfg, n_stamps, {"n_input_px_per_stamp": n_in, "parallelism": f"blocks x{world}"}),
                "e2e": {"value": world * px_per_step * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / args.steps,
                        "host_ms_prepare_run_download": e2e_host[-args.steps:]},
                "gpu_launches": int(launches), "stamps_per_sec": world * n_stamps * args.steps / t_res,
                "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
