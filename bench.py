#!/usr/bin/env python
"""bench.py -- coadd output pixels/sec of the per-postage-stamp coaddition hot path on B200.

    python bench.py --gpus N --steps K --warmup W           (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the CPU path on the host cores, same config)

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): a paper4-shaped synthetic
block -- n2 = 32, FADE = 3 (m = 1444 output px per stamp incl. fade), dtheta = 0.0390625", NPIXPSF = 48,
oversamp = 8 (395^2-entry padded PSF-overlap tables), INPAD = 1.24", KAPPAC = [6e-4], 6 input images,
n_inframe = 6 layers, CholKernel; n ~ 6.3 k selected input pixels per stamp -- reduced to n1P x n1P = 8 x 8
output stamps per block so that a step takes about a quarter of a second.  One STEP = one block (64 OutStamps, two pipelined
batches of 32): gather -> A / mBhalf assembly -> batched FP64 Cholesky + triangular solves (long-K panel updates as
error-free sliced INT8 products on tcgen05, the rest on the FP64 DMMA pipe) -> T apply -> overlap-add.
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

# DRAM traffic of the dominant kernel: NOT measured by this run (ncu cannot run inside it).  It is read from the summary
# of the committed `ncu --set full` capture of the same kernel on the same workload (profiles/ncu_r02_top.json, written
# by tools/ncu_summary.py from the .ncu-rep; fields: kernel, grid, dram_bytes_read, dram_bytes_write, flops, source).
def ncu_traffic():
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_r02_top.json")))
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), d
    except (OSError, KeyError, ValueError):
        return None, {"note": "profiles/ncu_r02_top.json missing"}


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
    # OutStamps per step: 16 (a 4x4 corner of the block, so that the reference's SysMatA reuse of InStamp-pair blocks is
    # as on the GPU arm), fewer only if K steps would not finish within ~7 minutes on this host
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_stamps(blk, 1, cores)
    t1 = cpu_stamps(blk, 2, cores) / 2
    sample = int(max(4, min(16, 420.0 / max(args.steps, 1) / max(t1, 1e-3))))
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
    from pyimcom_b200.shard import assign_stamp_groups, gather_cube, reduce_cube

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

    if args.strong:
        return run_strong(args, world, rank, barrier)
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
    # INT8 tensor peak of this GPU (the pipe the long-K updates run on): cuBLASLt IGEMM through torch._int_mm
    NI = 8192
    ai = torch.randint(-64, 64, (NI, NI), dtype=torch.int8, device="cuda")
    bi = torch.randint(-64, 64, (NI, NI), dtype=torch.int8, device="cuda")
    best = 1e9
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ci = torch._int_mm(ai, bi)
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1) * 1e-3)
    int8_peak = 2 * NI**3 / best / 1e12
    del ai, bi, ci

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
    gather_verified = None
    if world > 1:
        tt = torch.tensor([t_res, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(tt[0]), float(tt[1])
        # the gathered cube is checked, outside the clock: every rank coadded a DIFFERENT block (seed 1000 + rank); its
        # checksums (sum, sum of squares, max |.| in float64) travel by all_gather and rank 0 recomputes them from the slice
        # of the gathered cube that should hold that rank's block
        def checks(t):
            t = t.double()
            return torch.stack([t.sum(), t.square().sum(), t.abs().max()])

        cube = gather_cube(gb.out_map, world, rank)
        mine = checks(gb.out_map)
        allc = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        if rank == 0:
            allc = torch.stack(allc)
            got = torch.stack([checks(cube[r]) for r in range(world)])
            distinct = len({tuple(row.tolist()) for row in allc}) == world
            gather_verified = bool(torch.equal(got, allc) and distinct and bool(torch.isfinite(allc).all()))
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        TENSOR = ("chol_super_update", "chol_panel", "chol_inner_update", "back_super_update", "back_diag",
                  "back_inner_update", "gemm_nt", "potrf_diag", "oz_gemm")
        ns = int(_lib.lib.b200_ozaki_slices())
        i8_per_f64 = ns * (ns + 1) // 2  # INT8 products evaluated per float64 product (digit pairs t + u < NS)

        def stage_table(pr, t_total):
            out = {}
            for name, (ms, work, cnt) in pr.items():
                tensor = name in TENSOR
                rate = work / (ms * 1e-3) / (1e12 if tensor else 1e9) if ms > 0 else 0.0
                out[name] = {"launches": cnt, "ms_total": round(ms, 3), "share_of_step": round(ms * 1e-3 / t_total, 4),
                             "achieved": round(rate, 3), "unit": "TFLOP/s" if tensor else "GB/s",
                             "frac": round(rate / (fp64_peak if tensor else hbm_peak), 4)}
                if name == "oz_gemm":  # work is counted in float64-equivalent flops; the pipe executes 36x that in INT8
                    out[name].update(unit="TFLOP/s float64-equivalent", frac_of_dgemm_peak=out[name].pop("frac"),
                                     int8_tops=round(rate * i8_per_f64, 1),
                                     frac_of_int8_peak=round(rate * i8_per_f64 / int8_peak, 4))
            return out

        def tensor_total(pr):
            ks = [k for k in pr if k.startswith(("chol_", "back_", "oz_gemm"))]
            return sum(pr[k][0] for k in ks), sum(pr[k][1] for k in ks)

        stages = stage_table(prof, t_prof)
        dom = "oz_gemm" if "oz_gemm" in prof else "chol_super_update"
        ms, work, cnt = prof.get(dom, (0.0, 0.0, 0))
        eq = work / (ms * 1e-3) / 1e12 if ms > 0 else 0.0  # float64-equivalent TFLOP/s of the dominant kernel
        tensor_ms, tensor_fl = tensor_total(prof)
        traffic, traffic_src = ncu_traffic()
        if dom == "oz_gemm":
            ach, peak, kname = eq * i8_per_f64, int8_peak, (
                "k_oz_gemm: long-K panel updates of the Cholesky / triangular solves as error-free sliced INT8 products on "
                f"tcgen05 (TMA-fed, TMEM accumulators; {ns} digit planes per operand = {i8_per_f64} INT8 products per "
                "float64 product)")
            peak_src = (f"cuBLASLt INT8 GEMM {NI}^3 through torch._int_mm, best of 5, measured in this run (nominal dense "
                        "INT8 4500 TOP/s; MEASURED_PEAKS.json bf16 burst x 2 = "
                        f"{2 * float(peaks.get('bf16_tflops', 0)):.0f})")
            unit = "TFLOP/s"
        else:
            ach, peak, kname = eq, fp64_peak, "k_chol_super_update (FP64 DMMA m8n8k4 tile GEMM, long-K)"
            peak_src = f"cuBLAS DGEMM {N}^3 through torch.matmul, best of 5, measured in this run"
            unit = "TFLOP/s"
        roofline = {"bound": "tensor", "kernel": kname, "achieved": round(ach, 3), "peak": round(peak, 3), "unit": unit,
                    "frac": round(ach / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                    "ops": ("INT8 multiply-adds x 2 executed by tcgen05.mma.kind::i8 = float64-equivalent flops x "
                            f"{i8_per_f64}") if dom == "oz_gemm" else "float64 flops",
                    "fp64_equivalent": {"achieved": round(eq, 3), "unit": "TFLOP/s", "dgemm_peak": round(fp64_peak, 3),
                                        "ratio_to_dgemm_peak": round(eq / fp64_peak, 4),
                                        "dgemm_peak_source": f"cuBLAS DGEMM {N}^3 through torch.matmul, best of 5, this run"},
                    "peak_source": peak_src, "launches": cnt, "avg_launch_ms": round(ms / max(cnt, 1), 4),
                    "flops_per_launch": work / max(cnt, 1) * (i8_per_f64 if dom == "oz_gemm" else 1),
                    "note": (f"per-launch events are recorded on a repeat of the K timed steps (same work, same {n_streams} "
                             "concurrent solve streams, batches pipelined): a launch's event interval includes the time it "
                             "shares the SMs with the other streams' kernels; 'serialized' repeats one step on one stream"),
                    "all_tensor_kernels": {"fp64_equivalent_achieved": round(tensor_fl / (tensor_ms * 1e-3) / 1e12, 3) if tensor_ms else 0,
                                           "share_of_step": round(tensor_ms * 1e-3 / t_prof, 4),
                                           "fp64_equivalent_flops_per_step": tensor_fl / max(args.steps, 1),
                                           "fp64_equivalent_achieved_wall": round(tensor_fl / t_res / 1e12, 3),
                                           "ratio_to_dgemm_peak_wall": round(tensor_fl / t_res / 1e12 / fp64_peak, 4)},
                    "ms_per_step_with_events": round(1e3 * t_prof / args.steps, 3),
                    "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src, "stages": stages}
        if prof_serial:
            t_ser = sum(v[0] for v in prof_serial.values()) * 1e-3
            ms_s, work_s, cnt_s = prof_serial.get(dom, (0.0, 0.0, 0))
            eq_s = work_s / (ms_s * 1e-3) / 1e12 if ms_s > 0 else 0.0
            ach_s = eq_s * (i8_per_f64 if dom == "oz_gemm" else 1)
            sm, sf = tensor_total(prof_serial)
            roofline["serialized"] = {"achieved": round(ach_s, 3), "frac": round(ach_s / peak, 4), "launches": cnt_s,
                                      "avg_launch_ms": round(ms_s / max(cnt_s, 1), 4),
                                      "fp64_equivalent_achieved": round(eq_s, 3),
                                      "all_tensor_kernels": {"fp64_equivalent_achieved": round(sf / (sm * 1e-3) / 1e12, 3) if sm else 0},
                                      "stages": stage_table(prof_serial, max(t_ser, 1e-9))}
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu:
            ncpu = 16  # ~16 s of CPU work on 16 cores (a 4x4 corner of the block: InStamp-pair blocks are reused as on the GPU)
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
        if world > 1:
            line["gather_verified"] = gather_verified
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_strong(args, world, rank, barrier):
    """Strong scaling of ONE block: contiguous strips of whole 2x2 stamp-group rows per rank, every rank holds
    the block's pixels and tables (the one-stamp halo of InStamps comes for free), seam InStamp-pair blocks are
    recomputed on both sides, zero-initialised full cubes are sum-reduced on rank 0.  Rank 0 also coadds the unsharded
    block (outside the clock) and checks the reduced cube against it: identical away from the strip seams, and equal to
    float32 rounding of the seam overlap-adds on them."""
    import torch
    import torch.distributed as dist

    from pyimcom_b200 import _lib
    from pyimcom_b200 import pyimcom_croutines as G
    from pyimcom_b200.coadd import GpuBlock
    from pyimcom_b200.psfovl_host import PSFTables
    from pyimcom_b200.shard import assign_stamp_groups, reduce_cube

    blk = make_block(0, n1=args.strong_n1)
    cfg = blk.cfg
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
    mine = assign_stamp_groups(cfg.n1P, world, rank)
    gb = GpuBlock(blk, tab).prepare(stamps=mine)

    def step():
        gb.reset_maps()
        gb.reset_cache()
        gb.run()
        cube = gb.out_map.clone()
        return reduce_cube(cube, world)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        cube = step()
    e1.record()
    barrier()
    launches = int(_lib.launch_count() - n0)
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = float(t[0])
    # every rank's own cube also travels to rank 0, which recomputes each strip itself: are the ranks' GPUs bit-consistent?
    from pyimcom_b200.shard import gather_cube

    gb.reset_maps()
    gb.reset_cache()
    gb.run()
    parts = gather_cube(gb.out_map, world, rank)
    if rank == 0:
        strips_same = []
        for r in range(world):
            g_r = GpuBlock(blk, tab).prepare(stamps=assign_stamp_groups(cfg.n1P, world, r)).run()
            strips_same.append(bool(torch.equal(g_r.out_map, parts[r])))
            del g_r
        full = GpuBlock(blk, tab).prepare()
        full.run()
        ref = full.out_map
        same = (cube == ref)
        scale = float(ref.abs().max())
        # rows of the map that belong to a strip seam: the fade borders of the stamps on either side overlap there
        fk, n2 = cfg.fade_kernel, cfg.n2
        per = (len(range(1, cfg.n1P + 1, 2)) + world - 1) // world
        seam = torch.zeros(ref.shape[-2], dtype=torch.bool, device=ref.device)
        for r in range(1, world):
            y = 2 * per * r * n2
            seam[max(0, y - 0):y + 2 * fk] = True
        off_seam_equal = bool(same[..., ~seam, :].all())
        max_dev = float((cube - ref).abs().max()) / scale
        rows_diff = torch.nonzero((cube != ref).any(dim=-1).any(dim=0).any(dim=0)).flatten().tolist()
        n_total = cfg.n1P * cfg.n1P
        line = {"metric": METRIC, "value": n_total * cfg.n2**2 * args.steps / t, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(cfg, n_total, {"parallelism": f"one block, {world} strips of 2x2 stamp-group rows",
                                                     "stamps_per_rank": [len(assign_stamp_groups(cfg.n1P, world, r))
                                                                         for r in range(world)]}),
                "gpu_launches": launches,
                "strong_verified": {"identical_off_seams": off_seam_equal, "max_rel_deviation": max_dev,
                                    "seam_rows": int(seam.sum()), "rows_that_differ": rows_diff[:24],
                                    "strips_recomputed_on_rank0_identical": strips_same,
                                    "ok": bool(off_seam_equal and max_dev < 1e-6)}}
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
    ap.add_argument("--strong", action="store_true", help="strong scaling: ONE block split into strips of stamp groups")
    ap.add_argument("--strong-n1", type=int, default=14, help="n1 of the block of the --strong run (n1P = n1 + 2)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
