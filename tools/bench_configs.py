#!/usr/bin/env python
"""Parity + timing of the other BASELINE.json configurations on one B200 (the driver's bench.py covers configs[3]):

  1  tests-shaped block, CholKernel            (n2=25, FADE=1, NPIXPSF=42, oversamp=6, INPAD=0.8", kappaC=[5e-4])
  2  same block, EigenKernel, kappa bisection  (kappaC=[1e-5,1e-4,1e-3], nbis=13)
  3  large stamp, IterKernel                   (n2=32, FADE=0, oversamp=8, INPAD=0.6", kappaC=[0], rtol 1.5e-3, 30 its)
  5  n_out=3 with PSF splitting                (config-1 geometry, doubled overlap sampling)

For each: whole-block device time (GpuBlock.run, inputs resident), the CPU oracle on a bounded sample of stamps, and the
parity of the first sampled stamp (P-f64 on A / mBhalf, P-f32 on T / outimage).  One JSON line per config.
"""
import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lakernel as OL  # noqa: E402
from oracle import routines as R  # noqa: E402
from oracle.sysmat import OracleOutStamp  # noqa: E402
from pyimcom_b200 import _lib  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock, GpuOutStamp  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402
from pyimcom_b200.synth import StampConfig, SynthBlock  # noqa: E402

T_SHAPE = dict(n1=4, n2=25, dtheta_arcsec=0.04, fade_kernel=1, postage_pad=0, npixpsf=42, oversamp=6,
               instamp_pad_arcsec=0.8, n_inframe=5, uctarget=1e-6, sigmamax=0.5)
CONFIGS = {
    "1_chol_tests_block": dict(cfg=dict(T_SHAPE, linear_algebra="Cholesky", kappaC_arr=[5e-4]), n_image=3, seed=12345,
                               sig=(0.85, 0.95, 1.05), tol=2e-6),
    "2_eigen_tests_block": dict(cfg=dict(T_SHAPE, linear_algebra="Eigen", kappaC_arr=[1e-5, 1e-4, 1e-3]), n_image=3,
                                seed=12345, sig=(0.85, 0.95, 1.05), tol=5e-5),
    "3_iter_large_stamp": dict(cfg=dict(n1=2, n2=32, dtheta_arcsec=0.0390625, fade_kernel=0, postage_pad=1, npixpsf=48,
                                        oversamp=8, instamp_pad_arcsec=0.6, n_inframe=6, linear_algebra="Iterative",
                                        kappaC_arr=[0.0], iter_rtol=1.5e-3, iter_max=30), n_image=6, seed=2024,
                               # kappa = 0 makes A singular: the reference's own CG result moves by 8e-3 under a 1e-15 relative
                               # perturbation of A (oracle run), so T is only defined to ~1e-2 here
                               sig=(0.85, 0.9, 0.95, 1.0, 1.05, 1.1), tol=3e-2),
    "5_nout3_psfsplit": dict(cfg=dict(T_SHAPE, linear_algebra="Cholesky", kappaC_arr=[5e-4], n_out=3, psfsplit=True,
                                      sigmatarget=0.85, sigmatarget_extra=(0.93, 1.02),
                                      outpsf_extra=("GAUSSIAN", "GAUSSIAN")), n_image=3, seed=777,
                             sig=(0.85, 0.95, 1.05), tol=2e-6),
}
KERN = {"Cholesky": OL.CholKernel, "Eigen": OL.EigenKernel, "Iterative": OL.IterKernel}


def rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    only = sys.argv[1:]
    cores = os.cpu_count() or 1
    R.set_threads(cores)
    for name, spec in CONFIGS.items():
        if only and not any(name.startswith(o) for o in only):
            continue
        cfg = StampConfig(**spec["cfg"])
        blk = SynthBlock(cfg, n_image=spec["n_image"], seed=spec["seed"], psf_sigmas=spec["sig"], star=True)
        tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
        gb = GpuBlock(blk, tab).prepare()
        ns = [gb.plans[ji].n for ji in gb.order]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for _ in range(2):
                gb.reset_maps(); gb.reset_cache(); gb.run()
            torch.cuda.synchronize()
            n0 = _lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                gb.reset_maps(); gb.reset_cache(); gb.run()
            e1.record()
            torch.cuda.synchronize()
            t_gpu = e0.elapsed_time(e1) * 1e-3 / reps
            launches = (_lib.launch_count() - n0) // reps
            # one more pass with per-launch events: device time and algorithmic work per kernel kind
            _lib.profile(1)
            gb.reset_maps(); gb.reset_cache(); gb.run()
            torch.cuda.synchronize()
            prof = {k: {"ms": round(v[0], 3), "launches": v[2],
                        "rate": (round(v[1] / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 and v[1] > 0 else None)}
                    for k, v in _lib.profile_read().items()}
            _lib.profile(0)
            # CPU oracle on a bounded sample + parity of the first stamp
            otab = PSFTables(blk, R.iD5512C, R.gridD5512C, dedup=True)
            sample = gb.order[:2]
            OracleOutStamp(blk, otab, *sample[0]).build_system_matrices()  # tables outside the clock
            t0 = time.perf_counter()
            outs = []
            for (j, i) in sample:
                o = OracleOutStamp(blk, otab, j, i)
                o.build_system_matrices()
                KERN[cfg.linear_algebra](o)()
                o.post_kernel()
                o.perform_coaddition()
                outs.append(o)
            t_cpu = (time.perf_counter() - t0) / len(sample)
            s = GpuOutStamp(gb, *sample[0])
        o = outs[0]
        line = {"config": name, "kernel": cfg.linear_algebra, "stamps": len(gb.order), "n_in_mean": int(np.mean(ns)),
                "m": cfg.n2f**2, "n_out": cfg.n_out, "gpu_s_per_block": t_gpu,
                "gpu_output_px_per_s": len(gb.order) * cfg.n2**2 / t_gpu, "gpu_launches_per_block": int(launches),
                "cpu_s_per_stamp": t_cpu, "cpu_output_px_per_s": cfg.n2**2 / t_cpu, "cpu_cores": cores,
                "parity": {"A_rel": rel(s.sysmata, o.sysmata), "mBhalf_rel": rel(s.mhalfb, o.mhalfb),
                           "T_rel": rel(s.T, o.T), "outimage_rel": rel(s.outimage, o.outimage),
                           "Sigma_rel": rel(s.Sigma, o.Sigma), "UC_abs": float(np.abs(s.UC - o.UC).max())},
                "tol_T": spec["tol"], "per_kind_ms_one_block (rate: GFLOP/s for tensor kinds, GB/s otherwise)": prof}
        line["parity_ok"] = bool(line["parity"]["A_rel"] < 1e-9 and line["parity"]["mBhalf_rel"] < 1e-9
                                 and line["parity"]["T_rel"] < spec["tol"])
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
