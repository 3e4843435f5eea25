"""Generate tests/golden/full_p4block.npz: the REFERENCE ITSELF on a whole paper-4-shaped block (build container only).

    python tests/golden/make_golden_block.py

The 4 x 4-stamp block of tests/cases.py::FULL_CASES['p4'] (bench.py's block with n1 = 2: 16 OutStamps of n ~ 6.2 k input
pixels, m = 1444) goes through the reference's own InStamp -> OutStamp -> SysMatA/SysMatB -> CholKernel ->
_perform_coaddition path, stamp by stamp in the reference's traversal order (oracle/refhost.run_block), and the per-stamp
results are overlap-added into the block maps exactly as Block._output_stamp_wrapper does (coadd.py:1976-1994: float32
maps, `+=` of the faded stamp arrays).  Stored: out_map, UC / Sigma / kappa / Tsum / Neff maps, T_weightmap (0.9 MB).
About four minutes on the 8 vCPUs of the build container.
"""

import contextlib
import io
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import refhost  # noqa: E402


def main():
    assert refhost.available(), "needs /root/reference (build container)"
    blk = cases.make_full_block("p4")
    cfg = blk.cfg
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = refhost.run_block(blk, cfg.linear_algebra, cfg.kappaC_arr)
    side = cfg.NsideP + 2 * cfg.fade_kernel
    f32 = np.float32
    maps = {"out_map": np.zeros((cfg.n_out, cfg.n_inframe, side, side), dtype=f32),
            "T_weightmap": np.zeros((cfg.n_out, blk.n_inimage, cfg.n1P, cfg.n1P), dtype=f32)}
    for nm in ("UC_map", "Sigma_map", "kappa_map", "Tsum_map", "Neff_map"):
        maps[nm] = np.zeros((cfg.n_out, side, side), dtype=f32)
    for (j, i) in blk.stamp_order():  # coadd.py:2056-2069 order, :1976-1994 accumulation
        d = res[(j, i)]
        b, l = (j - 1) * cfg.n2, (i - 1) * cfg.n2
        sl = (slice(b, b + cfg.n2f), slice(l, l + cfg.n2f))
        maps["out_map"][(slice(None), slice(None)) + sl] += d["outimage"]
        maps["UC_map"][(slice(None),) + sl] += d["UC"]
        maps["Sigma_map"][(slice(None),) + sl] += d["Sigma"]
        maps["kappa_map"][(slice(None),) + sl] += d["kappa"]
        maps["Tsum_map"][(slice(None),) + sl] += d["Tsum_inpix"]
        maps["Neff_map"][(slice(None),) + sl] += d["Neff"]
        maps["T_weightmap"][:, :, j - 1, i - 1] = d["Tsum_stamp"]
    np.savez_compressed(os.path.join(HERE, "full_p4block.npz"), **maps)
    print(f"wrote full_p4block.npz: {len(res)} stamps, reference time {time.perf_counter() - t0:.0f} s, "
          f"{os.path.getsize(os.path.join(HERE, 'full_p4block.npz')) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
