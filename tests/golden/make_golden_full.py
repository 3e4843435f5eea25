"""Generate tests/golden/full_*.npz: the REFERENCE ITSELF at BASELINE.json's stated sizes (build container only).

    python tests/golden/make_golden_full.py [case ...]

Same mechanism as make_golden.py (pyimcom.{routine,lakernel,psfutil,coadd} imported verbatim from /root/reference behind
oracle/refhost.py; nothing copied), on the blocks of tests/cases.py::FULL_CASES: config 1 (CholKernel, n ~ 1.5 k),
config 2 (EigenKernel + kappa bisection), config 3 (IterKernel, n ~ 2.8 k; and its well-posed kappa/C = 1 variant),
the paper-4 stamp of bench.py (n ~ 6.2 k, m = 1444) and config 5 (n_out = 3, PSF splitting, Roman-like PSFs).  One
OutStamp per case.  The matrices are too large to commit whole, so what is stored is: everything small in full
(inpix_cumsum, outovlc, UC, Sigma, kappa, outimage, Tsum_*, Neff), strided sub-samples of sysmata / mhalfb / the
float64 solution / T (strides in cases.FULL_SUB, coprime with the kernels' tile sizes), and two float64 checksums of
every T element (row sums, column sums of |T|).  Times on the 8 vCPUs of the build container: config 1 12 s,
paper-4 stamp 80 s.
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


def reduce_stamp(d):
    """The committed subset of one OutStamp's arrays (see module docstring)."""
    sa, sb, tr = cases.FULL_SUB["a"], cases.FULL_SUB["b"], cases.FULL_SUB["t_rows"]
    out = {k: np.asarray(d[k]) for k in ("inpix_cumsum", "outovlc", "UC", "Sigma", "kappa", "outimage", "Tsum_stamp",
                                          "Tsum_inpix", "Neff")}
    out["sysmata_sub"] = d["sysmata"][::sa[0], ::sa[1]].copy()
    out["sysmata_diag"] = np.diag(d["sysmata"]).copy()
    out["mhalfb_sub"] = d["mhalfb"][:, ::sb[0], ::sb[1]].copy()
    if "Ti64" in d:
        out["Ti64_sub"] = d["Ti64"][:, ::sb[0], ::sb[1]].copy()
    T = d["T"]
    out["T_rows"] = T[:, ::tr, :].copy()
    out["T_sub"] = T[:, ::sb[0], ::sb[1]].copy()
    T64 = T.astype(np.float64)
    out["T_rowsum"] = T64.sum(axis=-1)
    out["T_colabs"] = np.abs(T64).sum(axis=-2)
    return out


def main(names):
    assert refhost.available(), "needs /root/reference (build container)"
    for name in names:
        spec = cases.FULL_CASES[name]
        blk = cases.make_full_block(name)
        cfg = blk.cfg
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings(record=True) as wlist:
            warnings.simplefilter("always")
            res = refhost.run_block(blk, cfg.linear_algebra, cfg.kappaC_arr, stamps={spec["stamp"]}, only=True)
        dt = time.perf_counter() - t0
        out = reduce_stamp(res[spec["stamp"]])
        # repair-branch activations of CholKernel._cholesky_wrapper (lakernel.py:262-279 warns once per repaired matrix)
        out["n_repair"] = np.array(sum("holesky" in str(w.message) for w in wlist))
        np.savez_compressed(os.path.join(HERE, f"full_{name}.npz"), **out)
        n = int(out["inpix_cumsum"][-1])
        print(f"wrote full_{name}.npz: n={n} m={cfg.n2f**2} n_out={cfg.n_out} reference time {dt:.1f} s, "
              f"repairs {int(out['n_repair'])}, {os.path.getsize(os.path.join(HERE, f'full_{name}.npz')) / 1e6:.2f} MB",
              flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(cases.FULL_CASES))
