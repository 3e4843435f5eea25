"""GPU parity at BASELINE.json's full sizes (``-m gpu``): the paper-4 stamp shape of bench.py (n ~ 6.2 k selected input
pixels, m = 1444 output pixels, 6 images), where the CPU oracle needs > 1 s per stamp.  One stamp is compared with
the oracle outright (P-f64 / P-f32); the whole batched block is checked through size-independent properties:

* A is exactly symmetric and equals the fused-kernel assembly bit for bit (the cached pair blocks change nothing);
* the solve satisfies its defining equation: T (A + kappa I) = -B/2 to 1e-9 relative, verified with an independent
  float64 product (torch / cuBLAS, not the library's own GEMM);
* a unit point source rendered through each input PSF is recovered with the target PSF's amplitude
  (tests/pyimcom/test_pyimcom.py:943-978 re-expressed on the synthetic block);
* batching is invisible: a stamp solved alone gives the same T as inside a batch of 16.
"""

import os
import sys
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import lakernel as OL  # noqa: E402
from oracle import routines as R  # noqa: E402
from oracle.sysmat import OracleOutStamp  # noqa: E402
from pyimcom_b200 import lakernel as GL  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock, GpuOutStamp  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402


def rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def block():
    blk = bench.make_block(0, n1=2)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
    return blk, tab


def test_paper4_stamp_vs_oracle(block):
    blk, tab = block
    gb = GpuBlock(blk, tab).prepare(stamps=[(2, 2)])
    s = GpuOutStamp(gb, 2, 2)
    assert s.T.shape[-1] > 5000 and s.T.shape[1] == 1444
    R.set_threads(os.cpu_count() or 1)
    o = OracleOutStamp(blk, PSFTables(blk, R.iD5512C, R.gridD5512C, dedup=True), 2, 2)
    o.build_system_matrices()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        k = OL.CholKernel(o)
        k()
    o.post_kernel()
    o.perform_coaddition()
    assert rel(s.sysmata, o.sysmata) < 1e-9 and rel(s.mhalfb, o.mhalfb) < 1e-9  # P-f64
    assert rel(s.Ti64[0], k.f64[0]["Ti"]) < 1e-9  # P-f64 on the pre-cast solution
    for nm in ("T", "Sigma", "kappa", "outimage", "Tsum_stamp", "Tsum_inpix", "Neff"):  # P-f32
        assert rel(getattr(s, nm), getattr(o, nm)) < (2e-6 if nm == "T" else 1e-5), nm
    assert np.abs(s.UC - o.UC).max() < 2e-6


def test_paper4_block_properties(block):
    blk, tab = block
    cfg = blk.cfg
    gb = GpuBlock(blk, tab).prepare()
    fused = GpuBlock(blk, tab, a_cache=False).prepare()
    ks = list(range(len(gb.order)))
    kept = gb.coadd_batch(ks, keep=True)  # all 16 stamps through one batched factorisation (3 solve streams)
    torch.cuda.synchronize()
    kap = float(cfg.kappaC_arr[0]) * float(tab.outovlc[0])
    for q in (0, 5, 15):
        ds, res = kept[q]["ds"], kept[q][0]["res"]
        n, m = ds.n, ds.m
        A = ds.A[:n, :n]
        assert torch.equal(A, A.T)
        assert torch.equal(ds.A, fused.build_system(q)[0].A)
        T = res["Ti64"][:m, :n]
        lhs = T @ A + kap * T  # independent float64 product (cuBLAS)
        mB = ds.mB[0, :m, :n]
        assert float((lhs - mB).abs().max() / mB.abs().max()) < 1e-9
        # padding stays inert: identity block in A, zero columns in T
        assert float(res["Ti64"][:m, n:].abs().max()) == 0.0 if res["Ti64"].shape[1] > n else True
    # batching is invisible
    alone = GpuBlock(blk, tab).prepare(stamps=[gb.order[5]])
    one = alone.coadd_batch([0], keep=True)[0]
    assert rel(one[0]["res"]["Ti64"].cpu().numpy(), kept[5][0]["res"]["Ti64"].cpu().numpy()) < 1e-12
    # star recovery: layer 0 is a unit point source seen through every input PSF (synth.py); the coadd must show the
    # target PSF -- a Gaussian of sigma = cfg.sigmatarget native px, in flux per native pixel -- centred on the star
    maps = gb.download()
    fk = cfg.fade_kernel
    star = maps["out_map"][0, 0].astype(np.float64)
    sx, sy = blk.star_xy[0] + fk, blk.star_xy[1] + fk  # map index = output pixel + fade border
    s_out = cfg.sigmatarget * 0.11 / cfg.dtheta_arcsec
    yy, xx = np.mgrid[0:star.shape[0], 0:star.shape[1]]
    model = np.exp(-0.5 * ((xx - sx) ** 2 + (yy - sy) ** 2) / s_out**2) / (2 * np.pi * cfg.sigmatarget**2)
    win = (np.abs(xx - sx) < 4 * s_out) & (np.abs(yy - sy) < 4 * s_out)
    assert np.abs(star - model)[win].max() < 0.02 * model.max()
    assert abs((star * xx)[win].sum() / star[win].sum() - sx) < 0.05
    assert abs((star * yy)[win].sum() / star[win].sum() - sy) < 0.05


def test_tests_shaped_block_multikappa_batched():
    """BASELINE configs[0] geometry (n ~ 1.5 k, m = 729) with three kappa nodes: all 16 stamps x 3 nodes go through the
    batched factorisation together (48 systems, 3 launch sequences on 3 streams); two stamps are compared with the
    oracle, including the discrete bracket / branch decisions of build_reduced_T_wrap."""
    from pyimcom_b200.synth import StampConfig, SynthBlock

    cfg = StampConfig(n1=4, n2=25, dtheta_arcsec=0.04, fade_kernel=1, postage_pad=0, npixpsf=42, oversamp=6,
                      instamp_pad_arcsec=0.8, n_inframe=4, linear_algebra="Cholesky",
                      kappaC_arr=np.array([1e-5, 1e-4, 1e-3]), uctarget=1e-6, sigmamax=0.5)
    blk = SynthBlock(cfg, n_image=3, seed=12345, psf_sigmas=(0.85, 0.95, 1.05), star=True)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
    gb = GpuBlock(blk, tab).prepare()
    kept = gb.coadd_batch(list(range(len(gb.order))), keep=True)
    torch.cuda.synchronize()
    otab = PSFTables(blk, R.iD5512C, R.gridD5512C, dedup=True)
    for q in (3, 10):
        j, i = gb.order[q]
        o = OracleOutStamp(blk, otab, j, i)
        o.build_system_matrices()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            k = OL.CholKernel(o)
            k()
        ko, res, ds = kept[q][0]["ko"], kept[q][0]["res"], kept[q]["ds"]
        assert np.array_equal(ko.extras["iv"].cpu().numpy(), k.f64[0]["iv"])  # P-discrete
        assert np.array_equal(ko.extras["branch"].cpu().numpy(), k.f64[0]["branch"])
        assert rel(res["Ti64"][:ds.m, :ds.n].cpu().numpy(), k.f64[0]["Ti"]) < 1e-9  # P-f64
        o.post_kernel()
        o.perform_coaddition()  # fades o.T in place, as coadd.py:1321-1324 does
        assert rel(res["T32"][:ds.m, :ds.n].cpu().numpy(), o.T[0]) < 2e-6
        assert rel(res["outimage"].cpu().numpy().reshape(o.outimage[0].shape), o.outimage[0]) < 1e-5
