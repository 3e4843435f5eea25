"""Seeded inputs shared by the golden generator, the oracle tests and the GPU parity tests.

The known-answer inputs restate those of the reference's own tests (file:line given per builder);
the block cases are the parity configurations of SURVEY 8d scaled to sizes the CPU oracle finishes
in seconds.
"""

import types

import numpy as np

from pyimcom_b200.synth import ARCSEC, StampConfig, SynthBlock

GETW_FH = (-0.5, -0.3137, -0.01, 0.0, 0.2718, 0.49, 0.5)


def interp_inputs():
    """tests/pyimcom/test_routine.py:8-63: sin grid (2,64,32), 100 quasi-random points, 3x(12,20) grid."""
    nx, ny, N = 32, 64, 10
    npts = N * N
    infunc = np.sin(np.linspace(0, 200, 2 * nx * ny)).reshape((2, ny, nx))
    x_, _ = np.modf(np.arange(npts) / np.sqrt(5))
    x_ *= 40
    y_, _ = np.modf(np.arange(npts) * 2 / np.sqrt(5))
    y_ *= 40
    xs_, ys_ = x_.copy(), y_.copy()
    for i in range(1, N):
        for j in range(i):
            xs_[i * N + j] = xs_[j * N + i]
            ys_[i * N + j] = ys_[j * N + i]
    npi, nxo, nyo = 3, 12, 20
    xpos = np.zeros((npi, nxo))
    ypos = np.zeros((npi, nyo))
    for i in range(npi):
        xpos[i, :] = np.linspace(2 + i, nx - 2 - i, nxo)
        ypos[i, :] = np.linspace(2 + i, ny - 2 - i, nyo)
    return infunc, x_, y_, xs_, ys_, xpos, ypos


def kernel_toy():
    """tests/pyimcom/test_routine.py:66-108: 33x33 -> 25x25 Gaussian-overlap toy, scaled by 0.7."""
    sigma, m1, n1 = 4.0, 25, 33
    gi = np.arange(n1, dtype=np.float64)
    y = np.repeat(gi, n1)
    x = np.tile(gi, n1)
    go = 5 + 0.25 * np.arange(m1)
    yout = np.repeat(go, m1)
    xout = np.tile(go, m1)
    A = np.exp(-1.0 / sigma**2 * ((x[:, None] - x[None, :]) ** 2 + (y[:, None] - y[None, :]) ** 2))
    mBhalf = np.exp(-1.0 / sigma**2 * ((x[None, :] - xout[:, None]) ** 2 + (y[None, :] - yout[:, None]) ** 2))
    return A * 0.7, mBhalf * 0.7, 0.7


def reduced_inputs(m=257, nv=4, seed=7):
    """Random but well-posed node statistics for build_reduced_T_wrap (routine.py:487-588).

    Built the way CholKernel._call_multi_kappa builds them (lakernel.py:361-380) from a small random
    SPD system, so that bracket selection and bisection branches are exercised on both sides.
    """
    rng = np.random.default_rng(seed)
    n = 24
    G = rng.standard_normal((n, 3 * n))
    A = G @ G.T / (3 * n)
    w, Q = np.linalg.eigh(A)
    A = (Q * (w * np.logspace(-6, 0, n))) @ Q.T
    B = rng.standard_normal((m, n)) @ A * rng.uniform(0.2, 1.0, size=(m, 1))
    Cn = 1.3 * np.max(np.einsum("ai,ai->a", B @ np.linalg.pinv(A), B))
    kap = np.array([1e-5, 1e-3, 1e-1, 3.0])[:nv]
    Tpi = np.stack([np.linalg.solve(A + k * Cn * np.eye(n), B.T).T for k in kap])
    Dp = np.einsum("ai,pai->ap", B, Tpi)
    Npq = np.einsum("pai,qai->apq", Tpi, Tpi)
    Epq = np.zeros((m, nv, nv))
    for p in range(nv):
        for q in range(p + 1):
            Epq[:, q, p] = Epq[:, p, q] = Dp[:, q] - kap[p] * Cn * Npq[:, p, q]
    return Npq.ravel().copy(), (Dp.ravel() / Cn).copy(), (Epq.ravel() / Cn).copy(), kap, 1e-4, 0.6


def la_outst(kappaC, uctarget=1e-4, sigmamax=0.5, iterative=False):
    """tests/pyimcom/test_la.py:46-230: 6x6 circulant cosine system, 16 output pixels, duck-typed outst."""
    N = 6
    idx = np.arange(N)
    d = 2 * np.pi * (idx[:, None] - idx[None, :]) / N
    A = np.zeros((N, N))
    for k in range(1, N // 2 + 1):
        A += np.cos(k * d) / k / N
    mBhalf = np.zeros((1, 16, N))
    for i in range(N):
        for j in range(16):
            _d = 2 * np.pi * (i - 0.4 * j) / N
            for k in range(1, N // 2 + 1):
                mBhalf[0, j, i] += np.cos(k * _d) / k / N
    cfg = types.SimpleNamespace(n2f=4, n_out=1, kappaC_arr=np.asarray(kappaC, dtype=np.float64), uctarget=uctarget,
                                sigmamax=sigmamax, linear_algebra="x", fade_kernel=0)
    if iterative:
        cfg.instamp_pad = 2.0 * ARCSEC
        cfg.dtheta = 0.11 / 3600.0
        cfg.iter_rtol = 1e-2
        cfg.iter_max = 8
    outst = types.SimpleNamespace(blk=types.SimpleNamespace(cfg=cfg), inpix_cumsum=np.array([N]), sysmata=A,
                                  mhalfb=mBhalf, outovlc=np.array([A[0, 0]]))
    if iterative:
        outst.yx_val = [np.linspace(0, 6, 16), np.zeros(16)]
        outst.iny_val = np.zeros(N)
        outst.inx_val = np.linspace(0, N - 1, N)
    return outst


# name -> (kernel class, kappaC, la_outst kwargs)
LA_CASES = {
    "eigen1": ("EigenKernel", [1e-2], dict(sigmamax=0.5)),
    "eigen3": ("EigenKernel", [1e-4, 1e-3, 1e-2], dict(sigmamax=1.0)),
    "chol1": ("CholKernel", [1e-2], dict(sigmamax=0.5)),
    "chol3": ("CholKernel", [1e-4, 1e-3, 1e-2], dict(sigmamax=1.0)),
    "iter2": ("IterKernel", [1e-3, 1e-2], dict(sigmamax=1.0, iterative=True)),
    "iter1": ("IterKernel", [1e-3], dict(sigmamax=1.0, iterative=True)),
}

# Synthetic blocks pushed through the reference's own OutStamp path (make_golden.py).
_MINI = dict(n1=2, n2=8, dtheta_arcsec=0.04, fade_kernel=1, postage_pad=0, npixpsf=16, oversamp=4,
             instamp_pad_arcsec=0.3, n_inframe=3)
BLOCK_CASES = {
    # config 1 analogue: Cholesky, one kappa node
    "chol1": dict(cfg=_MINI, n_image=2, seed=1, kernel="Cholesky", kappaC=[5e-4], stamps=[(1, 1), (2, 2)],
                  store_ab=True, store_ti64=True),
    # Cholesky, three kappa nodes -> build_reduced_T_wrap
    "chol3": dict(cfg=_MINI, n_image=2, seed=2, kernel="Cholesky", kappaC=[1e-5, 1e-4, 1e-3], stamps=[(1, 2)]),
    # config 2 analogue: Eigen with per-pixel kappa bisection
    "eigen3": dict(cfg=_MINI, n_image=2, seed=3, kernel="Eigen", kappaC=[1e-5, 1e-4, 1e-3], stamps=[(2, 1)]),
    "eigen1": dict(cfg=_MINI, n_image=2, seed=3, kernel="Eigen", kappaC=[5e-4], stamps=[(2, 1)]),
    # config 3 analogue: Iterative, kappa = 0, no fade
    "iter0": dict(cfg=dict(_MINI, fade_kernel=0, instamp_pad_arcsec=0.25), n_image=3, seed=4, kernel="Iterative",
                  kappaC=[0.0], stamps=[(1, 1)], iter_rtol=1.5e-3, iter_max=30),
    "iter2": dict(cfg=dict(_MINI, fade_kernel=0, instamp_pad_arcsec=0.25), n_image=3, seed=4, kernel="Iterative",
                  kappaC=[1e-4, 1e-3], stamps=[(1, 1)], iter_rtol=1.5e-3, iter_max=30),
    # config 5 analogue: two output PSFs + PSF splitting (doubled overlap sampling), flat penalty on
    "nout2split": dict(cfg=dict(_MINI, n_out=2, psfsplit=True, sigmatarget=0.95, sigmatarget_extra=(1.1,),
                                outpsf_extra=("GAUSSIAN",), flat_penalty=1e-7), n_image=2, seed=5, kernel="Cholesky",
                       kappaC=[5e-4], stamps=[(1, 1)], store_ab=True),
    # 4x4 block with padding so that interior stamps have full 3x3 neighbourhoods and four PSF groups
    "pad4": dict(cfg=dict(_MINI, n1=2, postage_pad=1, n_inframe=2), n_image=3, seed=6, kernel="Cholesky",
                 kappaC=[5e-4], stamps=[(2, 2), (3, 2)], store_ab=True),
    # EmpirKernel (lakernel.py:747-805): empirical weights, exact U/C
    "empir": dict(cfg=dict(_MINI, n1=2, postage_pad=1, n_inframe=2, instamp_pad_arcsec=0.25), n_image=3, seed=8,
                  kernel="Empirical", kappaC=[5e-4], stamps=[(2, 2), (3, 3)]),
    # truncated PSF support (npixpsf=12) makes A + kappa*I indefinite -> the eigen-shift repair branch
    # of CholKernel._cholesky_wrapper (lakernel.py:262-279) fires in the reference
    "repair": dict(cfg=dict(_MINI, n1=2, postage_pad=1, n_inframe=2, npixpsf=12), n_image=3, seed=6, kernel="Cholesky",
                   kappaC=[5e-4], stamps=[(2, 2)]),
}


def make_block(spec):
    kw = dict(spec["cfg"])
    kw["kappaC_arr"] = np.asarray(spec["kappaC"], dtype=np.float64)
    kw["linear_algebra"] = spec["kernel"]
    for k in ("iter_rtol", "iter_max"):
        if k in spec:
            kw[k] = spec[k]
    cfg = StampConfig(**kw)
    return SynthBlock(cfg, n_image=spec["n_image"], seed=spec["seed"])


# ---------------------------------------------------------------------------------------------------
# block output assembly: seeded block maps as they stand after coadd_output_stamps
# ---------------------------------------------------------------------------------------------------
OUTPUT_CASES = {  # name -> (config overrides, n_inimage, pad_sides of the block, is_final)
    "final_nopad": (dict(n1=2, n2=8, fade_kernel=2, postage_pad=1, n_out=2, n_inframe=3), 4, "", True),
    "final_BL": (dict(n1=2, n2=8, fade_kernel=2, postage_pad=1, n_out=1, n_inframe=2), 3, "BL", True),
    "final_all_fk1": (dict(n1=3, n2=6, fade_kernel=1, postage_pad=2, n_out=1, n_inframe=1, outmaps="USK"), 2, "BTLR", True),
    "intermediate": (dict(n1=2, n2=8, fade_kernel=2, postage_pad=1, n_out=1, n_inframe=2), 3, "", False),
}


def output_case(name):
    """(cfg, maps, n_inimage, pad_sides, is_final): float32 block maps with the value ranges of real ones plus the
    corners of the encoding (zeros, negatives, values that saturate the 16-bit codes)."""
    from pyimcom_b200.synth import StampConfig

    over, n_inimage, pad_sides, is_final = OUTPUT_CASES[name]
    cfg = StampConfig(**over)
    rng = np.random.default_rng(sum(map(ord, name)))
    side = cfg.NsideP + 2 * cfg.fade_kernel
    q = (cfg.n_out, side, side)
    maps = {
        "out_map": rng.standard_normal((cfg.n_out, cfg.n_inframe, side, side)),
        "T_weightmap": rng.uniform(0.0, 0.4, (cfg.n_out, n_inimage, cfg.n1P, cfg.n1P)),
        "UC_map": 10.0 ** rng.uniform(-9.0, 0.2, q),
        "Sigma_map": 10.0 ** rng.uniform(-3.5, 3.5, q),
        "kappa_map": 10.0 ** rng.uniform(-14.0, -1.0, q),
        "Tsum_map": 1.0 + 0.05 * rng.standard_normal(q),
        "Neff_map": rng.uniform(0.3, 8.0, q),
    }
    for k in ("UC_map", "Sigma_map", "kappa_map", "Tsum_map", "Neff_map"):  # encoding corners
        maps[k][:, 5, 7] = 0.0
        maps[k][:, 6, 7] = -3.0
        maps[k][:, 7, 7] = 1e30
        maps[k][:, 8, 7] = 1e-40
        maps[k][:, 9, 7] = 1.0
    return cfg, {k: v.astype(np.float32) for k, v in maps.items()}, n_inimage, pad_sides, is_final


def codes_match(got, want, max_frac=0.005):
    """Log-integer quality-map codes: identical except for at most max_frac of the pixels, and those by one count
    (the float32 log10 of the host that produced `want` is not necessarily correctly rounded; see k_compress_map)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.dtype == want.dtype and got.shape == want.shape, (got.dtype, want.dtype, got.shape, want.shape)
    d = np.abs(got.astype(np.int64) - want.astype(np.int64))
    return d.max() <= 1 and (d > 0).mean() <= max_frac


def code_mismatches_are_ties(got, want, x, coef):
    """Integer output must be identical unless the reference's own float32 formula is ambiguous at that pixel.

    Block.compress_map evaluates floor(coef * log10(x) + 0.5) in float32 (coadd.py:2128).  NumPy's float32 log10 is the
    host libm's log10f (glibc: up to 2 ulp, not correctly rounded), the device takes the correctly rounded float32
    logarithm; the product and the sum round once more each.  A code may therefore differ -- by one count -- only where
    the exact value y = coef * log10(x) + 0.5 lies within that float32 evaluation error of an integer:
        |y - rint(y)| <= 2.5 ulp32(log10 x) * |coef| + ulp32(coef * log10 x).
    Asserts exactly that for EVERY mismatching pixel (and that nothing differs by more than one count); returns
    (number of mismatches, number of pixels inside the ambiguity band)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.dtype == want.dtype and got.shape == want.shape == np.shape(x)
    d = got.astype(np.int64) - want.astype(np.int64)
    x64 = np.clip(np.asarray(x, dtype=np.float32), np.float32(1e-32), None).astype(np.float64)
    lg = np.log10(x64)
    y = coef * lg + 0.5
    band = 2.5 * np.spacing(np.abs(lg).astype(np.float32)).astype(np.float64) * abs(coef) \
        + np.spacing(np.abs(coef * lg).astype(np.float32)).astype(np.float64)
    tie = np.abs(y - np.rint(y)) <= band
    mism = d != 0
    assert np.abs(d).max(initial=0) <= 1, "a 16-bit code differs by more than one count"
    assert np.all(tie[mism]), f"{int((mism & ~tie).sum())} codes differ where the float32 formula is not ambiguous"
    return int(mism.sum()), int(tie.sum())


# ---------------------------------------------------------------------------------------------------
# input-pixel partitioning
# ---------------------------------------------------------------------------------------------------
PARTITION_CASES = {  # name -> (config overrides, detector side, sp_res, centre of the block on the detector, rotation)
    "small_cells": (dict(n1=4, n2=8, postage_pad=1), 600, 40, (301.3, 287.9), 0.31),
    "uneven_cells": (dict(n1=2, n2=24, postage_pad=0), 1000, 90, (520.4, 610.2), -1.1),
    "edge_of_detector": (dict(n1=4, n2=8, postage_pad=1), 600, 40, (8.0, 590.0), 2.0),
    "not_relevant": (dict(n1=2, n2=8, postage_pad=0), 600, 40, (-900.0, 200.0), 0.0),
}


def partition_case(name):
    """(cfg, outpix, mask_a, mask_b, use_instamps, indata, sca_nside, sp_res): a rotated, slightly distorted pixel map in
    place of the WCS composition, two random masks (the reference ANDs its permanent / cosmic-ray / file masks), a few
    postage stamps switched off."""
    from pyimcom_b200.synth import StampConfig

    over, sca, sp_res, ctr, rot = PARTITION_CASES[name]
    cfg = StampConfig(dtheta_arcsec=0.04, fade_kernel=1, n_inframe=2, **over)
    rng = np.random.default_rng(sum(map(ord, name)) + 7)
    s = 0.11 / 0.04  # output pixels per detector pixel
    cs, sn = s * np.cos(rot), s * np.sin(rot)
    mid = cfg.NsideP / 2.0 - 0.5

    def outpix(inxys):
        d = np.asarray(inxys, dtype=np.float64) - np.asarray(ctr)
        dx, dy = d[:, 0], d[:, 1]
        x = cs * dx - sn * dy + 3e-5 * dx * dy + mid
        y = sn * dx + cs * dy - 2e-5 * dx * dx + mid
        return np.stack([x, y], axis=1)

    ns = cfg.n1P + 2
    use = np.ones((ns, ns), dtype=bool)
    use[0, 0] = use[ns // 2, ns // 2 + 1] = use[ns - 1, 1] = False
    mask_a = rng.random((sca, sca)) < 0.93
    mask_b = rng.random((sca, sca)) < 0.97
    indata = rng.standard_normal((cfg.n_inframe, sca, sca)).astype(np.float32)
    return cfg, outpix, mask_a, mask_b, use, indata, sca, sp_res


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configurations at their stated sizes (tests/golden/full_*.npz are made by the reference itself:
# tests/golden/make_golden_full.py; the GPU tests compare the CUDA path with them directly)
# ---------------------------------------------------------------------------------------------------
T_SHAPE = dict(n1=4, n2=25, dtheta_arcsec=0.04, fade_kernel=1, postage_pad=0, npixpsf=42, oversamp=6,
               instamp_pad_arcsec=0.8, n_inframe=5, uctarget=1e-6, sigmamax=0.5)
P4_SHAPE = dict(n1=2, n2=32, dtheta_arcsec=0.0390625, postage_pad=1, npixpsf=48, oversamp=8, n_inframe=6,
                uctarget=1e-6, sigmamax=0.5)
SIG3 = (0.85, 0.95, 1.05)
SIG6 = (0.85, 0.9, 0.95, 1.0, 1.05, 1.1)
FULL_CASES = {
    # configs[0]: tests-shaped block, CholKernel (n ~ 1.5 k, m = 729)
    "cfg1": dict(cfg=dict(T_SHAPE, linear_algebra="Cholesky", kappaC_arr=[5e-4]), n_image=3, seed=12345, sig=SIG3,
                 stamp=(2, 2)),
    # configs[1]: the same block, EigenKernel with the per-pixel kappa bisection (nbis = 13)
    "cfg2": dict(cfg=dict(T_SHAPE, linear_algebra="Eigen", kappaC_arr=[1e-5, 1e-4, 1e-3]), n_image=3, seed=12345,
                 sig=SIG3, stamp=(2, 2)),
    # configs[2]: paper-4 Iter stamp (FADE = 0, m = 1024, INPAD 0.6", kappa = 0, rtol 1.5e-3, 30 iterations, n ~ 2.8 k)
    "cfg3": dict(cfg=dict(P4_SHAPE, fade_kernel=0, instamp_pad_arcsec=0.6, linear_algebra="Iterative", kappaC_arr=[0.0],
                          iter_rtol=1.5e-3, iter_max=30), n_image=6, seed=2024, sig=SIG6, stamp=(2, 2)),
    # the same stamp with kappa/C = 1: the well-posed variant (10-13 iterations), where CG iteration counts must be
    # identical (at kappa/C <= 0.1 the recurrence runs 20-30 iterations and a 1e-15 relative perturbation of A already
    # changes the reference's own iteration counts: tests/test_oracle_golden.py::test_cg_sensitivity)
    "cfg3k": dict(cfg=dict(P4_SHAPE, fade_kernel=0, instamp_pad_arcsec=0.6, linear_algebra="Iterative",
                           kappaC_arr=[1.0], iter_rtol=1.5e-3, iter_max=30), n_image=6, seed=2024, sig=SIG6,
                  stamp=(2, 2)),
    # configs[3]: the paper-4 stamp of bench.py (FADE = 3, m = 1444, INPAD 1.24", kappa/C = 6e-4, n ~ 6.2 k)
    "p4": dict(cfg=dict(P4_SHAPE, fade_kernel=3, instamp_pad_arcsec=1.24, linear_algebra="Cholesky", kappaC_arr=[6e-4]),
               n_image=6, seed=1000, sig=SIG6, stamp=(2, 2)),
    # configs[4]: n_out = 3 with PSF splitting on Roman-like PSFs (obscured Airy (x) jitter Gaussian (x) pixel tophat,
    # circularly truncated short-range PSF), three Gaussian targets
    "cfg5": dict(cfg=dict(T_SHAPE, linear_algebra="Cholesky", kappaC_arr=[5e-4], n_out=3, psfsplit=True, sigmatarget=0.85,
                          sigmatarget_extra=(0.93, 1.02), outpsf_extra=("GAUSSIAN", "GAUSSIAN")), n_image=3, seed=777,
                 sig=(0.30, 0.34, 0.38), psf_kind="airy", stamp=(2, 2)),
}
# strides of the stored sub-samples (coprime with the tile sizes of the kernels)
FULL_SUB = dict(a=(37, 41), b=(29, 31), t_rows=61)


def make_full_block(name):
    spec = FULL_CASES[name]
    cfg = StampConfig(**spec["cfg"])
    return SynthBlock(cfg, n_image=spec["n_image"], seed=spec["seed"], psf_sigmas=spec["sig"], star=True,
                      psf_kind=spec.get("psf_kind", "gauss"))


def full_errors(s, g):
    """Deviations of one coadded OutStamp `s` (reference attribute names; a GpuOutStamp or an OracleOutStamp after
    post_kernel / perform_coaddition) from the reference-made golden `g` = np.load(full_<case>.npz):
    max|d| / max|ref| per array, UC absolute.  Index work (inpix_cumsum) must be identical and is asserted here."""
    sa, sb, tr = FULL_SUB["a"], FULL_SUB["b"], FULL_SUB["t_rows"]

    def rel(a, b):
        b = np.asarray(b, dtype=np.float64)
        return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-300))

    assert np.array_equal(np.asarray(s.inpix_cumsum), g["inpix_cumsum"])
    T = np.asarray(s.T)
    T64 = T.astype(np.float64)
    e = {
        "outovlc": rel(s.outovlc, g["outovlc"]),
        "sysmata": max(rel(s.sysmata[::sa[0], ::sa[1]], g["sysmata_sub"]), rel(np.diag(s.sysmata), g["sysmata_diag"])),
        "mhalfb": rel(s.mhalfb[:, ::sb[0], ::sb[1]], g["mhalfb_sub"]),
        "T": max(rel(T[:, ::tr, :], g["T_rows"]), rel(T[:, ::sb[0], ::sb[1]], g["T_sub"])),
        # every element of T enters these two: float64 sums of the float32 matrix, normalised by the sum of |T|
        "T_rowsum": float(np.abs(T64.sum(axis=-1) - g["T_rowsum"]).max() / np.abs(T64).sum(axis=-1).max()),
        "T_colabs": rel(np.abs(T64).sum(axis=-2), g["T_colabs"]),
        "UC_abs": float(np.abs(np.asarray(s.UC, dtype=np.float64) - g["UC"]).max()),
    }
    for nm in ("Sigma", "kappa", "outimage", "Tsum_stamp", "Tsum_inpix", "Neff"):
        e[nm] = rel(getattr(s, nm), g[nm])
    if "Ti64_sub" in g.files and getattr(s, "Ti64", None) is not None:
        e["Ti64"] = rel(np.asarray(s.Ti64)[:, ::sb[0], ::sb[1]], g["Ti64_sub"])
    return e
