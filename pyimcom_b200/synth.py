"""Synthetic, WCS-free block generator (host-side harness; NOT on the hot path).

The reference feeds ``OutStamp`` from FITS images through astropy WCS (coadd.py:51-655), which is
out of scope here (a9/c4 are the first rows in scope).  This module fabricates the same
data structures the reference's ``InStamp``/``OutStamp`` read -- per-image ``pix_count / x_val /
y_val / data`` binned into input postage stamps exactly like ``InImage.partition_pixels``
(coadd.py:198-358) -- from affine image placements and analytic multi-Gaussian PSFs, so that the
same seeded block can be pushed through (1) the reference itself (tests/golden/make_golden.py,
build container only), (2) the CPU oracle and (3) the CUDA path.

Attribute names deliberately mirror the reference (``cfg.n2f``, ``blk.inimages[k].pix_count`` ...)
so the duck-typed objects can be handed straight to ``pyimcom.coadd.InStamp/OutStamp``.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

ARCSEC = math.pi / 648000.0  # radians, config.py:87
DEGREE = math.pi / 180.0
PIXSCALE_NATIVE_ARCSEC = 0.11  # config.py:97


@dataclass
class StampConfig:
    """The hot-path subset of pyimcom.config.Config (config.py:412-449, 502-594)."""

    n1: int = 4
    n2: int = 25
    dtheta_arcsec: float = 0.04
    fade_kernel: int = 1
    postage_pad: int = 0
    npixpsf: int = 42
    oversamp: int = 6
    instamp_pad_arcsec: float = 0.8
    n_out: int = 1
    n_inframe: int = 4
    linear_algebra: str = "Cholesky"
    kappaC_arr: np.ndarray = field(default_factory=lambda: np.array([5e-4]))
    uctarget: float = 1e-6
    sigmamax: float = 0.5
    iter_rtol: float = 1.5e-3
    iter_max: int = 30
    flat_penalty: float = 0.0
    psfsplit: bool = False
    sigmatarget: float = 0.9265  # output Gaussian sigma in native pixels (EXTRASMOOTH)
    sigmatarget_extra: tuple = ()
    outpsf: str = "GAUSSIAN"
    outpsf_extra: tuple = ()
    psf_circ: bool = False
    psf_norm: bool = False
    amp_penalty: tuple = (0.0, 0.0)
    tempfile: object = None
    use_filter: int = 2
    outmaps: str = "USKTN"
    no_qlt_ctrl: bool = False

    def __post_init__(self):
        self.kappaC_arr = np.asarray(self.kappaC_arr, dtype=np.float64)
        self.dtheta = self.dtheta_arcsec / 3600.0  # degrees, config.py:502
        self.instamp_pad = self.instamp_pad_arcsec * ARCSEC  # radians, config.py:562
        self.n1P = self.n1 + 2 * self.postage_pad
        self.n2f = self.n2 + 2 * self.fade_kernel
        self.Nside = self.n1 * self.n2
        self.NsideP = self.Nside + 2 * self.postage_pad * self.n2
        # PSFGrp.setup / PSFOvl.setup (psfutil.py:568-613, 1065-1089)
        self.nsamp = self.npixpsf * self.oversamp - 1
        self.nfft = self.npixpsf * self.oversamp * 2
        self.dscale = PIXSCALE_NATIVE_ARCSEC / self.oversamp / self.dtheta_arcsec
        self.nsamp_ovl = 2 * self.nsamp + 1 if self.psfsplit else self.nsamp
        self.nc_ovl = self.nsamp_ovl // 2
        self.rpix_search = (self.instamp_pad / ARCSEC) / (self.dtheta * 3600.0)  # coadd.py:923


class GaussMixPSF:
    """Analytic PSF = sum of (possibly offset, elliptical) Gaussians, in native-pixel units."""

    def __init__(self, comps):
        # comps: list of (amp, x0, y0, sx, sy, theta)
        self.comps = [tuple(float(v) for v in c) for c in comps]

    def __call__(self, x, y):
        out = np.zeros(np.broadcast(x, y).shape)
        for amp, x0, y0, sx, sy, th in self.comps:
            c, s = math.cos(th), math.sin(th)
            u = c * (x - x0) + s * (y - y0)
            v = -s * (x - x0) + c * (y - y0)
            out += amp * np.exp(-0.5 * ((u / sx) ** 2 + (v / sy) ** 2)) / (2.0 * math.pi * sx * sy)
        return out

    def oversampled(self, npix, oversamp):
        """(npix*oversamp)^2 array centred at ((n-1)/2,(n-1)/2), flux per oversampled pixel."""
        n = npix * oversamp
        g = (np.arange(n) - (n - 1) / 2.0) / oversamp
        return self(g[None, :], g[:, None]) / oversamp**2


Q_FILTER_NATIVE = (1.155, 1.456, 1.250, 1.021, 0.834, 0.689, 0.491, 1.009, 0.000, 1.159, 1.685)  # config.py:91
OBSC = 0.31  # config.py:94


class AiryPSF:
    """Roman-like synthetic PSF: obscured Airy disc (x) Gaussian jitter (x) square pixel tophat, circularly truncated.

    The oversampled image is the spot ``OutPSF.psf_simple_airy`` draws (psfutil.py:149-224): amplitude
    ``J0 + J2`` of the full aperture minus ``obsc^2`` times that of the obscuration, squared, ``pi / (4 ldp^2
    (1 - obsc^2))`` normalisation, smoothing applied as a transfer function on the padded FFT grid; here the spot may
    also be decentred by a sub-pixel offset and is cut to zero outside ``rcut`` native pixels (the short-range PSF
    G^(S) that PSF splitting hands to PSFGrp, coadd.py:565-569).  ``__call__`` evaluates the same function off the
    grid (cubic spline through a finer rendering), for the point-source layer.
    """

    def __init__(self, ldp=1.25, obsc=OBSC, sigma=0.3, tophat=1.0, x0=0.0, y0=0.0, rcut=None):
        self.ldp, self.obsc, self.sigma, self.tophat = float(ldp), float(obsc), float(sigma), float(tophat)
        self.x0, self.y0, self.rcut = float(x0), float(y0), rcut
        self._fine = None

    def _render(self, n, ov):
        """n x n samples at 1/ov native pixels, centred at ((n-1)/2, (n-1)/2) + (x0, y0) * ov, flux per sample."""
        from scipy.special import jv

        ldp, sig, th, ob = self.ldp * ov, self.sigma * ov, self.tophat * ov, self.obsc
        kp = 1 + int(np.ceil(th + 6 * sig))
        npad = n + 2 * kp
        g = np.arange(npad) - (npad - 1) / 2.0
        r = np.hypot(g[None, :] - self.x0 * ov, g[:, None] - self.y0 * ov) / ldp
        amp = jv(0, np.pi * r) + jv(2, np.pi * r) - ob**2 * (jv(0, np.pi * r * ob) + jv(2, np.pi * r * ob))
        img = np.square(amp) * (np.pi / (4.0 * ldp**2 * (1.0 - ob**2)))
        u = np.fft.fftfreq(npad)
        ux, uy = u[None, :npad // 2 + 1], u[:, None]
        mtf = np.exp(-2.0 * np.pi**2 * sig**2 * (ux**2 + uy**2)) * np.sinc(ux * th) * np.sinc(uy * th)
        img = np.fft.irfft2(np.fft.rfft2(img) * mtf, s=(npad, npad))[kp:-kp, kp:-kp]
        if self.rcut is not None:
            c = np.arange(n) - (n - 1) / 2.0
            img = img * (np.hypot(c[None, :] - self.x0 * ov, c[:, None] - self.y0 * ov) < self.rcut * ov)
        return img

    def oversampled(self, npix, oversamp):
        return self._render(npix * oversamp, oversamp)

    def __call__(self, x, y):
        from scipy.ndimage import map_coordinates

        ov, half = 8, 24
        if self._fine is None:
            self._fine = self._render(2 * half * ov + 1, ov) * ov**2  # flux per native pixel, centre on a sample
        x, y = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64))
        c = half * ov
        return map_coordinates(self._fine, [y.ravel() * ov + c, x.ravel() * ov + c], order=3, mode="constant",
                               cval=0.0).reshape(x.shape)


class SynthImage:
    """Duck-typed stand-in for coadd.InImage (the attributes InStamp/PSFGrp touch)."""

    def __init__(self, idsca, M, t, psf: GaussMixPSF, cfg: StampConfig):
        self.idsca = idsca
        self.M = np.asarray(M, dtype=np.float64)  # native px -> output px (2x2)
        self.t = np.asarray(t, dtype=np.float64)
        self.Minv = np.linalg.inv(self.M)
        self.psf = psf
        self.cfg = cfg
        self._psf_arr = None

    # coadd.py:156-172 (WCS round trip replaced by the affine map)
    def outpix2world2inpix(self, outxys):
        return (np.asarray(outxys, dtype=np.float64) - self.t) @ self.Minv.T

    def inpix2outpix(self, inxys):
        return np.asarray(inxys, dtype=np.float64) @ self.M.T + self.t

    # coadd.py:540: oversampled PSF image at a position (position independent here)
    def get_psf_pos(self, point=None, use_shortrange=False):
        if self._psf_arr is None:
            self._psf_arr = self.psf.oversampled(self.cfg.npixpsf, self.cfg.oversamp)
        return self._psf_arr

    def partition(self, layers_rng, star_xy=None):
        """Bin native pixels into input stamps, as partition_pixels does (coadd.py:198-358)."""
        cfg = self.cfg
        ns = cfg.n1P + 2
        lo = -cfg.n2 - 0.5
        hi = cfg.NsideP + cfg.n2 - 0.5
        corners = np.array([[lo, lo], [lo, hi], [hi, lo], [hi, hi]])
        cin = self.outpix2world2inpix(corners)
        u0, v0 = np.floor(cin.min(axis=0)).astype(int) - 1
        u1, v1 = np.ceil(cin.max(axis=0)).astype(int) + 1
        uu, vv = np.meshgrid(np.arange(u0, u1 + 1), np.arange(v0, v1 + 1))
        inxy = np.stack([uu.ravel(), vv.ravel()], axis=1).astype(np.float64)
        oxy = self.inpix2outpix(inxy)
        ok = (oxy[:, 0] > lo) & (oxy[:, 0] < hi) & (oxy[:, 1] > lo) & (oxy[:, 1] < hi)
        inxy, oxy = inxy[ok], oxy[ok]
        i_st = ((oxy[:, 0] - lo) // cfg.n2).astype(int)
        j_st = ((oxy[:, 1] - lo) // cfg.n2).astype(int)
        key = j_st * ns + i_st
        order = np.argsort(key, kind="stable")
        key, oxy, inxy = key[order], oxy[order], inxy[order]
        counts = np.bincount(key, minlength=ns * ns)
        self.pix_count = counts.reshape(ns, ns).astype(np.uint32)
        self.max_count = int(counts.max()) if counts.size else 0
        start = np.concatenate([[0], np.cumsum(counts)])
        within = np.arange(key.size) - start[key]
        self.x_val = np.zeros((ns, ns, self.max_count))
        self.y_val = np.zeros((ns, ns, self.max_count))
        jj, ii = np.divmod(key, ns)
        self.x_val[jj, ii, within] = oxy[:, 0]
        self.y_val[jj, ii, within] = oxy[:, 1]
        # layers: layer 0 = unit point source seen through this image's PSF; the rest white noise
        vals = np.zeros((cfg.n_inframe, key.size), dtype=np.float32)
        if star_xy is not None:
            sin = self.outpix2world2inpix(np.asarray(star_xy, dtype=np.float64)[None, :])[0]
            vals[0] = self.psf(inxy[:, 0] - sin[0], inxy[:, 1] - sin[1]).astype(np.float32)
        for k in range(1 if star_xy is not None else 0, cfg.n_inframe):
            vals[k] = layers_rng.standard_normal(key.size).astype(np.float32)
        self.data = np.zeros((cfg.n_inframe, ns, ns, self.max_count), dtype=np.float32)
        self.data[:, jj, ii, within] = vals


class SynthInStamp:
    """coadd.InStamp (coadd.py:656-792) without the PSF-group refcount plumbing."""

    def __init__(self, blk, j_st, i_st):
        self.blk, self.j_st, self.i_st = blk, j_st, i_st
        self.pix_count = np.array([im.pix_count[j_st, i_st] for im in blk.inimages], dtype=np.uint32)
        self.pix_cumsum = np.cumsum([0] + list(self.pix_count), dtype=np.uint32)
        n = int(self.pix_cumsum[-1])
        self.y_val = np.empty(n)
        self.x_val = np.empty(n)
        self.data = np.empty((blk.cfg.n_inframe, n), dtype=np.float32)
        for k, im in enumerate(blk.inimages):
            a, b, c = int(self.pix_cumsum[k]), int(self.pix_cumsum[k + 1]), int(self.pix_count[k])
            self.y_val[a:b] = im.y_val[j_st, i_st, :c]
            self.x_val[a:b] = im.x_val[j_st, i_st, :c]
            self.data[:, a:b] = im.data[:, j_st, i_st, :c]
        if j_st % 2 == 0 and i_st % 2 == 0:  # coadd.py:709-714
            self.psf_compute_point_pix = [i_st * blk.cfg.n2 - 0.5, j_st * blk.cfg.n2 - 0.5]

    def make_selection(self, pivot=(None, None), radius=None):
        """coadd.py:716-749."""
        if pivot == (None, None) or radius is None:
            return None
        dist_sq = np.zeros(self.x_val.shape[0])
        if pivot[0] is not None:
            dist_sq += np.square(self.x_val - pivot[0])
        if pivot[1] is not None:
            dist_sq += np.square(self.y_val - pivot[1])
        sel = np.array(np.where(dist_sq < radius**2)[0], dtype=np.uint32)
        return sel if sel.shape[0] < self.x_val.shape[0] else None


class _IdentityWCS:
    @staticmethod
    def all_pix2world(xy, origin):  # only used to label the PSF compute point (psfutil.py:832)
        return np.asarray(xy, dtype=np.float64)


class SynthBlock:
    """Duck-typed coadd.Block: cfg, inimages, instamps (+ tables added by psfovl_host)."""

    def __init__(self, cfg: StampConfig, n_image=3, seed=12345, psf_sigmas=(0.85, 0.95, 1.05),
                 rot_deg=3.0, star=True, asym=0.08, psf_kind="gauss"):
        self.cfg = cfg
        self.this_sub = 0
        self.outwcs = _IdentityWCS()
        rng = np.random.default_rng(seed)
        s = PIXSCALE_NATIVE_ARCSEC / cfg.dtheta_arcsec
        self.inimages = []
        self.star_xy = None
        if star:
            mid = 0.5 * cfg.NsideP
            self.star_xy = (mid + 1.37, mid - 2.21)
        for k in range(n_image):
            th = math.radians(rng.uniform(-rot_deg, rot_deg))
            R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
            t = rng.uniform(0.0, 1.0, size=2) * s
            sig = psf_sigmas[k % len(psf_sigmas)]
            ang = rng.uniform(0, math.pi)
            if psf_kind == "airy":
                # Roman-like: lambda/D of the configured filter (a few per cent of defocus-like spread between the
                # exposures), jitter sigma = psf_sigmas[k] native px, decentred by up to 0.05 px, cut at 0.45 npixpsf
                psf = AiryPSF(ldp=Q_FILTER_NATIVE[cfg.use_filter] * (1.0 + 0.02 * math.cos(1.7 * k)), sigma=sig,
                              x0=0.05 * math.cos(ang), y0=0.05 * math.sin(ang), rcut=0.45 * cfg.npixpsf)
            else:
                comps = [(1.0 - asym, 0.0, 0.0, sig, sig * 1.06, ang)]
                if asym:
                    comps.append((asym, 0.6 * math.cos(2.1 * k + 0.3), 0.6 * math.sin(2.1 * k + 0.3), sig * 1.3,
                                  sig * 1.3, 0.0))
                psf = GaussMixPSF(comps)
            im = SynthImage((100 + k, 1 + k), s * R, t, psf, cfg)
            im.partition(rng, self.star_xy)
            self.inimages.append(im)
        self.n_inimage = len(self.inimages)
        ns = cfg.n1P + 2
        self.instamps = [[SynthInStamp(self, j, i) for i in range(ns)] for j in range(ns)]

    def stamp_order(self):
        """OutStamp traversal of coadd_output_stamps: 2x2 groups (coadd.py:2056-2060)."""
        n1P = self.cfg.n1P
        for j in range(1, n1P + 1, 2):
            for i in range(1, n1P + 1, 2):
                for dj in range(2):
                    for di in range(2):
                        if j + dj <= n1P and i + di <= n1P:
                            yield (j + dj, i + di)
