"""Host-side PSF-overlap table builder (harness for SURVEY 8f row f1; NOT on the timed hot path).

The hot path *reads* oversampled PSF-overlap tables (``PSFOvl.ovl_arr``); building them
(``PSFGrp._sample_psf`` + ``accel_pad_and_rfft2`` + ``PSFOvl._build_psfovl``, psfutil.py:709-795,
942-986, 1244-1294) is the step before it and is the first "next" row of the scope table.  Until that
row moves to the device this module restates it with NumPy FFTs so that synthetic blocks can be
prepared on any host.  The two interpolators are passed in (product seam ``pyimcom_croutines`` on a
GPU box; the CPU oracle inside CPU-only tests).

Table conventions:

* one PSF group per 2x2 InStamp group, anchored at even (j,i) (coadd.py:709-714, psfutil.py:1824);
* ``self_[G]``   : triangle-packed (n_psf(n_psf+1)/2, ns, ns), entry tri(j,i) = psf_j (*) psf_i, j<=i;
* ``cross[(G1,G2)]`` (G1 < G2): (n_psf1, n_psf2, ns, ns);
* ``io[G]``      : (n_psf, n_out, ns, ns);  ``outovlc`` (n_out,) = C.
"""

from __future__ import annotations

import numpy as np

from .adapter import adapt_block
from .synth import StampConfig


def psf_gaussian(n, sigmax, sigmay):
    """Target Gaussian, psfutil.py:117-146."""
    y, x = np.mgrid[(1 - n) / 2 / sigmay:(n - 1) / 2 / sigmay:n * 1j, (1 - n) / 2 / sigmax:(n - 1) / 2 / sigmax:n * 1j]
    return np.exp(-0.5 * (np.square(x) + np.square(y))) / (2.0 * np.pi * sigmax * sigmay)


def tri_index(n_psf, j, i):
    """psfutil.py:1139-1175."""
    assert j <= i
    return (2 * n_psf - j + 1) * j // 2 + i - j


class PSFTables:
    def __init__(self, blk, iC, gridC, dedup=False):
        blk = adapt_block(blk)  # (a reference Block: derived PSFGrp / PSFOvl quantities come from the adapter)
        self.blk = blk
        self.cfg: StampConfig = blk.cfg
        self.iC, self.gridC = iC, gridC
        self.dedup = dedup
        cfg = self.cfg
        ns = cfg.nsamp
        # PSFGrp.yxo (psfutil.py:599-602)
        self.yxo = np.mgrid[(1 - ns) / 2:(ns - 1) / 2:ns * 1j, (1 - ns) / 2:(ns - 1) / 2:ns * 1j]
        self.grp_rft = {}
        self.grp_imgs = {}
        self.self_ = {}
        self.cross = {}
        self.io = {}
        self._cache = {}
        self._build_out()

    # ---- sampling (psfutil.py:709-795) ----
    def _sample_in(self, inst, image):
        cfg = self.cfg
        psf = image.get_psf_pos(None, use_shortrange=True)
        ny, nx = psf.shape
        xctr, yctr = (nx - 1) / 2.0, (ny - 1) / 2.0
        pt = np.asarray(inst.psf_compute_point_pix, dtype=np.float64)
        if cfg.psfsplit:
            card = np.flip(image.outpix2world2inpix(pt[None, :] + np.array([[1, 0], [0, 1], [-1, 0], [0, -1]]) * cfg.oversamp),
                           axis=-1) / 2.0 * cfg.dscale
            yxco = np.tensordot(card[0] - card[2], self.yxo[1], axes=0) + np.tensordot(card[1] - card[3], self.yxo[0], axes=0)
        else:
            xyo = np.flip(self.yxo, axis=0).reshape((2, -1)).T * cfg.dscale
            yxco = image.outpix2world2inpix(xyo + pt)
            yxco -= image.outpix2world2inpix(pt[None, :])
            yxco = np.flip(yxco * cfg.oversamp, axis=-1).T.reshape(2, cfg.nsamp, cfg.nsamp)
        out = np.zeros((1, cfg.nsamp**2))
        self.iC(np.pad(psf, 6).reshape((1, ny + 12, nx + 12)), np.ascontiguousarray(yxco[1].ravel() + xctr + 6),
                np.ascontiguousarray(yxco[0].ravel() + yctr + 6), out)
        return out.reshape(cfg.nsamp, cfg.nsamp)

    def _finish(self, psf_arr):
        """circular cut / normalisation / padded rfft2 / amplitude penalty (psfutil.py:650-671, 942-986)."""
        cfg = self.cfg
        if cfg.psf_circ:
            psf_arr = psf_arr * (np.hypot(self.yxo[0], self.yxo[1]) < cfg.nsamp // 2 + 0.5)
        if cfg.psf_norm:
            psf_arr = psf_arr / psf_arr.sum(axis=(-2, -1))[:, None, None]
        rft = np.fft.rfft2(psf_arr, s=(cfg.nfft, cfg.nfft))
        if 0.0 not in cfg.amp_penalty:
            nfft = cfg.nfft
            u = np.linspace(0, 1 - 1 / nfft, nfft)
            u = np.where(u > 0.5, u - 1, u)
            u2 = np.square(u)
            ut2 = u2[None, :nfft // 2 + 1] + u2[:, None]
            rft = rft * (1.0 + cfg.amp_penalty[0] * np.exp(-2.0 * np.pi**2 * ut2 * (cfg.amp_penalty[1] * cfg.oversamp) ** 2))
        return rft

    def _ovl(self, rft):
        """irfft2 + roll + crop (psfutil.py:1226-1227)."""
        nc = self.cfg.nc_ovl
        return np.roll(np.fft.irfft2(rft), nc, axis=(-2, -1))[..., :2 * nc + 1, :2 * nc + 1]

    def _build_out(self):
        cfg = self.cfg
        ns = cfg.nsamp
        sig = (cfg.sigmatarget,) + tuple(cfg.sigmatarget_extra)
        assert len(sig) == cfg.n_out
        psf_arr = np.zeros((cfg.n_out, ns, ns))
        for k in range(cfg.n_out):  # psfutil.py:874-877, 784-794
            orig = psf_gaussian(ns + 1, sig[k] * cfg.oversamp, sig[k] * cfg.oversamp)
            ctr = ns / 2.0
            out = np.zeros((1, ns * ns))
            self.gridC(np.pad(orig, 6), np.ascontiguousarray(self.yxo[None, 1, 0, :] + ctr + 6),
                       np.ascontiguousarray(self.yxo[None, 0, :, 0] + ctr + 6), out)
            psf_arr[k] = out.reshape(ns, ns)
        self.out_rft = self._finish(psf_arr)
        oo = self._ovl(self.out_rft * self.out_rft.conjugate())
        self.outovlc = np.ascontiguousarray(oo[:, cfg.nc_ovl, cfg.nc_ovl])  # psfutil.py:1290

    # ---- groups ----
    def group(self, G):
        """PSF group anchored at even InStamp index G=(j,i): psfutil.py:797-851."""
        if G in self.grp_rft:
            return
        blk = self.blk
        inst = blk.instamps[G[0]][G[1]]
        ns_side = self.cfg.n1P + 2
        use = np.zeros(blk.n_inimage, dtype=bool)
        for dj in range(2):
            for di in range(2):
                if G[0] + dj < ns_side and G[1] + di < ns_side:
                    use |= blk.instamps[G[0] + dj][G[1] + di].pix_count.astype(bool)
        imgs = [k for k in range(blk.n_inimage) if use[k]]
        self.grp_imgs[G] = imgs
        key = ("grp", tuple(imgs))
        if self.dedup and key in self._cache:
            self.grp_rft[G] = self._cache[key]
            return
        arr = np.zeros((len(imgs), self.cfg.nsamp, self.cfg.nsamp))
        if self.dedup:
            # shared tables must not depend on which group happened to ask first (ranks that coadd different strips of one
            # block would otherwise hold tables that differ in the last bit): always sample at the first group's point
            inst = blk.instamps[0][0]
        for q, k in enumerate(imgs):
            arr[q] = self._sample_in(inst, blk.inimages[k])
        self.grp_rft[G] = self._finish(arr)
        self._cache[key] = self.grp_rft[G]

    def get_self(self, G):
        if G not in self.self_:
            self.group(G)
            key = ("self", tuple(self.grp_imgs[G]))
            if self.dedup and key in self._cache:
                self.self_[G] = self._cache[key]
                return self.self_[G]
            rft = self.grp_rft[G]
            n_psf = rft.shape[0]
            parts = [self._ovl(rft[j] * rft[j:].conjugate()) for j in range(n_psf)]
            self.self_[G] = np.ascontiguousarray(np.concatenate(parts, axis=0))
            self._cache[key] = self.self_[G]
        return self.self_[G]

    def get_cross(self, G1, G2):
        assert G1 < G2
        if (G1, G2) not in self.cross:
            self.group(G1)
            self.group(G2)
            key = ("cross", tuple(self.grp_imgs[G1]), tuple(self.grp_imgs[G2]))
            if self.dedup and key in self._cache:
                self.cross[(G1, G2)] = self._cache[key]
                return self.cross[(G1, G2)]
            r1, r2 = self.grp_rft[G1], self.grp_rft[G2]
            self.cross[(G1, G2)] = np.ascontiguousarray(np.stack([self._ovl(r1[j] * r2.conjugate()) for j in range(r1.shape[0])]))
            self._cache[key] = self.cross[(G1, G2)]
        return self.cross[(G1, G2)]

    def get_io(self, G):
        if G not in self.io:
            self.group(G)
            key = ("io", tuple(self.grp_imgs[G]))
            if self.dedup and key in self._cache:
                self.io[G] = self._cache[key]
                return self.io[G]
            r1 = self.grp_rft[G]
            self.io[G] = np.ascontiguousarray(np.stack([self._ovl(r1[j] * self.out_rft.conjugate()) for j in range(r1.shape[0])]))
            self._cache[key] = self.io[G]
        return self.io[G]

    def grp_index(self, G, k_img):
        """idx_blk2grp (psfutil.py:821-829)."""
        return self.grp_imgs[G].index(k_img)

    def table_ii_ref(self, Ga, ka, Gb, kb):
        """(table set, index into its leading axes, flip) for psf(ka@Ga) (*) psf(kb@Gb) as a function of (p_a - p_b).

        Same group: triangle entry, flipped in both axes when group index a > b (psfutil.py:1652-1665).
        Different groups: dense table of the ordered pair; the reversed order is the flipped table
        (identical to what PSFOvl(Gb,Ga) would hold up to FFT rounding).
        """
        if Ga == Gb:
            t = self.get_self(Ga)
            n_psf = len(self.grp_imgs[Ga])
            a, b = self.grp_index(Ga, ka), self.grp_index(Ga, kb)
            return (t, (tri_index(n_psf, a, b),), False) if a <= b else (t, (tri_index(n_psf, b, a),), True)
        if Ga < Gb:
            t = self.get_cross(Ga, Gb)
            return t, (self.grp_index(Ga, ka), self.grp_index(Gb, kb)), False
        t = self.get_cross(Gb, Ga)
        return t, (self.grp_index(Gb, kb), self.grp_index(Ga, ka)), True

    def table_ii(self, Ga, ka, Gb, kb):
        """(table, flip): the 2-D table selected by table_ii_ref."""
        t, idx, flip = self.table_ii_ref(Ga, ka, Gb, kb)
        return t[idx], flip


def anchor(ji):
    """SysMatA.ji_st2psf (psfutil.py:1803-1824)."""
    return (ji[0] >> 1 << 1, ji[1] >> 1 << 1)
