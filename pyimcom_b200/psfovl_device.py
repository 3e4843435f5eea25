"""PSF-overlap tables built on the device (the step before the hot path).

Same tables, conventions and bookkeeping as ``psfovl_host.PSFTables`` (which restates ``PSFGrp._sample_psf`` /
``accel_pad_and_rfft2`` / ``PSFOvl._build_psfovl``, psfutil.py:709-795, 942-986, 1177-1294), but everything after the
host has produced the sample coordinates stays in HBM:

* input PSFs are sampled with the device ``iD5512C`` kernel (``b200_dev_iD5512C``);
* the zero-padded ``rfft2`` / ``irfft2`` pair is evaluated as dense DFT products on the FP64 tensor pipe
  (``b200_dev_gemm_nt``): only ``ns`` of the ``nfft`` input samples are non-zero and only the ``2 nc + 1`` central lags of
  the correlation are kept (psfutil.py:1226-1227), so the partial DFTs are four small matrix products per direction
  instead of full-length transforms -- no FFT library, and the products run at DGEMM rate;
* spectra are kept transposed, ``FT[v, u]`` (v: half spectrum along x, u: full spectrum along y), as separate real and
  imaginary planes, which makes every product of the chain an ``A (M,K) . B (N,K)^T`` with K contiguous;
* ``rft_j * conj(rft_i)`` is ``b200_dev_cmul_conj``.

The resulting table sets are ``torch`` tensors on the device; ``coadd._Arena`` lays them out without a host round trip.
Agreement with the NumPy-FFT tables is at the 1e-13 level (tests/test_gpu_parity.py::test_device_tables).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .lakernel import ptr, rup, stream_handle
from .adapter import adapt_block
from .psfovl_host import PSFTables


def _phase(k, j, n):
    """cos, sin of 2 pi k j / n with the integer product reduced mod n first (full float64 accuracy)."""
    r = (np.asarray(k, dtype=np.int64)[:, None] * np.asarray(j, dtype=np.int64)[None, :]) % n
    ang = 2.0 * np.pi * r.astype(np.float64) / n
    return np.cos(ang), np.sin(ang)


def _pad(a, rows, cols):
    out = np.zeros((rows, cols))
    out[: a.shape[0], : a.shape[1]] = a
    return torch.from_numpy(out).cuda()


_DFT_CACHE = {}


class _Spectrum:
    """Transposed half spectra of a stack of PSFs: re, im (n, nvp, nup) float64 on the device."""

    def __init__(self, re, im):
        self.re, self.im = re, im

    @property
    def n(self):
        return self.re.shape[0]


class DeviceTables(PSFTables):
    """Drop-in for psfovl_host.PSFTables whose table sets live on the device."""

    def __init__(self, blk, iC, gridC, dedup=False):
        if not torch.cuda.is_available():
            raise RuntimeError("pyimcom_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        blk = adapt_block(blk)
        cfg = blk.cfg
        ns, nfft, nc = cfg.nsamp, cfg.nfft, cfg.nc_ovl
        nl = cfg.nsamp_ovl      # lags kept per axis: ns, or 2 ns + 1 with PSF splitting (psfutil.py:1075-1083)
        assert 2 * nc + 1 == nl and nl < nfft
        nv = nfft // 2 + 1
        self.ns, self.nfft, self.nv, self.nl = ns, nfft, nv, nl
        self.nsp = rup(ns)      # rows / K extent of the PSF samples
        self.nvp = rup(nv)      # half spectrum along x
        self.nup = rup(nfft)    # full spectrum along y
        self.nlp = rup(nl)      # kept lags
        key = (ns, nfft, nc, nl, torch.cuda.current_device())
        if key not in _DFT_CACHE:  # the partial DFT matrices depend on the geometry only
            x = np.arange(ns)
            v = np.arange(nv)
            u = np.arange(nfft)
            lag = np.arange(nl) - nc
            cvx, svx = _phase(v, x, nfft)       # forward along x: e^{-i 2 pi v x / nfft}
            cuy, suy = _phase(u, x, nfft)       # forward along y
            cdu, sdu = _phase(lag, u, nfft)     # inverse along y at the kept lags
            cdv, sdv = _phase(lag, v, nfft)     # inverse along x at the kept lags (Hermitian half, weights alpha_v)
            alpha = np.full(nv, 2.0)
            alpha[0] = 1.0
            if nfft % 2 == 0:
                alpha[-1] = 1.0
            alpha /= float(nfft) ** 2
            _DFT_CACHE[key] = (_pad(cvx, self.nvp, self.nsp), _pad(-svx, self.nvp, self.nsp),
                               _pad(cuy, self.nup, self.nsp), _pad(suy, self.nup, self.nsp),
                               _pad(cdu, self.nlp, self.nup), _pad(sdu, self.nlp, self.nup),
                               _pad(cdv * alpha[None, :], self.nlp, self.nvp), _pad(sdv * alpha[None, :], self.nlp, self.nvp))
        self.Cx, self.Sxn, self.Cu, self.Su, self.Cd, self.Sd, self.Cv, self.Sv = _DFT_CACHE[key]
        self.amp = None
        if 0.0 not in cfg.amp_penalty:  # psfutil.py:650-671: real factor on the spectra, here in the transposed layout
            uu = np.linspace(0, 1 - 1 / nfft, nfft)
            uu = np.where(uu > 0.5, uu - 1, uu)
            u2 = np.square(uu)
            ut2 = u2[None, : nfft // 2 + 1] + u2[:, None]  # [u (full, along y), v (half, along x)]
            fac = 1.0 + cfg.amp_penalty[0] * np.exp(-2.0 * np.pi**2 * ut2 * (cfg.amp_penalty[1] * cfg.oversamp) ** 2)
            self.amp = _pad(np.ascontiguousarray(fac.T), self.nvp, self.nup)
        super().__init__(blk, iC, gridC, dedup=dedup)

    # ---- dense products on the tensor pipe -------------------------------------------------------------------
    @staticmethod
    def _gemm(A, B, Cm, acc):
        """Cm (M,N) (=, +=, -=) A (M,K) B (N,K)^T for contiguous 2-D views."""
        M, K = A.shape
        N = B.shape[0]
        _lib.dev_gemm_nt(ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(Cm), Cm.stride(0), M, N, K, acc, stream_handle())

    def _forward(self, psf):
        """psf (n, ns, ns) device -> transposed half spectra (n, nvp, nup)."""
        n = psf.shape[0]
        X = torch.zeros((n, self.nsp, self.nsp), dtype=torch.float64, device="cuda")
        X[:, : self.ns, : self.ns] = psf
        re = torch.empty((n, self.nvp, self.nup), dtype=torch.float64, device="cuda")
        im = torch.empty_like(re)
        ytr = torch.empty((self.nvp, self.nsp), dtype=torch.float64, device="cuda")
        yti = torch.empty_like(ytr)
        for k in range(n):
            # along x:  YT[v, y] = sum_x (cos - i sin)[v, x] X[y, x]
            self._gemm(self.Cx, X[k], ytr, 0)
            self._gemm(self.Sxn, X[k], yti, 0)
            # along y:  FT[v, u] = sum_y YT[v, y] (cos - i sin)[u, y]
            self._gemm(ytr, self.Cu, re[k], 0)
            self._gemm(yti, self.Su, re[k], 1)
            self._gemm(yti, self.Cu, im[k], 0)
            self._gemm(ytr, self.Su, im[k], -1)
        return _Spectrum(re, im)

    def _inverse(self, gr, gi):
        """Product spectra (n, nvp, nup) -> central (nl, nl) lags of irfft2, as a contiguous (n, nl, nl) tensor.

        All n tables go through each product together (stacked along M), so the tile grids fill the GPU."""
        n = gr.shape[0]
        out = torch.empty((n, self.nl, self.nl), dtype=torch.float64, device="cuda")
        CH = 32
        for k0 in range(0, n, CH):
            k1 = min(k0 + CH, n)
            nb = k1 - k0
            g_r = gr[k0:k1].reshape(nb * self.nvp, self.nup)
            g_i = gi[k0:k1].reshape(nb * self.nvp, self.nup)
            htr = torch.empty((nb * self.nvp, self.nlp), dtype=torch.float64, device="cuda")
            hti = torch.empty_like(htr)
            # along y:  HT[v, dy] = sum_u G[v, u] (cos + i sin)[dy, u]
            self._gemm(g_r, self.Cd, htr, 0)
            self._gemm(g_i, self.Sd, htr, -1)
            self._gemm(g_r, self.Sd, hti, 0)
            self._gemm(g_i, self.Cd, hti, 1)
            # (v, dy) -> (dy, v) per table, so that v is the contiguous K of the last product
            hr = htr.view(nb, self.nvp, self.nlp).transpose(1, 2).contiguous().view(nb * self.nlp, self.nvp)
            hi = hti.view(nb, self.nvp, self.nlp).transpose(1, 2).contiguous().view(nb * self.nlp, self.nvp)
            # along x (Hermitian half):  ovl[dy, dx] = sum_v alpha_v (Hr cos - Hi sin)[dy, v ; dx, v]
            ovl = torch.empty((nb * self.nlp, self.nlp), dtype=torch.float64, device="cuda")
            self._gemm(hr, self.Cv, ovl, 0)
            self._gemm(hi, self.Sv, ovl, -1)
            out[k0:k1].copy_(ovl.view(nb, self.nlp, self.nlp)[:, : self.nl, : self.nl])
        return out

    def _cmul_conj(self, a: _Spectrum, b: _Spectrum, b_index=None, conj_a=False):
        """a[k] * conj(b[k]) (matching stacks) or a[k] * conj(b[b_index]) (broadcast); conj_a: conj(a[k]) * b[...]."""
        gr, gi = torch.empty_like(a.re), torch.empty_like(a.im)
        per = a.re.shape[1] * a.re.shape[2]
        if b_index is None:
            br, bi, stride = b.re, b.im, per
        else:
            br, bi, stride = b.re[b_index], b.im[b_index], 0
        _lib.dev_cmul_conj(ptr(a.re), ptr(a.im), ptr(br), ptr(bi), per, stride, a.n, -1.0 if conj_a else 1.0, ptr(gr),
                           ptr(gi), stream_handle())
        return gr, gi

    # ---- PSFTables hooks ---------------------------------------------------------------------------------------
    def _sample_in(self, inst, image):
        """psfutil.py:709-795: same coordinates as the host builder, interpolation on the device."""
        cfg = self.cfg
        psf = image.get_psf_pos(None, use_shortrange=True)
        ny, nx = psf.shape
        xctr, yctr = (nx - 1) / 2.0, (ny - 1) / 2.0
        pt = np.asarray(inst.psf_compute_point_pix, dtype=np.float64)
        if cfg.psfsplit:  # linearised distortion from the four cardinal offsets (psfutil.py:739-753)
            card = np.flip(image.outpix2world2inpix(pt[None, :] + np.array([[1, 0], [0, 1], [-1, 0], [0, -1]]) * cfg.oversamp),
                           axis=-1) / 2.0 * cfg.dscale
            yxco = np.tensordot(card[0] - card[2], self.yxo[1], axes=0) + np.tensordot(card[1] - card[3], self.yxo[0], axes=0)
        else:
            xyo = np.flip(self.yxo, axis=0).reshape((2, -1)).T * cfg.dscale
            yxco = image.outpix2world2inpix(xyo + pt)
            yxco -= image.outpix2world2inpix(pt[None, :])
            yxco = np.flip(yxco * cfg.oversamp, axis=-1).T.reshape(2, cfg.nsamp, cfg.nsamp)
        tab = torch.from_numpy(np.pad(psf, 6)).cuda()
        xs = torch.from_numpy(np.ascontiguousarray(yxco[1].ravel() + xctr + 6)).cuda()
        ys = torch.from_numpy(np.ascontiguousarray(yxco[0].ravel() + yctr + 6)).cuda()
        out = torch.zeros(cfg.nsamp**2, dtype=torch.float64, device="cuda")
        _lib.dev_iD5512C(ptr(tab), 1, ny + 12, nx + 12, ptr(xs), ptr(ys), xs.numel(), ptr(out), stream_handle())
        return out.reshape(cfg.nsamp, cfg.nsamp)

    def _finish(self, psf_arr):
        """Circular cut / normalisation / zero-padded rfft2 (psfutil.py:650-671, 942-986) -> _Spectrum."""
        cfg = self.cfg
        if isinstance(psf_arr, np.ndarray):
            psf_arr = torch.from_numpy(np.ascontiguousarray(psf_arr)).cuda()
        if cfg.psf_circ:
            mask = torch.from_numpy((np.hypot(self.yxo[0], self.yxo[1]) < cfg.nsamp // 2 + 0.5).astype(np.float64)).cuda()
            psf_arr = psf_arr * mask
        if cfg.psf_norm:
            psf_arr = psf_arr / psf_arr.sum(dim=(-2, -1), keepdim=True)
        spec = self._forward(psf_arr.contiguous())
        if self.amp is not None:
            spec.re.mul_(self.amp)
            spec.im.mul_(self.amp)
        return spec

    def _build_out(self):
        """Output PSFs (psfutil.py:874-877, 784-794) -> spectra; C = out (*) out at zero lag (psfutil.py:1290)."""
        from .psfovl_host import psf_gaussian

        cfg = self.cfg
        ns = cfg.nsamp
        sig = (cfg.sigmatarget,) + tuple(cfg.sigmatarget_extra)
        assert len(sig) == cfg.n_out
        psf_arr = np.zeros((cfg.n_out, ns, ns))
        for k in range(cfg.n_out):
            orig = psf_gaussian(ns + 1, sig[k] * cfg.oversamp, sig[k] * cfg.oversamp)
            ctr = ns / 2.0
            out = np.zeros((1, ns * ns))
            self.gridC(np.pad(orig, 6), np.ascontiguousarray(self.yxo[None, 1, 0, :] + ctr + 6),
                       np.ascontiguousarray(self.yxo[None, 0, :, 0] + ctr + 6), out)
            psf_arr[k] = out.reshape(ns, ns)
        self.out_rft = self._finish(psf_arr)
        gr, gi = self._cmul_conj(self.out_rft, self.out_rft)
        oo = self._inverse(gr, gi)
        self.outovlc = np.ascontiguousarray(oo[:, cfg.nc_ovl, cfg.nc_ovl].cpu().numpy())

    def group(self, G):
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
        arr = torch.zeros((len(imgs), self.cfg.nsamp, self.cfg.nsamp), dtype=torch.float64, device="cuda")
        if self.dedup:  # (shared tables are sampled at the first group's point whoever asks first: see psfovl_host)
            inst = blk.instamps[0][0]
        for q, k in enumerate(imgs):
            arr[q] = self._sample_in(inst, blk.inimages[k])
        self.grp_rft[G] = self._finish(arr)
        self._cache[key] = self.grp_rft[G]

    def _memo(self, store, skey, ckey, build):
        if skey not in store:
            if self.dedup and ckey in self._cache:
                store[skey] = self._cache[ckey]
            else:
                store[skey] = build()
                self._cache[ckey] = store[skey]
        return store[skey]

    def get_self(self, G):
        self.group(G)

        def build():
            s = self.grp_rft[G]
            grs, gis = [], []
            for j in range(s.n):  # psf_j (*) psf_i for i >= j (triangle order of psfutil.py:1139-1175)
                a = _Spectrum(s.re[j:], s.im[j:])
                gr, gi = self._cmul_conj(a, s, b_index=j, conj_a=True)  # rft[j] * conj(rft[i]), i >= j
                grs.append(gr)
                gis.append(gi)
            return self._inverse(torch.cat(grs), torch.cat(gis))

        return self._memo(self.self_, G, ("self", tuple(self.grp_imgs[G])), build)

    def get_cross(self, G1, G2):
        assert G1 < G2
        self.group(G1)
        self.group(G2)

        def build():
            s1, s2 = self.grp_rft[G1], self.grp_rft[G2]
            grs, gis = [], []
            for j in range(s1.n):
                gr, gi = self._cmul_conj(s2, s1, b_index=j, conj_a=True)  # rft1[j] * conj(rft2[i])
                grs.append(gr)
                gis.append(gi)
            return self._inverse(torch.cat(grs), torch.cat(gis)).view(s1.n, s2.n, self.nl, self.nl)

        return self._memo(self.cross, (G1, G2), ("cross", tuple(self.grp_imgs[G1]), tuple(self.grp_imgs[G2])), build)

    def get_io(self, G):
        self.group(G)

        def build():
            s1, so = self.grp_rft[G], self.out_rft
            grs, gis = [], []
            for j in range(s1.n):
                gr, gi = self._cmul_conj(so, s1, b_index=j, conj_a=True)  # rft1[j] * conj(out[o])
                grs.append(gr)
                gis.append(gi)
            return self._inverse(torch.cat(grs), torch.cat(gis)).view(s1.n, so.n, self.nl, self.nl)

        return self._memo(self.io, G, ("io", tuple(self.grp_imgs[G])), build)
