"""CPU oracle of system-matrix assembly and the coaddition tail (TEST INFRASTRUCTURE).

Restates, for one output postage stamp:

* ``OutStamp.__init__ / _process_input_stamps``            coadd.py:846-977
* ``PSFOvl._call_ii_self / _call_ii_cross / _call_io_cross`` psfutil.py:1401-1732
* ``OutStamp._build_system_matrices`` (assembly)           coadd.py:1002-1085
* ``OutStamp.trapezoid / _perform_coaddition``             coadd.py:1221-1363

on top of ``oracle.routines`` and a ``PSFTables`` object (pyimcom_b200/psfovl_host.py) holding the
PSF-overlap tables.  The reference's reference-counted caches (SysMatA/SysMatB) only bound CPU RAM
and do not change values beyond FFT/flip rounding, so every block is computed here
directly from the rule set of SURVEY App. B.
"""

from itertools import combinations

import numpy as np

from . import routines as R


def anchor(ji):
    return (ji[0] >> 1 << 1, ji[1] >> 1 << 1)


class OracleOutStamp:
    def __init__(self, blk, tables, j_st, i_st, ii_cache=None):
        """ii_cache: optional dict shared by the OutStamps of a block -- the reference's SysMatA cache of InStamp-pair
        blocks (psfutil.py:1764-2092), so that a timed CPU run does not redo what the reference would reuse."""
        self.blk, self.tab, self.j_st, self.i_st = blk, tables, j_st, i_st
        self.ii_cache = ii_cache
        cfg = blk.cfg
        self.ji_st_in_s = [(j_st + dj, i_st + di) for dj in range(-1, 2) for di in range(-1, 2)]
        self.bottom = (j_st - 1) * cfg.n2
        self.top = self.bottom + cfg.n2 - 1
        self.left = (i_st - 1) * cfg.n2
        self.right = self.left + cfg.n2 - 1
        fk = cfg.fade_kernel
        self.yx_val = np.mgrid[self.bottom - fk:self.top + fk + 1, self.left - fk:self.right + fk + 1]
        # coadd.py:917-977
        self.instamps, self.selections = [], []
        self.inpix_count = np.zeros(9, dtype=np.uint32)
        r = cfg.rpix_search
        for idx, ji in enumerate(self.ji_st_in_s):
            inst = blk.instamps[ji[0]][ji[1]]
            xp = [self.left - 0.5, None, self.right + 0.5][ji[1] - i_st + 1]
            yp = [self.bottom - 0.5, None, self.top + 0.5][ji[0] - j_st + 1]
            sel = inst.make_selection((xp, yp), r)
            self.instamps.append(inst)
            self.selections.append(sel)
            self.inpix_count[idx] = inst.pix_cumsum[-1] if sel is None else sel.shape[0]
        self.inpix_cumsum = np.cumsum([0] + list(self.inpix_count), dtype=np.uint32)
        pick = lambda a, s: a if s is None else a[..., s]  # noqa: E731
        self.iny_val = np.hstack([pick(i.y_val, s) for i, s in zip(self.instamps, self.selections)])
        self.inx_val = np.hstack([pick(i.x_val, s) for i, s in zip(self.instamps, self.selections)])
        self.indata = np.hstack([pick(i.data, s) for i, s in zip(self.instamps, self.selections)])

    # ---- psfutil.py:1401-1495 + 1597-1732: one InStamp-pair block of A ----
    def _ii_block(self, st1, st2):
        if self.ii_cache is not None:
            key = (st1.j_st, st1.i_st, st2.j_st, st2.i_st)
            if key not in self.ii_cache:
                self.ii_cache[key] = self._ii_block_compute(st1, st2)
            return self.ii_cache[key]
        return self._ii_block_compute(st1, st2)

    def _ii_block_compute(self, st1, st2):
        cfg, tab = self.blk.cfg, self.tab
        same = st1 is st2
        G1, G2 = anchor((st1.j_st, st1.i_st)), anchor((st2.j_st, st2.i_st))
        res = np.zeros((int(st1.pix_cumsum[-1]), int(st2.pix_cumsum[-1])))
        ddx = st1.x_val[:, None] - st2.x_val[None, :]
        ddx /= cfg.dscale
        ddx += cfg.nc_ovl
        ddy = st1.y_val[:, None] - st2.y_val[None, :]
        ddy /= cfg.dscale
        ddy += cfg.nc_ovl
        nso = cfg.nsamp_ovl
        tab.group(G1)
        tab.group(G2)
        n_in = len(tab.grp_imgs[G1]) if G1 == G2 else (len(tab.grp_imgs[G1]) * len(tab.grp_imgs[G2])) ** 0.5
        for j_im in range(self.blk.n_inimage):
            if st1.pix_count[j_im] == 0:
                continue
            for i_im in range(j_im if same else 0, self.blk.n_inimage):
                if st2.pix_count[i_im] == 0:
                    continue
                table, flip = tab.table_ii(G1, j_im, G2, i_im)
                if flip:
                    table = np.flip(table)
                sl = np.s_[int(st1.pix_cumsum[j_im]):int(st1.pix_cumsum[j_im + 1]),
                           int(st2.pix_cumsum[i_im]):int(st2.pix_cumsum[i_im + 1])]
                out = np.zeros((1, int(st1.pix_count[j_im]) * int(st2.pix_count[i_im])))
                fn = R.iD5512C_sym if (same and j_im == i_im) else R.iD5512C
                fn(np.ascontiguousarray(np.pad(table, 6)).reshape((1, nso + 12, nso + 12)),
                   np.ascontiguousarray(ddx[sl].ravel() + 6), np.ascontiguousarray(ddy[sl].ravel() + 6), out)
                res[sl] = out.reshape(int(st1.pix_count[j_im]), int(st2.pix_count[i_im]))
                if cfg.flat_penalty != 0.0:  # psfutil.py:1483-1486, 1705-1708
                    res[sl] -= cfg.flat_penalty / n_in
                    if j_im == i_im:
                        res[sl] += cfg.flat_penalty
                if same and j_im < i_im:
                    res[sl[1], sl[0]] = res[sl].T
        return res

    # ---- psfutil.py:1497-1595: one InStamp's columns of mBhalf ----
    def _io_block(self, st1, selection):
        cfg, tab = self.blk.cfg, self.tab
        G = anchor((st1.j_st, st1.i_st))
        io = tab.get_io(G)
        m = cfg.n2f**2
        x_, y_ = st1.x_val, st1.y_val
        cum = st1.pix_cumsum
        if selection is not None:
            x_, y_ = x_[selection], y_[selection]
            cum = np.searchsorted(selection, st1.pix_cumsum)
        cnt = np.diff(cum)
        res = np.zeros((cfg.n_out, m, x_.shape[0]))
        ddx = x_[:, None] - self.yx_val[None, 1, 0, :]
        ddx /= cfg.dscale
        ddx += cfg.nc_ovl
        ddy = y_[:, None] - self.yx_val[None, 0, :, 0]
        ddy /= cfg.dscale
        ddy += cfg.nc_ovl
        for i_psf in range(cfg.n_out):
            for j_im in range(self.blk.n_inimage):
                if st1.pix_count[j_im] == 0:
                    continue
                out = np.zeros((int(cnt[j_im]), m))
                if cnt[j_im] > 0:
                    a, b = int(cum[j_im]), int(cum[j_im + 1])
                    R.gridD5512C(np.ascontiguousarray(np.pad(io[tab.grp_index(G, j_im), i_psf], 6)),
                                 np.ascontiguousarray(ddx[a:b] + 6), np.ascontiguousarray(ddy[a:b] + 6), out)
                    res[i_psf, :, a:b] = out.T
        return res

    def build_system_matrices(self):
        n = int(self.inpix_cumsum[-1])
        cs = [int(v) for v in self.inpix_cumsum]
        self.sysmata = np.zeros((n, n))
        for idx in range(9):
            sub = self._ii_block(self.instamps[idx], self.instamps[idx])
            s = self.selections[idx]
            if s is not None:
                sub = sub[np.ix_(s, s)]
            self.sysmata[cs[idx]:cs[idx + 1], cs[idx]:cs[idx + 1]] = sub
        for (a, b) in combinations(range(9), 2):
            sub = self._ii_block(self.instamps[a], self.instamps[b])
            sa, sb = self.selections[a], self.selections[b]
            if sa is not None:
                sub = sub[sa, :]
            if sb is not None:
                sub = sub[:, sb]
            self.sysmata[cs[a]:cs[a + 1], cs[b]:cs[b + 1]] = sub
            self.sysmata[cs[b]:cs[b + 1], cs[a]:cs[a + 1]] = sub.T
        self.mhalfb = np.zeros((self.blk.cfg.n_out, self.blk.cfg.n2f**2, n))
        for idx in range(9):
            self.mhalfb[:, :, cs[idx]:cs[idx + 1]] = self._io_block(self.instamps[idx], self.selections[idx])
        self.outovlc = self.tab.outovlc

    # ---- coadd.py:1104-1122 ----
    def post_kernel(self):
        cfg = self.blk.cfg
        if cfg.linear_algebra == "Iterative":
            self.UC = np.maximum(self.UC, 1e-32)
            self.Sigma = np.maximum(self.Sigma, 1e-32)
        if cfg.fade_kernel > 0:
            trapezoid(self.kappa, cfg.fade_kernel)
            trapezoid(self.Sigma, cfg.fade_kernel)
            trapezoid(self.UC, cfg.fade_kernel)

    # ---- coadd.py:1294-1363 ----
    def perform_coaddition(self):
        cfg = self.blk.cfg
        n_out, n2f, fk = cfg.n_out, cfg.n2f, cfg.fade_kernel
        n = int(self.inpix_cumsum[-1])
        if fk > 0:
            T_view = np.moveaxis(self.T, 1, -1).reshape((n_out, n, n2f, n2f))
            trapezoid(T_view, fk)
        Tsum_image = np.zeros(self.T.shape[:2] + (self.blk.n_inimage,))
        for j_st, (inst, sel) in enumerate(zip(self.instamps, self.selections)):
            cum = inst.pix_cumsum.astype(np.int64) if sel is None else np.searchsorted(sel, inst.pix_cumsum).astype(np.int64)
            cum = cum + int(self.inpix_cumsum[j_st])
            for k in range(self.blk.n_inimage):
                Tsum_image[:, :, k] += np.sum(self.T[:, :, cum[k]:cum[k + 1]], axis=2)
        self.Tsum_stamp = np.sum(Tsum_image, axis=1) / cfg.n2**2
        self.Tsum_inpix = np.sum(Tsum_image, axis=2).reshape((n_out, n2f, n2f))
        with np.errstate(invalid="ignore", divide="ignore"):
            Tsum_norm = Tsum_image / np.abs(Tsum_image).sum(axis=2)[:, :, None]
            self.Neff = 1.0 / np.sum(np.square(Tsum_norm), axis=2).reshape((n_out, n2f, n2f))
        if fk > 0:
            trapezoid(self.Neff, fk)
        self.outimage = np.einsum("oaj,ij->oia", self.T, self.indata).reshape((n_out, cfg.n_inframe, n2f, n2f))


def trapezoid(arr, fade_kernel, recover_mode=False):
    """coadd.py:1221-1292 (default pad_widths, all sides, truncated sinc)."""
    fk2 = fade_kernel * 2
    if not fk2 > 0:
        return
    ny, nx = arr.shape[-2:]
    s = np.arange(1, fk2 + 1, dtype=np.float64) / (fk2 + 1)
    s -= np.sin(2 * np.pi * s) / (2 * np.pi)
    sT = s[None, :].T
    if not recover_mode:
        arr[..., 0:fk2, :] *= sT
        arr[..., ny - 1:ny - 1 - fk2:-1, :] *= sT
        arr[..., :, 0:fk2] *= s
        arr[..., :, nx - 1:nx - 1 - fk2:-1] *= s
    else:
        arr[..., 0:fk2, :] /= sT
        arr[..., ny - 1:ny - 1 - fk2:-1, :] /= sT
        arr[..., :, 0:fk2] /= s
        arr[..., :, nx - 1:nx - 1 - fk2:-1] /= s
