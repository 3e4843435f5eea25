"""System-matrix seam: drop-ins for psfutil.SysMatA / psfutil.SysMatB of the reference.

``OutStamp.__init__`` and ``OutStamp._build_system_matrices`` talk to the system matrices through two objects hanging
off the block (coadd.py:863-867, 1032-1082):

    blk.sysmata.get_iisubmat(ji_st1, ji_st2, sim_mode=False, ji_st_out=None)  ->  (npix1, npix2) float64
    blk.sysmatb.get_iosubmat(ji_st_in, ji_st_out, sim_mode=False)              ->  (n_out, n2f^2, npix_selected) float64

with a two-pass protocol: every OutStamp is first constructed with ``sim_mode=True`` (references are only counted,
nothing is returned), the caches are cleared, and the real pass then asks in the same order (coadd.py:1620-1622,
2056-2069); a block is dropped when its count returns to zero (psfutil.py:2073-2076, 2181-2185).  The classes here keep
those names, arguments, return shapes and the reference counts; the arithmetic -- the D5512 interpolation of the
PSF-overlap tables -- runs on the device (``k_pair_blocks``, ``k_build_B_sep``) against tables and pixels uploaded once
per block, and each call downloads its result, so this seam gives the reference's OutStamp the device interpolator
without any other change.  (The block seam, ``coadd.GpuBlock``, keeps everything in HBM instead.)

There is no CPU fallback: without a CUDA device the constructors raise.
"""

from __future__ import annotations

import numpy as np

from .coadd import GpuBlock
from .psfovl_host import PSFTables, anchor


def _context(blk, tables=None) -> GpuBlock:
    """One device context per block, shared by its SysMatA and SysMatB: pixels and PSF-overlap tables in HBM."""
    ctx = getattr(blk, "_b200_sysmat_ctx", None)
    if ctx is None:
        if tables is None:
            from . import pyimcom_croutines as G

            tables = PSFTables(blk, G.iD5512C, G.gridD5512C)
        ctx = GpuBlock(blk, tables, kernel="Cholesky").prepare()
        blk._b200_sysmat_ctx = ctx
    return ctx


class SysMatA:
    """psfutil.SysMatA (psfutil.py:1764-2092): reference-counted InStamp-pair blocks of the system matrix A."""

    def __init__(self, blk, tables=None):
        self.blk = blk
        self._ctx = _context(blk, tables)  # (reference Blocks are wrapped by pyimcom_b200.adapter inside GpuBlock)
        self.iisubmats = {}  # psfutil.py:1797
        ns = blk.cfg.n1P + 2
        self.iisubmats_ref = np.zeros((ns, ns, 13), dtype=np.uint8)  # psfutil.py:1800

    ji_st2psf = staticmethod(anchor)  # psfutil.py:1803-1824

    @staticmethod
    def shift_ji_st(ji_st, dji_st):
        """psfutil.py:1826-1847."""
        return (ji_st[0] + dji_st[0], ji_st[1] + dji_st[1])

    @staticmethod
    def iisubmat_dist(ji_st1, ji_st2):
        """Index into iisubmats_ref, or None when the two InStamps are too far apart (psfutil.py:1849-1903)."""
        assert ji_st1 <= ji_st2, f"{ji_st1=} should precede {ji_st2=}"
        dj_st = ji_st2[0] - ji_st1[0]
        if not 0 <= dj_st <= 2:
            return None
        di_st = ji_st2[1] - ji_st1[1]
        if abs(di_st) > 2 or (dj_st == 0 and di_st < 0):
            return None
        return (*ji_st1, dj_st * 5 + di_st)

    def get_iisubmat(self, ji_st1, ji_st2, sim_mode: bool = False, ji_st_out=None, visualize: bool = False):
        """psfutil.py:2009-2092.  ji_st_out (the virtual-memory spill of the reference) is accepted and ignored: blocks
        that are still referenced stay in ``iisubmats``."""
        assert ji_st1 <= ji_st2, f"{ji_st1=} should precede {ji_st2=}"
        ji_dist = SysMatA.iisubmat_dist(ji_st1, ji_st2)
        assert ji_dist is not None, f"distance between InStamps {ji_st1} and {ji_st2} is out of range"
        key = (ji_st1, ji_st2)
        if sim_mode:
            self.iisubmats_ref[ji_dist] += 1
            self.iisubmats.setdefault(key, None)
            return None
        if self.iisubmats.get(key) is None:
            self.iisubmats[key] = self._ctx.pair_block(ji_st1, ji_st2).cpu().numpy()
        arr = self.iisubmats[key]
        if self.iisubmats_ref[ji_dist] > 0:  # (a caller that skipped the counting pass keeps everything cached)
            self.iisubmats_ref[ji_dist] -= 1
            if self.iisubmats_ref[ji_dist] == 0:
                del self.iisubmats[key]
        return arr

    def clear(self) -> None:
        """psfutil.py:2088-2092."""
        self.iisubmats.clear()
        del self.iisubmats_ref


class SysMatB:
    """psfutil.SysMatB (psfutil.py:2095-2199): one InStamp's slab of -B/2 for one OutStamp (the -2 is not included, as
    upstream); reference counts per 2x2 PSF group."""

    def __init__(self, blk, tables=None):
        self.blk = blk
        self._ctx = _context(blk, tables)
        self.iopsfovls = {}  # psfutil.py:2120
        self.iopsfovls_ref = np.zeros((blk.cfg.n1P // 2 + 1, blk.cfg.n1P // 2 + 1), dtype=np.uint8)

    def get_iosubmat(self, ji_st_in, ji_st_out, sim_mode: bool = False, visualize: bool = False):
        """psfutil.py:2125-2185; the OutStamp is blk.outstamps[j][i] (its ``selections`` and ``yx_val`` are read, as
        PSFOvl._call_io_cross does, psfutil.py:1526-1541)."""
        assert max(abs(ji_st_in[0] - ji_st_out[0]), abs(ji_st_in[1] - ji_st_out[1])) <= 1, (
            f"distance between InStamp {ji_st_in} and OutStamp {ji_st_out} is out of range")
        inpsf_key = tuple(ji >> 1 for ji in anchor(ji_st_in))
        if sim_mode:
            self.iopsfovls_ref[inpsf_key] += 1
            self.iopsfovls.setdefault(inpsf_key, None)
            return None
        self.iopsfovls.setdefault(inpsf_key, None)
        if self.iopsfovls_ref[inpsf_key] > 0:
            self.iopsfovls_ref[inpsf_key] -= 1
        st2 = self.blk.outstamps[ji_st_out[0]][ji_st_out[1]]
        selection = st2.selections[(ji_st_in[0] - ji_st_out[0] + 1) * 3 + (ji_st_in[1] - ji_st_out[1] + 1)]
        x0out, y0out = float(st2.yx_val[1, 0, 0]), float(st2.yx_val[0, 0, 0])
        res = self._ctx.io_block(ji_st_in, selection, x0out, y0out).cpu().numpy()
        if self.iopsfovls_ref[inpsf_key] == 0:
            del self.iopsfovls[inpsf_key]
        return res

    def clear(self) -> None:
        """psfutil.py:2187-2199."""
        self.iopsfovls.clear()
        del self.iopsfovls_ref
