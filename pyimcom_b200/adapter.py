"""Adapter between the reference's own ``Config`` / ``Block`` objects and the attributes the device path reads.

``pyimcom_b200.coadd.GpuBlock`` and the ``sysmat`` drop-ins were written against the synthetic harness
(``synth.StampConfig`` / ``SynthBlock``), which carries a few DERIVED quantities as plain attributes.  The reference
keeps the same quantities elsewhere: ``PSFGrp.setup`` class state (psfutil.py:568-613: ``oversamp, nsamp, nfft,
dscale``), ``PSFOvl.setup`` (psfutil.py:1065-1089: ``nsamp, nc`` of the overlap tables) and an expression inside
``OutStamp._process_input_stamps`` (coadd.py:923: the search radius in output pixels).  ``adapt_config`` wraps a reference
``Config`` (or anything with its attribute names) in a view that answers both vocabularies; ``adapt_block`` does the
same for a ``Block`` (``stamp_order`` = the 2x2-group traversal of ``coadd_output_stamps``, coadd.py:2056-2060).
Nothing is copied or modified: unknown attributes fall through to the wrapped object.
"""

from __future__ import annotations

import math

ARCSEC = math.pi / 648000.0  # config.py:87
PIXSCALE_NATIVE_ARCSEC = 0.11  # config.py:97


class ConfigView:
    def __init__(self, cfg):
        object.__setattr__(self, "_cfg", cfg)

    def __getattr__(self, k):
        cfg = object.__getattribute__(self, "_cfg")
        if hasattr(cfg, k):
            return getattr(cfg, k)
        derive = _DERIVED.get(k)
        if derive is None:
            raise AttributeError(f"{type(cfg).__name__} has no attribute {k!r} and the adapter cannot derive it")
        return derive(self)

    def __setattr__(self, k, v):
        setattr(object.__getattribute__(self, "_cfg"), k, v)


def _oversamp(c):
    return int(c.inpsf_oversamp)  # coadd.py:1603: PSFGrp.setup(oversamp=cfg.inpsf_oversamp)


_DERIVED = {
    "oversamp": _oversamp,
    "dtheta_arcsec": lambda c: c.dtheta * 3600.0,  # config.py:502: dtheta is stored in degrees
    "nsamp": lambda c: c.npixpsf * c.oversamp - 1,  # psfutil.py:593
    "nfft": lambda c: c.npixpsf * c.oversamp * 2,  # psfutil.py:607
    "dscale": lambda c: PIXSCALE_NATIVE_ARCSEC / c.oversamp / (c.dtheta * 3600.0),  # psfutil.py:610
    "nsamp_ovl": lambda c: 2 * c.nsamp + 1 if c.psfsplit else c.nsamp,  # psfutil.py:1088
    "nc_ovl": lambda c: c.nsamp_ovl // 2,  # psfutil.py:1089
    "rpix_search": lambda c: (c.instamp_pad / ARCSEC) / (c.dtheta * 3600.0),  # coadd.py:923
    "n2": lambda c: c.n2f - 2 * c.fade_kernel,  # config.py: n2f = n2 + 2 fade_kernel
    "n1": lambda c: c.n1P - 2 * c.postage_pad,
    "psf_circ": lambda c: False,
    "psf_norm": lambda c: False,
    "amp_penalty": lambda c: (0.0, 0.0),
    "flat_penalty": lambda c: 0.0,
    "psfsplit": lambda c: False,
    "outmaps": lambda c: "USKTN",
    "no_qlt_ctrl": lambda c: False,
    "sigmatarget_extra": lambda c: (),
    "outpsf_extra": lambda c: (),
}


def adapt_config(cfg):
    """A view of `cfg` with every attribute the device path reads (see module docstring); idempotent."""
    return cfg if isinstance(cfg, ConfigView) or all(hasattr(cfg, k) for k in ("dscale", "nc_ovl", "rpix_search")) \
        else ConfigView(cfg)


class BlockView:
    def __init__(self, blk):
        object.__setattr__(self, "_blk", blk)
        object.__setattr__(self, "cfg", adapt_config(blk.cfg))

    def __getattr__(self, k):
        return getattr(object.__getattribute__(self, "_blk"), k)

    def __setattr__(self, k, v):
        setattr(object.__getattribute__(self, "_blk"), k, v)

    def stamp_order(self):
        blk = object.__getattribute__(self, "_blk")
        if hasattr(blk, "stamp_order"):
            return blk.stamp_order()
        return default_stamp_order(self.cfg.n1P)


def default_stamp_order(n1P):
    """OutStamp traversal of Block.coadd_output_stamps: 2x2 groups in raster order (coadd.py:2056-2060)."""
    for j in range(1, n1P + 1, 2):
        for i in range(1, n1P + 1, 2):
            for dj in range(2):
                for di in range(2):
                    if j + dj <= n1P and i + di <= n1P:
                        yield (j + dj, i + di)


def adapt_block(blk):
    """A view of a reference Block (or a SynthBlock) whose cfg answers the derived attributes and which has
    stamp_order(); idempotent."""
    if isinstance(blk, BlockView):
        return blk
    if hasattr(blk, "stamp_order") and adapt_config(blk.cfg) is blk.cfg:
        return blk
    return BlockView(blk)
