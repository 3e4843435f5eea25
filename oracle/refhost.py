"""Import the *reference itself* from /root/reference (build container only; TEST INFRASTRUCTURE).

``/root/reference`` is read-only and does not exist on the GPU box, and several of the reference's
third-party imports (astropy, galsim, matplotlib, asdf, fitsio, ...; furry_parakeet too) are not
installed.  ``load()`` serves permissive dummy modules for the absent packages, gives
``astropy.units`` the three conversions the hot path uses, pre-registers a bare ``pyimcom`` package
(so pyimcom/__init__.py, which needs asdf, is not executed) and imports
``pyimcom.{config,routine,lakernel,psfutil,coadd}`` verbatim from the reference tree.
Nothing is copied.  Used only by tests/golden/make_golden.py and the container-only cross-checks.
"""

import contextlib
import importlib.abc
import importlib.machinery
import math
import os
import sys
import types

REF_SRC = "/root/reference/src/pyimcom"
MISSING = ("astropy", "galsim", "matplotlib", "asdf", "fitsio", "pytz", "gwcs", "healpy", "skimage", "treecorr",
           "piff", "memory_profiler", "furry_parakeet", "pyimcom_croutines")


def available() -> bool:
    return os.path.isdir(REF_SRC)


class _Meta(type):
    def __getattr__(cls, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _mk(k)


def _mk(k):
    return _Meta(k, (), {"__init__": lambda s, *a, **kw: None, "__call__": lambda s, *a, **kw: None})


class _Any(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        v = _mk(k)
        setattr(self, k, v)
        return v


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        root = name.split(".")[0]
        if root in ("furry_parakeet", "pyimcom_croutines"):
            return None  # genuinely absent: the reference falls back to .routine (lakernel.py:41-47)
        if root in MISSING:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Any(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


_loaded = None


def load():
    """Returns a namespace with the reference modules: coadd, psfutil, lakernel, routine, config."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree /root/reference not present (GPU box?)")
    sys.meta_path.insert(0, _Finder())
    import astropy.units as u  # noqa: E402  (stub)

    class _U:
        def __init__(self, rad):
            self.rad = rad

        def to(self, o):
            return self.rad / {"rad": 1.0, "degree": math.pi / 180, "arcmin": math.pi / 10800,
                               "arcsec": math.pi / 648000}[o]

    u.degree, u.arcmin, u.arcsec = _U(math.pi / 180), _U(math.pi / 10800), _U(math.pi / 648000)
    import matplotlib as mpl  # noqa: E402  (stub)

    mpl.rc_context = lambda *a, **k: contextlib.nullcontext()
    pkg = types.ModuleType("pyimcom")
    pkg.__path__ = [REF_SRC]
    pkg.__version__ = "ref"
    sys.modules["pyimcom"] = pkg
    import pyimcom.coadd as coadd  # noqa: E402
    import pyimcom.config as config  # noqa: E402
    import pyimcom.lakernel as lakernel  # noqa: E402
    import pyimcom.psfutil as psfutil  # noqa: E402
    import pyimcom.routine as routine  # noqa: E402

    _loaded = types.SimpleNamespace(coadd=coadd, psfutil=psfutil, lakernel=lakernel, routine=routine, config=config)
    return _loaded


def run_block(blk, kernel: str, kappaC, stamps=None, exact=None, only=False):
    """Drive the reference's own InStamp -> OutStamp path on a SynthBlock.

    Returns {(j,i): dict of the reference's arrays} for the requested OutStamps.  ``blk`` is a
    pyimcom_b200.synth.SynthBlock; its cfg is mutated to the requested kernel/kappa set.  ``only=True`` constructs and
    runs just the requested OutStamps (the reference counts of the two-pass protocol then cover that subset).
    """
    import numpy as np

    ref = load()
    cfg = blk.cfg
    cfg.linear_algebra = kernel
    cfg.kappaC_arr = np.asarray(kappaC, dtype=np.float64)
    psfutil, coadd = ref.psfutil, ref.coadd
    psfutil.PSFGrp.setup(npixpsf=cfg.npixpsf, oversamp=cfg.oversamp, dtheta=cfg.dtheta, psfsplit=cfg.psfsplit)
    psfutil.PSFOvl.setup(flat_penalty=cfg.flat_penalty)
    rb = types.SimpleNamespace(cfg=cfg, n_inimage=blk.n_inimage, timer=ref.config.Timer(), this_sub=0,
                               inimages=blk.inimages, outwcs=blk.outwcs, cache_dir=None)
    ns = cfg.n1P + 2
    rb.instamps = [[coadd.InStamp(rb, j, i) for i in range(ns)] for j in range(ns)]
    rb.outstamps = [[None] * ns for _ in range(ns)]
    if cfg.n_out > 1:  # reference bug psfutil.py:924 (self.n_out does not exist): patch the one attribute
        psfutil.PSFGrp.n_out = cfg.n_out
    rb.outpsfgrp = psfutil.PSFGrp(in_or_out=False, blk=rb)
    rb.outpsfovl = psfutil.PSFOvl(rb.outpsfgrp, None)
    rb.sysmata = psfutil.SysMatA(rb)
    rb.sysmatb = psfutil.SysMatB(rb)
    order = list(blk.stamp_order())
    if only:
        order = [ji for ji in order if ji in stamps]
    for (j, i) in order:
        rb.outstamps[j][i] = coadd.OutStamp(rb, j, i)
    rb.sysmata.iisubmats.clear()
    rb.sysmatb.iopsfovls.clear()
    out = {}
    f64 = {}
    lak = ref.lakernel
    # capture the f64 solution before the float32 cast (lakernel.py:304,358)
    orig_cho_solve = lak.cho_solve

    def spy(c_and_lower, b, **kw):
        r = orig_cho_solve(c_and_lower, b, **kw)
        f64.setdefault("Ti", []).append(r.T.copy())
        return r

    lak.cho_solve = spy
    try:
        for (j, i) in order:
            f64.clear()
            st = rb.outstamps[j][i]
            st(save_abc=True, save_t=True)
            if stamps is None or (j, i) in stamps:
                d = dict(sysmata=st.sysmata, mhalfb=st.mhalfb, outovlc=np.asarray(st.outovlc), T=st.T, UC=st.UC,
                         Sigma=st.Sigma, kappa=st.kappa, outimage=st.outimage, Tsum_stamp=st.Tsum_stamp,
                         Tsum_inpix=st.Tsum_inpix, Neff=st.Neff, inpix_cumsum=st.inpix_cumsum.copy())
                if "Ti" in f64:
                    d["Ti64"] = np.stack(f64["Ti"])
                out[(j, i)] = d
            rb.outstamps[j][i] = None
    finally:
        lak.cho_solve = orig_cho_solve
    return out
