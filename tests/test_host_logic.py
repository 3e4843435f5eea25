"""Host-side logic of the product that needs no GPU: the chunked OutStamp planning of GpuBlock (device uploads replaced by
host stand-ins), the reference-count indexing of the system-matrix seam, the sparse-grid pass of the partitioning."""

import numpy as np
import pytest

import cases

torch = pytest.importorskip("torch")


@pytest.fixture
def host_gpublock(monkeypatch):
    """pyimcom_b200.coadd with every device touch of prepare() replaced by a host stand-in."""
    from pyimcom_b200 import coadd

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(coadd, "h2d", lambda a: torch.from_numpy(np.ascontiguousarray(a)))
    monkeypatch.setattr(coadd._Arena, "upload", lambda self: (setattr(self, "h2d_bytes", 0), torch.zeros(1))[1])
    monkeypatch.setattr(coadd.GpuBlock, "reset_maps", lambda self: None)
    monkeypatch.setattr(coadd, "hbm_free_estimate", lambda: 100 << 30)
    return coadd


@pytest.mark.parametrize("a_cache", [True, False])
def test_chunked_planning_equals_upfront_planning(host_gpublock, a_cache):
    """prepare() plans the first chunk only; later chunks are planned on first use.  The metadata of every stamp must be
    what planning everything at once gives, no table may be registered after the arena upload, and every table a plan
    refers to must have an offset."""
    from oracle import routines as R
    from pyimcom_b200.psfovl_host import PSFTables

    coadd = host_gpublock
    blk = cases.make_block(cases.BLOCK_CASES["pad4"])
    tab = PSFTables(blk, R.iD5512C, R.gridD5512C, dedup=True)
    lazy = coadd.GpuBlock(blk, tab, a_cache=a_cache)
    lazy.plan_chunk = 3
    lazy.prepare()
    full = coadd.GpuBlock(blk, tab, a_cache=a_cache)
    full.plan_chunk = 1000
    full.prepare()
    nplanned = len(dict.keys(lazy.plans))
    assert len(dict.keys(full.plans)) == len(full.order) == 16
    assert nplanned == (3 if a_cache else 16)  # the fused-A route needs every plan before the tables are uploaded
    size0 = lazy.arena.size
    assert size0 == full.arena.size and len(lazy._pair_lut_list) == len(full._pair_lut_list)
    for k, ji in enumerate(lazy.order):
        ch, q = lazy._meta(k)
        chf, qf = full._meta(k)
        p, pf = lazy.plans[ji], full.plans[ji]
        assert p.n == pf.n and np.array_equal(p.idx, pf.idx) and np.array_equal(p.pcode, pf.pcode)
        o, of = int(ch["off_pix"][q]), int(chf["off_pix"][qf])
        for nm in ("idx", "pcode"):
            assert torch.equal(ch[nm][o:o + p.n], chf[nm][of:of + p.n]), nm
        a, b = int(ch["off_seg"][q]), int(ch["off_seg"][q + 1])
        af, bf = int(chf["off_seg"][qf]), int(chf["off_seg"][qf + 1])
        assert torch.equal(ch["seg_end"][a:b], chf["seg_end"][af:bf]) and torch.equal(ch["seg_img"][a:b], chf["seg_img"][af:bf])
        assert torch.equal(ch["lut_io"][q], chf["lut_io"][qf])
        if p.n:
            assert (p.lut_io[np.unique(p.pcode)] >= 0).all()
        if not a_cache:
            assert torch.equal(ch["lut"][q], chf["lut"][qf])
    assert lazy.arena.size == size0, "a table set was registered after the arena upload"
    assert len(dict.keys(lazy.plans)) == 16
    assert lazy.h2d_bytes == full.h2d_bytes
    if a_cache:
        assert lazy._pool_target() == full._pool_target() > 0
    with pytest.raises(KeyError):
        lazy.plans[(99, 99)]


def test_sysmat_reference_count_index():
    """SysMatA.iisubmat_dist against the diagram of psfutil.py:1875-1886 (13 'distances' of a later InStamp)."""
    from pyimcom_b200.sysmat import SysMatA

    d = SysMatA.iisubmat_dist
    assert d((2, 2), (2, 2)) == (2, 2, 0) and d((2, 2), (2, 3)) == (2, 2, 1) and d((2, 2), (2, 4)) == (2, 2, 2)
    assert [d((2, 2), (3, i))[2] for i in range(0, 5)] == [3, 4, 5, 6, 7]
    assert [d((2, 2), (4, i))[2] for i in range(0, 5)] == [8, 9, 10, 11, 12]
    assert d((2, 2), (5, 2)) is None and d((2, 2), (3, 5)) is None and d((2, 5), (3, 2)) is None
    with pytest.raises(AssertionError):
        d((2, 2), (2, 1))
    assert SysMatA.ji_st2psf((3, 5)) == (2, 4) and SysMatA.shift_ji_st((2, 4), (1, 0)) == (3, 4)


def test_cell_descriptor_layout():
    """The partition cell descriptors the host writes are the struct the kernels read (include/pyimcom_b200.h)."""
    import ctypes

    from pyimcom_b200 import _lib
    from pyimcom_b200.partition import CELL_DTYPE

    assert CELL_DTYPE.itemsize == ctypes.sizeof(_lib.PartCell) == 24
    for name, fld in zip(CELL_DTYPE.names, _lib.PartCell._fields_):
        assert name == fld[0] and CELL_DTYPE.fields[name][1] == getattr(_lib.PartCell, name).offset


def test_adapter_derives_the_reference_quantities():
    """pyimcom_b200.adapter: a reference-shaped Config (only the reference's own attribute names) gets the derived
    quantities the device path reads, equal to what StampConfig carries and -- in the build container -- to the
    reference's own PSFGrp.setup / PSFOvl.setup class state for configs/paper4_configs/H158_Chol_benchmark.json."""
    import types

    from pyimcom_b200.adapter import adapt_block, adapt_config, default_stamp_order
    from pyimcom_b200.synth import StampConfig

    sc = StampConfig(n1=6, n2=32, dtheta_arcsec=0.0390625, fade_kernel=3, postage_pad=1, npixpsf=48, oversamp=8,
                     instamp_pad_arcsec=1.24, psfsplit=True)
    ref_like = types.SimpleNamespace(n1P=sc.n1P, n2f=sc.n2f, dtheta=sc.dtheta, instamp_pad=sc.instamp_pad,
                                     npixpsf=sc.npixpsf, inpsf_oversamp=sc.oversamp, psfsplit=True, fade_kernel=3,
                                     postage_pad=1, NsideP=sc.NsideP, n_out=1, kappaC_arr=sc.kappaC_arr)
    v = adapt_config(ref_like)
    for k in ("oversamp", "nsamp", "nfft", "nsamp_ovl", "nc_ovl", "n2", "n1"):
        assert getattr(v, k) == getattr(sc, k), k
    for k in ("dscale", "rpix_search"):
        assert abs(getattr(v, k) - getattr(sc, k)) < 1e-12 * abs(getattr(sc, k)), k
    assert v.kappaC_arr is ref_like.kappaC_arr and adapt_config(sc) is sc  # pass-through, idempotent
    with pytest.raises(AttributeError):
        v.not_a_config_key
    blk = types.SimpleNamespace(cfg=ref_like, n_inimage=2)
    b = adapt_block(blk)
    assert list(b.stamp_order()) == list(default_stamp_order(sc.n1P)) and b.n_inimage == 2
    assert list(default_stamp_order(2)) == [(1, 1), (1, 2), (2, 1), (2, 2)]

    from oracle import refhost

    if not refhost.available():
        return
    import contextlib
    import io

    ref = refhost.load()
    with contextlib.redirect_stdout(io.StringIO()):
        c = ref.config.Config("/root/reference/configs/paper4_configs/H158_Chol_benchmark.json")
    v = adapt_config(c)
    G, O = ref.psfutil.PSFGrp, ref.psfutil.PSFOvl
    saved = {k: getattr(G, k, None) for k in ("oversamp", "nsamp", "nfft", "dscale", "psfsplit", "nc", "yxo")}
    try:
        G.setup(npixpsf=c.npixpsf, oversamp=c.inpsf_oversamp, dtheta=c.dtheta, psfsplit=bool(c.psfsplit))
        O.setup(flat_penalty=c.flat_penalty)
        assert (v.oversamp, v.nsamp, v.nfft, v.nsamp_ovl, v.nc_ovl) == (G.oversamp, G.nsamp, G.nfft, O.nsamp, O.nc)
        assert v.dscale == G.dscale and (v.n1, v.n2, v.n1P, v.n2f) == (80, 32, 84, 38)
        assert abs(v.rpix_search - 1.24 / 0.0390625) < 1e-9
    finally:
        for k, val in saved.items():
            if val is not None:
                setattr(G, k, val)
