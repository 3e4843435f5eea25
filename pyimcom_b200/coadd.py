"""Device-resident block driver: the OutStamp loop of one mosaic block on a B200.

Mirrors, for the hot path only, ``coadd.Block.coadd_output_stamps`` / ``_output_stamp_wrapper``
(coadd.py:1939-2084) and ``OutStamp.__call__`` (coadd.py:979-1363):

    per OutStamp:  _process_input_stamps  -> gather of the selected input pixels      (coadd.py:886-977)
                   _build_system_matrices -> fused A / mBhalf assembly kernels        (coadd.py:1002-1085)
                   LAKERNEL[...]          -> pyimcom_b200.lakernel solvers            (lakernel.py)
                   trapezoid, _perform_coaddition -> finalize / stamp_maps kernels    (coadd.py:1221-1363)
                   overlap-add into the block maps                                    (coadd.py:1976-1994)

Inputs are a block object exposing what the reference's ``Block`` exposes to ``OutStamp`` (``cfg``,
``n_inimage``, ``instamps[j][i]`` with ``x_val, y_val, data, pix_count, pix_cumsum, make_selection``) and a
``PSFTables`` object holding the PSF-overlap tables (psfovl_host.py).  Everything between the upload of
those inputs and the download of the block maps stays in HBM.

The reference's SysMatA/SysMatB reference-counted caches (psfutil.py:1764-2199) exist to bound CPU RAM;
here every stamp's A and mBhalf are assembled directly from positions and tables by one kernel each.
"""

from __future__ import annotations

import ctypes as C
import os
import types
import weakref

import numpy as np
import torch

from . import _lib
from .adapter import adapt_block
from .lakernel import SOLVERS, ApplySpec, DeviceSystem, apply_T, eigen_decompose_batch, ptr, rup, solve_chol_batch, solve_eigen
from .lakernel import solve_chol_finish, solve_chol_launch, side_streams_in_use
from .lakernel import stream_handle
from .lakernel import trapezoid_weights
from .psfovl_host import anchor

TABLEREF_DTYPE = np.dtype([("offset", "<i8"), ("flip", "<i4"), ("pad_", "<i4"), ("penalty_sub", "<f8")])
assert TABLEREF_DTYPE.itemsize == C.sizeof(_lib.TableRef)
PAIRDESC_DTYPE = np.dtype([("offA", "<i4"), ("nA", "<i4"), ("offB", "<i4"), ("nB", "<i4"), ("out", "<i8"), ("ld", "<i4"),
                           ("lut", "<i4"), ("same", "<i4"), ("pad_", "<i4")])
assert PAIRDESC_DTYPE.itemsize == C.sizeof(_lib.PairDesc)


_PINNED = {}  # (address, bytes) of a large host array -> (the array, its page-locked copy)


def h2d(a: np.ndarray, keep_pinned: bool = False) -> torch.Tensor:
    """Host array -> device tensor through pinned memory, asynchronous on the current stream.

    keep_pinned: the page-locked staging copy of a LARGE array is kept (keyed by the array's buffer, which is kept alive)
    so that uploading the same host array again -- the PSF-overlap table sets, which every block of a mosaic built from
    the same PSFTables object sends again -- is one DMA from pinned memory without the pageable-to-pinned memcpy
    (81 MB of tables: ~7 ms of host time per block).  The caller must not modify such an array afterwards."""
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a)
    if t.numel() == 0:
        return t.cuda()
    if keep_pinned and a.nbytes >= (1 << 20):
        key = (a.__array_interface__["data"][0], a.nbytes, str(a.dtype))
        hit = _PINNED.get(key)
        if hit is None:
            if len(_PINNED) >= 64:
                _PINNED.clear()
            hit = _PINNED[key] = (a, t.pin_memory())
        return hit[1].to("cuda", non_blocking=True)
    return t.pin_memory().to("cuda", non_blocking=True)


POLY_PERIODS = (2, 3, 4, 5, 6, 8, 10, 12, 16)  # instantiations of the polyphase A-assembly kernel


class _Arena:
    """PSF-overlap tables in HBM: each table set is stored once, zero-padded by 6 (np.pad(ovl, 6), psfutil.py:1471).

    In-out tables (read by the mBhalf kernel along the dense output grid) stay row-major.  In-in tables (read by
    the A kernel at positions one native pixel = ``poly`` table samples apart) are stored polyphase,
    ``T'[y % P][x % P][y // P][x // P]``, so that a warp's 32 neighbouring input pixels read neighbouring doubles
    (include/pyimcom_b200.h, b200_dev_build_A)."""

    def __init__(self, nsamp_ovl, poly=0):
        self.ns = nsamp_ovl
        self.ngrid = nsamp_ovl + 12
        self.poly = int(poly) if int(poly) in POLY_PERIODS else 0
        self.ncell = -(-self.ngrid // self.poly) if self.poly else 0
        self.sets = []  # (array, lead, poly period, base offset)
        self.base = {}
        self.stride = {}
        self.size = 0

    def register(self, arr, poly=False):
        """(base offset, stride between consecutive tables) of the table set arr inside the arena, in doubles."""
        self.offset(arr, (0,) * (arr.ndim - 2), poly=poly)
        key = (id(arr), self.poly if poly else 0)
        return self.base[key], self.stride[key]

    def offset(self, arr, idx, poly=False):
        """Offset (in doubles) of table arr[idx] inside the arena; registers arr on first use."""
        P = self.poly if poly else 0
        key = (id(arr), P)
        if key not in self.base:
            assert arr.shape[-1] == self.ns and arr.shape[-2] == self.ns
            lead = int(np.prod(arr.shape[:-2]))
            self.base[key] = self.size
            self.stride[key] = P * P * self.ncell * self.ncell if P else self.ngrid * self.ngrid
            self.sets.append((arr, lead, P, self.size))  # keeps arr alive: ids stay unique
            self.size += lead * self.stride[key]
        flat = int(np.ravel_multi_index(idx, arr.shape[:-2]))
        return self.base[key] + flat * self.stride[key]

    def upload(self):
        """Raw tables go up as they are (one pinned copy per table set); padding and re-layout happen on the device."""
        # (+2: the strip stage of k_build_B_sep reads aligned 16-byte pairs and may touch the double after a window)
        dst = torch.empty(max(1, self.size) + 2, dtype=torch.float64, device="cuda")
        self.h2d_bytes = 0
        st = stream_handle()
        for arr, lead, P, base in self.sets:
            if torch.is_tensor(arr):  # built on the device (psfovl_device.DeviceTables): nothing crosses PCIe
                src = arr.reshape(lead, self.ns, self.ns).contiguous()
            else:
                src = h2d(np.asarray(arr, dtype=np.float64).reshape(lead, self.ns, self.ns), keep_pinned=True)
                self.h2d_bytes += src.numel() * 8
            _lib.dev_layout_tables(ptr(src), lead, self.ns, 6, self.ngrid, P, C.c_void_p(dst.data_ptr() + 8 * base), st)
        return dst


_TOTAL_HBM = {}
_MESH = {}  # (na, nb) -> index grids of an (na x nb) image-pair block
# Pair-block pools outlive the GpuBlock that used them: a fresh block takes a parked pool instead of asking the
# allocator for another multi-GB segment (a cudaMalloc of that size was measured to stall a step by > 100 ms).
_PARKED_POOLS = {}


def hbm_free_estimate() -> int:
    """Bytes this process can still place on the current device, from torch's own counters.  cudaMemGetInfo is not
    used on the hot path: it takes a driver-wide lock and was measured to stall for tens of ms while work is queued."""
    dev = torch.cuda.current_device()
    if dev not in _TOTAL_HBM:
        _TOTAL_HBM[dev] = torch.cuda.get_device_properties(dev).total_memory
    return max(0, int(0.92 * _TOTAL_HBM[dev]) - torch.cuda.memory_allocated(dev))


def release_parked_pools() -> None:
    """Give the parked pair-block pools (at most two per device) back to the allocator."""
    for parked in _PARKED_POOLS.values():
        parked.clear()


class StampPlan:
    """Host-side description of one OutStamp (what OutStamp.__init__ / _process_input_stamps derive)."""

    __slots__ = ("j_st", "i_st", "n", "idx", "pcode", "groups", "seg_end", "seg_img", "inpix_cumsum", "x0out", "y0out",
                 "lut", "lut_io", "insts")


class _Plans(dict):
    """StampPlans by (j_st, i_st); a stamp that has not been planned yet is planned (with its chunk) on first access."""

    def __init__(self, gblk):
        super().__init__()
        self._gblk = weakref.ref(gblk)  # no reference cycle: a dropped GpuBlock frees its HBM at once

    def __missing__(self, ji):
        g = self._gblk()
        if g is None or ji not in g._pos:
            raise KeyError(ji)
        g._ensure_planned(g._pos[ji])
        return dict.__getitem__(self, ji)


class GpuBlock:
    """One mosaic block on one GPU."""

    def __init__(self, blk, tables, kernel: str | None = None, a_cache: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("pyimcom_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        blk = adapt_block(blk)  # a reference Block / Config gets the derived attributes (dscale, nc_ovl, rpix_search, ...)
        self.blk, self.tab = blk, tables
        self.cfg = blk.cfg
        self.kernel = kernel or self.cfg.linear_algebra
        if self.kernel not in SOLVERS:
            raise ValueError(f"unsupported LAKERNEL {self.kernel!r} (Cholesky, Eigen, Iterative, Empirical)")
        self.plans = _Plans(self)
        self.order = []
        self._pos = {}
        self._uploaded = False
        # period of the polyphase in-in tables = native pixel pitch in table samples = oversamp (psfutil.py:610)
        self.poly = int(getattr(self.cfg, "oversamp", 0))
        # A from cached InStamp-pair blocks (the device form of SysMatA's cache) or, if False, one fused kernel per
        # OutStamp that interpolates all of its entries; both give bit-identical matrices
        self.a_cache = bool(a_cache)
        self.pool_bytes = None  # pair-block pool size; default: a quarter of the free HBM, at most 32 GB

    # ---------------------------------------------------------------------------------------------
    # host-side planning (coadd.py:846-977)
    # ---------------------------------------------------------------------------------------------
    def _global_pixels(self):
        blk, cfg = self.blk, self.cfg
        ns = cfg.n1P + 2
        dp = getattr(blk, "device_pixels", None)
        if dp is not None:  # InStamps binned on the device (partition.PartitionedBlock): the pixels are already in HBM
            self.inst_off = np.asarray(dp["inst_off"], dtype=np.int64)
            self.h_x, self.h_y, self.h_img, self.h_data = dp["h_x"], dp["h_y"], dp["h_img"], None
            self.npix_total = int(dp["npix_total"])
            return
        self.inst_off = np.zeros((ns, ns), dtype=np.int64)
        xs, ys, imgs, datas = [], [], [], []
        off = 0
        for j in range(ns):
            for i in range(ns):
                st = blk.instamps[j][i]
                self.inst_off[j, i] = off
                npix = int(st.pix_cumsum[-1])
                xs.append(st.x_val)
                ys.append(st.y_val)
                imgs.append(np.repeat(np.arange(blk.n_inimage, dtype=np.int32), st.pix_count.astype(np.int64)))
                datas.append(st.data)
                off += npix
        self.h_x = np.concatenate(xs) if xs else np.zeros(0)
        self.h_y = np.concatenate(ys) if ys else np.zeros(0)
        self.h_img = np.concatenate(imgs).astype(np.int32) if imgs else np.zeros(0, dtype=np.int32)
        self.h_data = np.ascontiguousarray(np.concatenate(datas, axis=1), dtype=np.float32)
        self.npix_total = off

    def plan_stamp(self, j_st, i_st) -> StampPlan:
        blk, cfg, tab = self.blk, self.cfg, self.tab
        nimg = blk.n_inimage
        p = StampPlan()
        p.j_st, p.i_st = j_st, i_st
        bottom, left = (j_st - 1) * cfg.n2, (i_st - 1) * cfg.n2
        top, right = bottom + cfg.n2 - 1, left + cfg.n2 - 1
        fk = cfg.fade_kernel
        p.x0out, p.y0out = float(left - fk), float(bottom - fk)
        r = cfg.rpix_search
        idx, pcode, seg_end, seg_img, counts = [], [], [], [], []
        groups = []
        pos = 0
        p.insts = [(j_st + dj, i_st + di) for dj in (-1, 0, 1) for di in (-1, 0, 1)]
        for dj in (-1, 0, 1):
            for di in (-1, 0, 1):
                jj, ii = j_st + dj, i_st + di
                st = blk.instamps[jj][ii]
                xp = [left - 0.5, None, right + 0.5][di + 1]  # coadd.py:929-931
                yp = [bottom - 0.5, None, top + 0.5][dj + 1]
                sel = st.make_selection((xp, yp), r)
                G = anchor((jj, ii))
                if G not in groups:
                    groups.append(G)
                lg = groups.index(G)
                base = int(self.inst_off[jj, ii])
                if sel is None:
                    loc = np.arange(int(st.pix_cumsum[-1]), dtype=np.int64)
                    cum = st.pix_cumsum.astype(np.int64)
                else:
                    loc = sel.astype(np.int64)
                    cum = np.searchsorted(sel, st.pix_cumsum).astype(np.int64)
                idx.append(base + loc)
                pcode.append(lg * nimg + self.h_img[base + loc])
                for k in range(nimg):  # (instamp, image) segments for Tsum_image (coadd.py:1327-1337)
                    if cum[k + 1] > cum[k]:
                        seg_end.append(pos + int(cum[k + 1]))
                        seg_img.append(k)
                counts.append(loc.size)
                pos += loc.size
        p.n = pos
        p.idx = np.concatenate(idx).astype(np.int32)
        p.pcode = np.concatenate(pcode).astype(np.int32)
        p.groups = groups
        p.seg_end = np.asarray(seg_end, dtype=np.int32)
        p.seg_img = np.asarray(seg_img, dtype=np.int32)
        p.inpix_cumsum = np.cumsum([0] + counts, dtype=np.uint32)
        # table look-ups for the (<= 4 groups) x (n_inimage) codes of this stamp
        ncode = 4 * nimg
        lut = np.zeros((ncode, ncode), dtype=TABLEREF_DTYPE)
        lut["offset"] = -1
        lut_io = np.full((ncode, cfg.n_out), -1, dtype=np.int64)
        present = np.unique(p.pcode)
        pair, io_c = self._pair_cache, self._io_cache  # (group, image) look-ups are shared by neighbouring stamps
        gk = [(groups[c // nimg], int(c % nimg)) for c in present]
        rows = []
        for ca, (Ga, ka) in zip(present, gk):
            if (Ga, ka) not in io_c:
                tab.group(Ga)
                io = tab.get_io(Ga)  # registered already when the pair-block cache is on (_register_tables)
                io_c[(Ga, ka)] = [self.arena.offset(io, (tab.grp_index(Ga, ka), o)) for o in range(cfg.n_out)]
            lut_io[ca, :] = io_c[(Ga, ka)]
            if self.a_cache:  # the per-stamp in-in look-up table is only read by the fused kernel
                continue
            for (Gb, kb) in gk:
                key = (Ga, ka, Gb, kb)
                if key not in pair:
                    tab.group(Gb)
                    t, tidx, flip = tab.table_ii_ref(Ga, ka, Gb, kb)
                    n_in = len(tab.grp_imgs[Ga]) if Ga == Gb else (len(tab.grp_imgs[Ga]) * len(tab.grp_imgs[Gb])) ** 0.5
                    pair[key] = (self.arena.offset(t, tidx, poly=True), int(flip), 0, cfg.flat_penalty / n_in)
                rows.append(pair[key])
        if rows:
            lut[np.ix_(present, present)] = np.array(rows, dtype=TABLEREF_DTYPE).reshape(len(present), len(present))
        p.lut, p.lut_io = lut, lut_io
        return p

    def _pair_lut(self, Ga, Gb) -> int:
        """Index of the (nimg x nimg) table-reference block for pixels of group Ga (rows) against group Gb (columns).

        Same rules as PSFTables.table_ii_ref (same group: triangle entry, flipped when a > b, psfutil.py:1652-1665;
        different groups: dense table of the ordered pair, the reversed order being the flipped table), evaluated for
        the whole image x image block at once."""
        key = (Ga, Gb)
        if key not in self._pair_lut_idx:
            tab, cfg, nimg = self.tab, self.cfg, self.blk.n_inimage
            arr = np.zeros((nimg, nimg), dtype=TABLEREF_DTYPE)
            arr["offset"] = -1
            tab.group(Ga)
            tab.group(Gb)
            ia, ib = np.asarray(tab.grp_imgs[Ga], dtype=np.int64), np.asarray(tab.grp_imgs[Gb], dtype=np.int64)
            na, nb = ia.size, ib.size
            if na and nb:
                if (na, nb) not in _MESH:
                    _MESH[(na, nb)] = np.meshgrid(np.arange(na), np.arange(nb), indexing="ij")
                qa, qb = _MESH[(na, nb)]  # positions inside the groups
                if Ga == Gb:
                    base, stride = self.arena.register(tab.get_self(Ga), poly=True)
                    lo, hi = np.minimum(qa, qb), np.maximum(qa, qb)
                    flat = (2 * na - lo + 1) * lo // 2 + hi - lo  # tri_index(n_psf, lo, hi), psfutil.py:1139-1175
                    flip = qa > qb
                    n_in = float(na)
                elif Ga < Gb:
                    base, stride = self.arena.register(tab.get_cross(Ga, Gb), poly=True)
                    flat, flip = qa * nb + qb, np.zeros_like(qa, dtype=bool)
                    n_in = (na * nb) ** 0.5
                else:
                    base, stride = self.arena.register(tab.get_cross(Gb, Ga), poly=True)
                    flat, flip = qb * na + qa, np.ones_like(qa, dtype=bool)
                    n_in = (na * nb) ** 0.5
                full = na == nimg and nb == nimg  # every image covers both groups: fill the block in place
                sub = arr if full else np.zeros((na, nb), dtype=TABLEREF_DTYPE)
                sub["offset"] = base + flat * stride
                sub["flip"] = flip
                sub["penalty_sub"] = cfg.flat_penalty / n_in
                if not full:
                    arr[np.ix_(ia, ib)] = sub
            self._pair_lut_idx[key] = len(self._pair_lut_list)
            self._pair_lut_list.append(arr)
        return self._pair_lut_idx[key]

    def prepare(self, stamps=None):
        """Register the tables of the requested OutStamps (default: the whole block in the reference's 2x2-group
        order), upload pixels and tables, and plan the first chunk of stamps.  Later chunks are planned when they are
        first needed; the pipelined run asks for the next chunk while the device factorises the current one."""
        import time

        t0 = time.perf_counter()
        self._global_pixels()
        self.arena = _Arena(self.cfg.nsamp_ovl, poly=self.poly)
        self._pair_cache, self._io_cache = {}, {}
        self._pair_lut_idx, self._pair_lut_list = {}, []
        self.order = list(stamps) if stamps is not None else list(self.blk.stamp_order())
        self._pos = {ji: k for k, ji in enumerate(self.order)}
        self.plans = _Plans(self)
        self._chunks = {}
        if self.a_cache:
            self._register_tables()
        else:  # the fused A kernel's per-stamp look-up tables register their table sets while the stamps are planned
            for ji in self.order:
                dict.__setitem__(self.plans, ji, self.plan_stamp(*ji))
        t1 = time.perf_counter()
        self.upload()
        if self.order:
            self._ensure_planned(0)
        self.host_seconds = {"plan": t1 - t0, "upload_enqueue": time.perf_counter() - t1}
        return self

    def _register_tables(self):
        """Every table set the planned OutStamps can touch (in-out tables per 2x2 group, in-in tables per ordered
        group pair of one 3x3 neighbourhood) gets its place in the arena before the arena is uploaded."""
        tab, seen = self.tab, set()
        for (j_st, i_st) in self.order:
            groups = []
            for dj in (-1, 0, 1):
                for di in (-1, 0, 1):
                    G = anchor((j_st + dj, i_st + di))
                    if G not in groups:
                        groups.append(G)
            for G in groups:
                if G not in seen:
                    seen.add(G)
                    tab.group(G)
                    self.arena.register(tab.get_io(G))
            for Ga in groups:
                for Gb in groups:
                    self._pair_lut(Ga, Gb)

    def upload(self):
        cfg = self.cfg
        dp = getattr(self.blk, "device_pixels", None)
        if dp is not None:
            self.d_x, self.d_y, self.d_data, self.d_img = dp["d_x"], dp["d_y"], dp["d_data"], dp["d_img"]
            self.d_tables = self.arena.upload()
        else:
            self.d_x = h2d(self.h_x)
            self.d_y = h2d(self.h_y)
            self.d_data = h2d(self.h_data)
            self.d_tables = self.arena.upload()
            self.d_img = h2d(self.h_img)
        plut = np.stack(self._pair_lut_list) if self._pair_lut_list else np.zeros((1, 1, 1), dtype=TABLEREF_DTYPE)
        self.d_pair_lut = h2d(plut.view(np.uint8).reshape(-1))
        self._pairs = {}  # (inst_a, inst_b) -> (pool offset in doubles, ld)
        self._pool = None
        self._pool_used = 0
        self._pool_gen = 0  # bumped whenever cached blocks are dropped: older assemble closures refuse to run
        self.pool_evictions = 0
        self.pair_points = 0  # entries interpolated so far (vs sum of n^2/2 without the cache)
        self.d_fade_w = h2d(trapezoid_weights(cfg.fade_kernel)) if cfg.fade_kernel > 0 else None
        crossed = (self.d_pair_lut,) if dp is not None else (self.d_x, self.d_y, self.d_data, self.d_img, self.d_pair_lut)
        self.h2d_bytes = self.arena.h2d_bytes + sum(t.numel() * t.element_size() for t in crossed)
        self.reset_maps()
        self._uploaded = True

    plan_chunk = 16  # OutStamps planned (and their metadata uploaded) together

    def _ensure_planned(self, k: int):
        """Plan the chunk of stamps that holds position k of self.order and upload its metadata (pixel indices, table
        codes, segment ends, in-out look-ups), packed into a few arrays with one H2D copy each."""
        c = k // self.plan_chunk
        if c in self._chunks:
            return self._chunks[c]
        ks = range(c * self.plan_chunk, min((c + 1) * self.plan_chunk, len(self.order)))
        plans = []
        for q in ks:
            ji = self.order[q]
            if not dict.__contains__(self.plans, ji):
                dict.__setitem__(self.plans, ji, self.plan_stamp(*ji))
            plans.append(dict.__getitem__(self.plans, ji))
        cat = lambda arrs, dt: np.concatenate(arrs).astype(dt) if arrs else np.zeros(0, dtype=dt)  # noqa: E731
        ch = {
            "off_pix": np.concatenate([[0], np.cumsum([p.n for p in plans])]).astype(np.int64),
            "off_seg": np.concatenate([[0], np.cumsum([p.seg_end.size for p in plans])]).astype(np.int64),
            "idx": h2d(cat([p.idx for p in plans], np.int32)),
            "pcode": h2d(cat([p.pcode for p in plans], np.int32)),
            "seg_end": h2d(cat([p.seg_end for p in plans], np.int32)),
            "seg_img": h2d(cat([p.seg_img for p in plans], np.int32)),
            "lut_io": h2d(np.ascontiguousarray(np.stack([p.lut_io for p in plans]))),
        }
        dev = ["idx", "pcode", "seg_end", "seg_img", "lut_io"]
        if not self.a_cache:  # the per-stamp in-in look-up table is only read by the fused kernel
            ch["lut"] = h2d(np.stack([p.lut for p in plans]).view(np.uint8).reshape(len(plans), -1))
            dev.append("lut")
        self.h2d_bytes += sum(ch[nm].numel() * ch[nm].element_size() for nm in dev)
        self._chunks[c] = ch
        return ch

    def _meta(self, k: int):
        """(chunk, position inside the chunk) of stamp k of self.order."""
        return self._ensure_planned(k), k % self.plan_chunk

    def reset_cache(self):
        """Forget every cached InStamp-pair block (a new mosaic block starts with an empty SysMatA cache)."""
        self._pairs.clear()
        self._pool_used = 0
        self._pool_gen += 1
        self.pair_points = 0

    def reset_maps(self):
        """Zero-initialised block maps (coadd.py:2028-2047)."""
        cfg, dev = self.cfg, "cuda"
        side = cfg.NsideP + 2 * cfg.fade_kernel
        self.side = side
        n_out = cfg.n_out
        self.out_map = torch.zeros((n_out, cfg.n_inframe, side, side), dtype=torch.float32, device=dev)
        self.T_weightmap = torch.zeros((n_out, self.blk.n_inimage, cfg.n1P, cfg.n1P), dtype=torch.float32, device=dev)
        self.UC_map = torch.zeros((n_out, side, side), dtype=torch.float32, device=dev)
        self.Sigma_map = torch.zeros_like(self.UC_map)
        self.kappa_map = torch.zeros_like(self.UC_map)
        self.Tsum_map = torch.zeros_like(self.UC_map)
        self.Neff_map = torch.zeros_like(self.UC_map)

    # ---------------------------------------------------------------------------------------------
    # device pipeline of one OutStamp
    # ---------------------------------------------------------------------------------------------
    # ---------------------------------------------------------------------------------------------
    # InStamp-pair block cache (SysMatA, psfutil.py:1764-2092)
    # ---------------------------------------------------------------------------------------------
    def _inst_count(self, ji):
        return int(self.blk.instamps[ji[0]][ji[1]].pix_cumsum[-1])

    def ensure_pairs(self, plans, before_evict=None):
        """Interpolate, in one launch, every InStamp-pair block the given OutStamps need that is not cached yet.

        Blocks live in one pool (bump allocation).  When the pool is full the cache is dropped and the blocks of the
        current OutStamps are recomputed: the traversal order makes older blocks dead anyway, as the reference's
        reference counts do (psfutil.py:1997-2004).  ``before_evict`` is called first in that case: a pipelined batch
        whose systems are still unmaterialised views of the pool (DeviceSystem.assemble; the repair branch of the
        Cholesky kernel re-assembles A from them) has to be finished before its blocks are overwritten.  Every
        eviction bumps ``_pool_gen``; an assemble closure of an older generation refuses to run."""
        def wanted():
            seen, out = set(), []
            for p in plans:
                for a in range(len(p.insts)):
                    for b in range(a, len(p.insts)):
                        key = (p.insts[a], p.insts[b])
                        if key in self._pairs or key in seen:
                            continue
                        seen.add(key)
                        nA, nB = self._inst_count(key[0]), self._inst_count(key[1])
                        if nA and nB:
                            out.append((key, nA, nB))
            return out

        need = wanted()
        if not need:
            return
        size = lambda nA, nB: nA * ((nB + 3) // 4 * 4)  # noqa: E731
        total = sum(size(nA, nB) for _, nA, nB in need)
        cap = self._pool.numel() if self._pool is not None else 0
        if self._pool_used + total > cap:
            if before_evict is not None:
                before_evict()
            self._pool_gen += 1
            self.pool_evictions += 1
            cur = torch.cuda.current_stream()
            for st in side_streams_in_use():  # a pipelined batch may still be cutting its A out of the pool there
                cur.wait_stream(st)
            self._pairs.clear()  # evict everything (stream order keeps earlier readers safe)
            self._pool_used = 0
            need = wanted()
            total = sum(size(nA, nB) for _, nA, nB in need)
            if total > cap:
                self._release_pool()
                want = max(total, self._pool_target() // 8)
                parked = _PARKED_POOLS.setdefault(torch.cuda.current_device(), [])
                fit = [t for t in parked if t.numel() >= want]
                if fit:
                    self._pool = min(fit, key=lambda t: t.numel())
                    parked[:] = [t for t in parked if t is not self._pool]  # (list.remove would compare tensors)
                else:
                    self._pool = torch.empty(want, dtype=torch.float64, device="cuda")
        desc = np.zeros(len(need), dtype=PAIRDESC_DTYPE)
        tiles = np.zeros(len(need) + 1, dtype=np.int64)
        points = 0.0
        for q, (key, nA, nB) in enumerate(need):
            a, b = key
            ld = (nB + 3) // 4 * 4
            desc[q] = (int(self.inst_off[a]), nA, int(self.inst_off[b]), nB, self._pool_used, ld,
                       self._pair_lut_idx[(anchor(a), anchor(b))], int(a == b), 0)
            self._pairs[key] = (self._pool_used, ld)
            self._pool_used += nA * ld
            tiles[q + 1] = tiles[q] + ((nA + 31) // 32) * ((nB + 31) // 32)
            points += nA * (nA + 1) / 2 if a == b else nA * nB
        self.pair_points += points
        cfg = self.cfg
        d_desc = h2d(desc.view(np.uint8).reshape(-1))
        d_tiles = h2d(tiles[:-1].astype(np.int32))
        _lib.dev_pair_blocks(ptr(self.d_x), ptr(self.d_y), ptr(self.d_img), ptr(d_desc), ptr(d_tiles), len(need),
                             int(tiles[-1]), ptr(self.d_tables), ptr(self.d_pair_lut), self.blk.n_inimage,
                             self.arena.ngrid, float(cfg.dscale), float(cfg.nc_ovl), float(cfg.flat_penalty),
                             self.arena.poly, ptr(self._pool), float(points), stream_handle())

    def pair_block(self, a, b) -> torch.Tensor:
        """The full InStamp-pair block of A, (npix_a, npix_b) float64 on the device (a <= b in raster order):
        what SysMatA.get_iisubmat returns (psfutil.py:2009-2092).  A view into the pair-block pool: copy it if it has to
        outlive the next ensure_pairs call."""
        assert a <= b, f"ji_st1={a} should precede ji_st2={b}"
        nA, nB = self._inst_count(a), self._inst_count(b)
        if nA == 0 or nB == 0:
            return torch.zeros((nA, nB), dtype=torch.float64, device="cuda")

        if (a, b) not in self._pairs:  # (ensure_pairs reads only .insts of a StampPlan)
            self.ensure_pairs([types.SimpleNamespace(insts=[a] if a == b else [a, b])])
        off, ld = self._pairs[(a, b)]
        return self._pool[off:off + nA * ld].view(nA, ld)[:, :nB]

    def io_block(self, ji_in, selection, x0out: float, y0out: float) -> torch.Tensor:
        """One InStamp's columns of -B/2, (n_out, m, n_selected) float64 on the device, for the output stamp whose
        first pixel sits at (x0out, y0out): what SysMatB.get_iosubmat returns (psfutil.py:2125-2185, 1497-1595).
        selection: indices into the InStamp's pixels (None: all of them)."""
        cfg, tab = self.cfg, self.tab
        nimg = self.blk.n_inimage
        base = int(self.inst_off[ji_in])
        ntot = self._inst_count(ji_in)
        loc = np.arange(ntot, dtype=np.int64) if selection is None else np.asarray(selection, dtype=np.int64)
        n, m = int(loc.size), cfg.n2f**2
        if n == 0:
            return torch.zeros((cfg.n_out, m, 0), dtype=torch.float64, device="cuda")
        G = anchor(ji_in)
        tab.group(G)
        io = tab.get_io(G)
        lut_io = np.full((nimg, cfg.n_out), -1, dtype=np.int64)
        for k in tab.grp_imgs[G]:
            lut_io[k, :] = [self.arena.offset(io, (tab.grp_index(G, k), o)) for o in range(cfg.n_out)]
        npad, mpad = rup(n), rup(m)
        gidx = torch.from_numpy(base + loc).cuda()
        px = torch.zeros(npad, dtype=torch.float64, device="cuda")
        py = torch.zeros(npad, dtype=torch.float64, device="cuda")
        px[:n], py[:n] = self.d_x[gidx], self.d_y[gidx]
        pcode = torch.zeros(npad, dtype=torch.int32, device="cuda")
        pcode[:n] = self.d_img[gidx]
        d_lut = h2d(lut_io)
        mB = torch.empty((cfg.n_out, mpad, npad), dtype=torch.float64, device="cuda")
        _lib.dev_build_B(ptr(px), ptr(py), ptr(pcode), n, npad, ptr(self.d_tables), ptr(d_lut), cfg.n_out,
                         self.arena.ngrid, float(cfg.dscale), float(cfg.nc_ovl), cfg.n2f, mpad, float(x0out), float(y0out),
                         ptr(mB), mB.stride(1), mB.stride(0), stream_handle())
        return mB[:, :m, :n]

    def _release_pool(self):
        """Park the pair-block pool for the next GpuBlock on this device (at most two are kept)."""
        pool, self._pool = getattr(self, "_pool", None), None
        if pool is not None and _PARKED_POOLS is not None:
            parked = _PARKED_POOLS.setdefault(pool.device.index, [])
            if len(parked) < 2:
                parked.append(pool)

    def __del__(self):
        try:
            self._release_pool()
        except Exception:  # interpreter shutdown
            pass

    def _pool_target(self) -> int:
        """Bytes of the pair-block pool: everything the planned OutStamps can ever need if that fits the budget
        (pool_bytes, default 16 GB or a quarter of the free HBM), else the budget (older blocks are then evicted)."""
        if self.pool_bytes:
            return int(self.pool_bytes)
        seen, tot = set(), 0
        for (j_st, i_st) in self.order:  # (stamps still unplanned count too: only the InStamp sizes are needed)
            insts = [(j_st + dj, i_st + di) for dj in (-1, 0, 1) for di in (-1, 0, 1)]
            for a in range(9):
                for b in range(a, 9):
                    key = (insts[a], insts[b])
                    if key not in seen:
                        seen.add(key)
                        tot += self._inst_count(key[0]) * ((self._inst_count(key[1]) + 3) // 4 * 4)
        budget = min(16 << 30, (hbm_free_estimate() // 4) >> 28 << 28)
        return int(min(8 * tot, budget))

    def _asm_desc(self, p) -> _lib.AsmDesc:
        d = _lib.AsmDesc()
        for a in range(9):
            d.inst_off[a] = int(self.inst_off[p.insts[a]])
            for b in range(a, 9):
                off, ld = self._pairs.get((p.insts[a], p.insts[b]), (0, 4))
                d.blk[9 * a + b], d.ld[9 * a + b] = off, ld
        for a in range(10):
            d.seg_start[a] = int(p.inpix_cumsum[a])
        return d

    def build_system(self, k: int, need_A: bool = True, skip_AB: bool = False):
        """Stage (a): gather + A + mBhalf for stamp number k of self.order.  Returns (DeviceSystem, indata).

        skip_AB: the no-quality-control mode of the empirical kernel (coadd.py:1020-1025): neither system matrix is
        interpolated (mBhalf is left as zeros of the right shape).

        need_A=False (pair-block cache on): A is not materialised; the solver cuts A + kappa I straight out of the
        cached blocks through DeviceSystem.assemble."""
        cfg = self.cfg
        p = self.plans[self.order[k]]
        st = stream_handle()
        n, m = p.n, cfg.n2f**2
        npad, mpad = rup(n), rup(m)
        nimg = self.blk.n_inimage
        px = torch.empty(npad, dtype=torch.float64, device="cuda")
        py = torch.empty(npad, dtype=torch.float64, device="cuda")
        indata = torch.empty((cfg.n_inframe, npad), dtype=torch.float32, device="cuda")
        ch, q = self._meta(k)
        o = int(ch["off_pix"][q])
        idx = ch["idx"][o:o + n]
        pcode = ch["pcode"][o:o + n]
        # positions and layers: gathered through idx; the table codes are per (stamp, pixel) and were planned on the host
        _lib.dev_gather_stamp(ptr(idx), n, npad, ptr(self.d_x), ptr(self.d_y), None, ptr(self.d_data),
                              self.d_data.stride(0), cfg.n_inframe, ptr(px), ptr(py), None, ptr(indata),
                              indata.stride(0), st)
        ncode = 4 * nimg
        A, assemble = None, None
        if skip_AB:
            mB = torch.zeros((cfg.n_out, mpad, npad), dtype=torch.float64, device="cuda")
            ds = DeviceSystem(n=n, m=m, n2f=cfg.n2f, A=None, mB=mB, C=np.asarray(self.tab.outovlc, dtype=np.float64),
                              px=px, py=py, assemble=None)
            g = torch.arange(cfg.n2f, dtype=torch.float64, device="cuda")
            ds.outy = (p.y0out + g).repeat_interleave(cfg.n2f).contiguous()
            ds.outx = (p.x0out + g).repeat(cfg.n2f).contiguous()
            return ds, indata
        if self.a_cache:
            self.ensure_pairs([p])  # no-op when coadd_batch has already requested the whole batch's blocks
            desc, pool, gen = self._asm_desc(p), self._pool, self._pool_gen

            def assemble(diag_add, desc=desc, pool=pool, idx=idx, n=n, npad=npad, gen=gen):
                if gen != self._pool_gen:
                    raise RuntimeError("pair-block pool was evicted after this system was planned: its blocks are gone "
                                       "(finish the batch before the next ensure_pairs)")
                W = torch.empty((npad, npad), dtype=torch.float64, device="cuda")
                _lib.dev_assemble_A(C.byref(desc), ptr(idx), n, npad, ptr(pool), ptr(W), W.stride(0), float(diag_add),
                                    stream_handle())
                return W

            if need_A:
                A = assemble(0.0)
        else:
            A = torch.empty((npad, npad), dtype=torch.float64, device="cuda")
            _lib.dev_build_A(ptr(px), ptr(py), ptr(pcode), n, npad, ptr(self.d_tables), ptr(ch["lut"][q]), nimg, ncode,
                             self.arena.ngrid, float(cfg.dscale), float(cfg.nc_ovl), float(cfg.flat_penalty), ptr(A),
                             A.stride(0), 0.0, self.arena.poly, st)
        mB = torch.empty((cfg.n_out, mpad, npad), dtype=torch.float64, device="cuda")
        _lib.dev_build_B(ptr(px), ptr(py), ptr(pcode), n, npad, ptr(self.d_tables), ptr(ch["lut_io"][q]), cfg.n_out,
                         self.arena.ngrid, float(cfg.dscale), float(cfg.nc_ovl), cfg.n2f, mpad, p.x0out, p.y0out, ptr(mB),
                         mB.stride(1), mB.stride(0), st)
        ds = DeviceSystem(n=n, m=m, n2f=cfg.n2f, A=A, mB=mB, C=np.asarray(self.tab.outovlc, dtype=np.float64),
                          px=px, py=py, assemble=assemble)
        if self.kernel in ("Iterative", "Empirical"):
            g = torch.arange(cfg.n2f, dtype=torch.float64, device="cuda")
            ds.outy = (p.y0out + g).repeat_interleave(cfg.n2f).contiguous()
            ds.outx = (p.x0out + g).repeat(cfg.n2f).contiguous()
        return ds, indata

    def apply_spec(self, k: int, indata, want_T32=True, want_Ti64=False) -> ApplySpec:
        cfg = self.cfg
        ch, q = self._meta(k)
        a, b = int(ch["off_seg"][q]), int(ch["off_seg"][q + 1])
        return ApplySpec(fade=cfg.fade_kernel, fade_w=self.d_fade_w, indata=indata, seg_end=ch["seg_end"][a:b],
                         seg_img=ch["seg_img"][a:b], n_img=self.blk.n_inimage, n2=cfg.n2,
                         clamp_iter=(self.kernel == "Iterative"), want_T32=want_T32, want_Ti64=want_Ti64)

    def coadd_stamp(self, k: int, keep: bool = False):
        """OutStamp.__call__ (coadd.py:979-1000) + overlap-add (coadd.py:1976-1994) for stamp k of self.order.

        keep=True returns the per-stamp device results (parity tests); otherwise nothing is retained."""
        return self.coadd_batch([k], keep=keep)[0]

    def coadd_batch(self, ks, keep: bool = False):
        """The OutStamps ks (positions in self.order) together: stage (a) per stamp, then ONE batched solve across
        the stamps (CholKernel: every (stamp, kappa node) system shares the factorisation launches; the other
        kernels run stamp by stamp), then T-apply and overlap-add per stamp.  Stamps are independent given their
        3x3 InStamp neighbourhoods (coadd.py:2056-2060), so the order inside a batch does not matter."""
        cfg = self.cfg
        kept = [dict() for _ in ks]
        live = []
        # single-kappa Cholesky needs A only as A + kappa I: skip materialising it unless the caller keeps the stamp
        need_A = keep or self.kernel != "Cholesky" or len(np.atleast_1d(cfg.kappaC_arr)) > 1
        # EmpirKernel without quality control (coadd.py:856-858, 1020-1025): no system matrices, U/C = Sigma = kappa = 0
        no_qlt = self.kernel == "Empirical" and bool(getattr(cfg, "no_qlt_ctrl", False))
        if self.a_cache and not no_qlt:
            self.ensure_pairs([pl for pl in (self.plans[self.order[k]] for k in ks) if pl.n > 0])
        for q, k in enumerate(ks):
            p = self.plans[self.order[k]]
            if p.n == 0:  # lakernel.py:110-119 and coadd.py:1094-1100: nothing to add except UC = kappa = 1
                self._empty_stamp(p, keep)
                continue
            ds, indata = self.build_system(k, need_A=need_A, skip_AB=no_qlt)
            live.append((q, k, p, ds, indata))
        if self.kernel == "Eigen" and live:  # one decomposition per stamp serves every output PSF; all stamps together
            for (q, k, p, ds, indata), eig in zip(live, eigen_decompose_batch([t[3] for t in live])):
                kept[q]["_eig"] = eig
        # CholKernel: the factorisations of every output PSF are enqueued before the first one is waited for
        # (kappa = kappa/C * C_j differs per output PSF, so each has its own systems: lakernel.py:295-299)
        handles = ([solve_chol_launch([t[3] for t in live], cfg, j) for j in range(cfg.n_out)]
                   if self.kernel == "Cholesky" and live else [])
        for j_out in range(cfg.n_out):
            if self.kernel == "Cholesky":
                kos = solve_chol_finish(handles[j_out]) if live else []
                if live:
                    handles[j_out] = None
            for u, (q, k, p, ds, indata) in enumerate(live):
                if self.kernel == "Cholesky":
                    ko = kos[u]
                elif self.kernel == "Eigen":
                    ko = solve_eigen(ds, cfg, j_out, eig=kept[q]["_eig"])
                elif no_qlt:
                    ko = SOLVERS[self.kernel](ds, cfg, j_out, no_qlt_ctrl=True)
                else:
                    ko = SOLVERS[self.kernel](ds, cfg, j_out)
                spec = self.apply_spec(k, indata, want_T32=keep, want_Ti64=keep)
                res = apply_T(ds, ko, j_out, spec)
                self._overlap_add(p, j_out, res)
                if keep:
                    kept[q][j_out] = dict(res=res, ko=ko)
            if self.kernel == "Cholesky":
                del kos
        for (q, k, p, ds, indata) in live:
            kept[q].pop("_eig", None)
            if keep:
                kept[q]["ds"], kept[q]["indata"], kept[q]["plan"] = ds, indata, p
        return kept

    def _overlap_add(self, p, j_out, res):
        """coadd.py:1976-1994: add one stamp's faded results into the block maps."""
        cfg = self.cfg
        y0, x0 = (p.j_st - 1) * cfg.n2, (p.i_st - 1) * cfg.n2
        tw = self.T_weightmap[j_out, :, p.j_st - 1, p.i_st - 1]  # (n_inimage,) strided view
        _lib.dev_accumulate_stamp(ptr(res["outimage"]), cfg.n_inframe, ptr(res["UC"]), ptr(res["Sigma"]),
                                  ptr(res["kappa"]), ptr(res["Tsum_inpix"]), ptr(res["Neff"]), ptr(res["Tsum_stamp"]),
                                  self.blk.n_inimage, cfg.n2f, ptr(self.out_map[j_out]), ptr(self.UC_map[j_out]),
                                  ptr(self.Sigma_map[j_out]), ptr(self.kappa_map[j_out]), ptr(self.Tsum_map[j_out]),
                                  ptr(self.Neff_map[j_out]), self.side, y0, x0, ptr(tw), tw.stride(0), stream_handle())

    def _empty_stamp(self, p, keep):
        """An OutStamp without input pixels (lakernel.py:110-119, coadd.py:1094-1100): UC = kappa = 1 (faded), nothing
        else.  Deliberate difference: the reference's _perform_coaddition then forms Neff = 1 / sum((Tsum / sum|Tsum|)^2)
        from empty sums, i.e. 0/0 = NaN, and adds that NaN into Neff_map; here an empty stamp adds nothing to Neff_map."""
        cfg = self.cfg
        y0, x0 = (p.j_st - 1) * cfg.n2, (p.i_st - 1) * cfg.n2
        sl = (slice(None), slice(y0, y0 + cfg.n2f), slice(x0, x0 + cfg.n2f))
        w = trapezoid_weights(cfg.fade_kernel)
        one = np.ones((cfg.n2f, cfg.n2f), dtype=np.float32)
        if cfg.fade_kernel > 0:
            fk2 = 2 * cfg.fade_kernel
            for arr in (one,):
                arr[:fk2, :] *= w[:, None]
                arr[::-1][:fk2, :] *= w[:, None]
                arr[:, :fk2] *= w[None, :]
                arr[:, ::-1][:, :fk2] *= w[None, :]
        t = torch.from_numpy(one).cuda()
        self.UC_map[sl] += t
        self.kappa_map[sl] += t
        return {}

    def batch_size(self) -> int:
        """OutStamps solved together: as many as MAXB systems allow, within an HBM budget (A, W, mBhalf and X of
        every stamp of the batch are live at once: ~(2 npad^2 + 2 n_out mpad npad) * 8 bytes per kappa node)."""
        if self.kernel == "Eigen" and self.order:
            # per stamp: A, its working copy, Vt, Q, the solver's saved copy, Gram matrix and transposed vectors
            # (7 npad^2), the inverse-iteration factors (5 n npad), -B/2, P, tt and T per output PSF
            nmax = max(p.n for p in self.plans.values())
            npad, mpad = rup(nmax), rup(self.cfg.n2f**2)
            per = 8.0 * (12 * npad * npad + (3 + self.cfg.n_out) * mpad * npad)
            return int(max(1, min(0.5 * hbm_free_estimate() // per, self.max_batch, _lib.MAXB)))
        if self.kernel != "Cholesky" or not self.order:
            return 1
        cfg = self.cfg
        nv = max(1, len(np.atleast_1d(cfg.kappaC_arr)))
        nmax = max(p.n for p in self.plans.values())  # of the stamps planned so far (at least the first chunk)
        npad, mpad = rup(nmax), rup(cfg.n2f**2)
        # A once; W and the digit planes of L per system; -B/2 per output PSF; X and the digit planes of Z per system
        # (systems per stamp: kappa nodes x output PSFs -- every output PSF's factorisations are enqueued together)
        nsys = nv * cfg.n_out
        per = 8.0 * ((1 + 2 * nsys) * npad * npad + (cfg.n_out + 2 * nsys) * mpad * npad)
        free = hbm_free_estimate()
        return int(max(1, min(0.5 * free // per, self.max_batch)))

    # 32 OutStamps per batch = two groups of 16 systems on the two solve streams: every serial k_potrf_diag launch and
    # every partial last wave then serves 16 systems instead of 8 (64-stamp block 281.8 -> 276.1 ms, 256 stamps
    # 1100.9 -> 1075.7 ms; 48 stamps on 3 streams 1070.0, 64 on 2 / 4 streams 1092.7 / 1077.6)
    max_batch = int(os.environ.get("B200_MAX_BATCH", "32"))

    def run(self, batch: int | None = None):
        """coadd_output_stamps(sim_mode=False) (coadd.py:2056-2069): every planned stamp, in batches of
        independent stamps (order inside the block is the reference's)."""
        assert self._uploaded, "call prepare() first"
        nb = batch or self.batch_size()
        batches = [list(range(k0, min(k0 + nb, len(self.order)))) for k0 in range(0, len(self.order), nb)]
        if self.kernel == "Cholesky" and self.cfg.n_out == 1 and len(batches) > 1 and self.a_cache:
            return self._run_pipelined(batches)
        for ks in batches:
            self.coadd_batch(ks)
        return self

    def _run_pipelined(self, batches):
        """Software pipeline over batches (CholKernel, one output PSF): the stage-(a) kernels of batch k+1 are enqueued
        on the caller's stream before anybody waits for batch k, whose factorisation runs on the solve streams; the
        T-apply of batch k then overlaps the factorisation of batch k+1.  The solve streams never run dry and their
        partial waves are filled by the interpolation / apply kernels."""
        cfg = self.cfg
        need_A = len(np.atleast_1d(cfg.kappaC_arr)) > 1
        pending = None

        def finish(pend):
            live, handle = pend
            kos = solve_chol_finish(handle)
            for u, (k, p, ds, indata) in enumerate(live):
                spec = self.apply_spec(k, indata, want_T32=False, want_Ti64=False)
                self._overlap_add(p, 0, apply_T(ds, kos[u], 0, spec))

        def finish_pending():
            nonlocal pending
            if pending is not None:
                finish(pending)
                pending = None

        for b, ks in enumerate(batches):
            plans = [self.plans[self.order[k]] for k in ks]
            # (if the pool has to be evicted for this batch, the previous batch is completed first: its repair branch
            # would otherwise re-assemble A from overwritten blocks)
            self.ensure_pairs([pl for pl in plans if pl.n > 0], before_evict=finish_pending)
            live = []
            for k, p in zip(ks, plans):
                if p.n == 0:
                    self._empty_stamp(p, False)
                    continue
                ds, indata = self.build_system(k, need_A=need_A)
                live.append((k, p, ds, indata))
            handle = solve_chol_launch([t[2] for t in live], cfg, 0)
            if b + 1 < len(batches):  # host planning of the next batch hides behind the work just enqueued
                for k in (batches[b + 1][0], batches[b + 1][-1]):
                    self._ensure_planned(k)
            finish_pending()
            pending = (live, handle)
        finish_pending()
        return self

    def build_output(self, is_final: bool = True, pad_sides: str = "", download: bool = True):
        """Block.build_output_file up to the FITS container for this block's maps: see assemble_output."""
        names = ("out_map", "T_weightmap", "UC_map", "Sigma_map", "kappa_map", "Tsum_map", "Neff_map")
        return assemble_output({nm: getattr(self, nm) for nm in names}, self.cfg, self.blk.n_inimage, is_final,
                               pad_sides, download)

    def download(self):
        """Block maps to host (the final gather of the output cube starts from these)."""
        torch.cuda.current_stream().synchronize()
        names = ("out_map", "T_weightmap", "UC_map", "Sigma_map", "kappa_map", "Tsum_map", "Neff_map")
        return {nm: getattr(self, nm).cpu().numpy() for nm in names}


# quality maps of cfg.outmaps: letter -> (block map, EXTNAME, coefficient, unsigned) (coadd.py:2245-2303)
OUTPUT_ENCODING = {"U": ("UC_map", "FIDELITY", -5000, True), "S": ("Sigma_map", "SIGMA", -10000, False),
                   "K": ("kappa_map", "KAPPA", -5000, True), "T": ("Tsum_map", "INWTSUM", 200000, False),
                   "N": ("Neff_map", "EFFCOVER", 50000, True)}


def assemble_output(maps, cfg, n_inimage: int, is_final: bool = True, pad_sides: str = "", download: bool = True):
    """Block.build_output_file up to the FITS container (coadd.py:2139-2303), on the device.

    maps: float32 CUDA tensors out_map (n_out, n_inframe, side, side), T_weightmap (n_out, n_inimage, n1P, n1P) and the
    quality maps UC_map / Sigma_map / kappa_map / Tsum_map / Neff_map (n_out, side, side), side = NsideP + 2 fade_kernel.
    The faded block boundary is recovered (is_final), the fade margin cropped, and the quality maps named in
    cfg.outmaps are log-encoded to (u)int16.  Returns {EXTNAME: array}: PRIMARY, INWEIGHT, INWTFLAT, FIDELITY, SIGMA,
    KAPPA, INWTSUM, EFFCOVER, as NumPy arrays (download=True: 2 bytes per quality-map pixel cross PCIe instead of 4)
    or device tensors (the unsigned codes then sit in int16 storage).  The input maps are not modified.
    pad_sides: the sides this block pads ("BTLR" subset, Block._handle_postage_pad, coadd.py:1810-1837)."""
    st = stream_handle()
    fk = cfg.fade_kernel
    side = cfg.NsideP + 2 * fk
    so = side - 2 * fk
    width = cfg.postage_pad * cfg.n2
    d_w = h2d(trapezoid_weights(fk)) if fk > 0 else None

    def crop(src, pads):
        flat = src.contiguous().reshape(-1, side, side)
        dst = torch.empty((flat.shape[0], so, so), dtype=torch.float32, device="cuda")
        _lib.dev_unfade_crop(ptr(flat), flat.shape[0], side, fk, int(bool(is_final)), *pads, ptr(d_w), ptr(dst), st)
        return dst.reshape(tuple(src.shape[:-2]) + (so, so))

    T_w = maps["T_weightmap"]
    n_out = T_w.shape[0]
    out = {"PRIMARY": crop(maps["out_map"], (0, 0, 0, 0)), "INWEIGHT": T_w.clone(),
           "INWTFLAT": T_w.permute(0, 2, 1, 3).reshape(n_out * cfg.n1P, n_inimage * cfg.n1P)}
    pads = tuple(width * (sd not in pad_sides) for sd in "BTLR")
    outmaps = getattr(cfg, "outmaps", "USKTN")
    for letter in "USKTN":
        if letter not in outmaps:
            continue
        name, ext, coef, unsigned = OUTPUT_ENCODING[letter]
        cropped = crop(maps[name], pads)
        # (torch has no uint16 arithmetic: unsigned codes are written into int16 storage and re-viewed on the host)
        codes = torch.empty(cropped.shape, dtype=torch.int16, device="cuda")
        _lib.dev_compress_map(ptr(cropped), cropped.numel(), coef, int(unsigned), ptr(codes), st)
        out[ext] = codes
    if not download:
        return out
    torch.cuda.current_stream().synchronize()
    host = {k: v.contiguous().cpu().numpy() for k, v in out.items()}
    for name, ext, coef, unsigned in OUTPUT_ENCODING.values():
        if unsigned and ext in host:
            host[ext] = host[ext].view(np.uint16)
    return host


class GpuOutStamp:
    """Host view of one coadded OutStamp with the reference's attribute names (coadd.py:795-1363)."""

    def __init__(self, gblk: GpuBlock, j_st: int, i_st: int):
        k = gblk.order.index((j_st, i_st))
        cfg = gblk.cfg
        kept = gblk.coadd_stamp(k, keep=True)
        torch.cuda.synchronize()
        p = gblk.plans[(j_st, i_st)]
        n, m, n2f = p.n, cfg.n2f**2, cfg.n2f
        self.inpix_cumsum = p.inpix_cumsum
        ds = kept["ds"]
        self.sysmata = ds.A[:n, :n].cpu().numpy()
        self.mhalfb = ds.mB[:, :m, :n].cpu().numpy()
        self.outovlc = np.asarray(gblk.tab.outovlc)
        self.indata = kept["indata"][:, :n].cpu().numpy()
        n_out = cfg.n_out
        g = lambda key, j: kept[j]["res"][key].cpu().numpy()  # noqa: E731
        self.T = np.stack([g("T32", j)[:, :n] for j in range(n_out)])
        self.Ti64 = np.stack([g("Ti64", j)[:, :n] for j in range(n_out)])
        self.UC = np.stack([g("UC", j).reshape(n2f, n2f) for j in range(n_out)])
        self.Sigma = np.stack([g("Sigma", j).reshape(n2f, n2f) for j in range(n_out)])
        self.kappa = np.stack([g("kappa", j).reshape(n2f, n2f) for j in range(n_out)])
        self.outimage = np.stack([g("outimage", j).reshape(cfg.n_inframe, n2f, n2f) for j in range(n_out)])
        self.Tsum_stamp = np.stack([g("Tsum_stamp", j)[: gblk.blk.n_inimage] for j in range(n_out)])
        self.Tsum_inpix = np.stack([g("Tsum_inpix", j).reshape(n2f, n2f) for j in range(n_out)])
        self.Neff = np.stack([g("Neff", j).reshape(n2f, n2f) for j in range(n_out)])
        self.extras = [
            {k_: (v.cpu().numpy() if torch.is_tensor(v) else v) for k_, v in kept[j]["ko"].extras.items()}
            for j in range(n_out)
        ]
