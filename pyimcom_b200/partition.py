"""Input-pixel partitioning on the device: InImage.partition_pixels + extract_layers.

The reference maps every detector pixel of the relevant sparse-grid cells through the WCS composition
``_inpix2world2outpix`` (astropy / gwcs, vectorised) and then walks them in two nested Python loops to append each to
the list of its postage stamp (coadd.py:333-360) -- on a full block that loop costs more than the whole device
coaddition.  Here the WCS call stays where it is (``outpix``, any callable (npix, 2) -> (npix, 2)); the sparse-grid
relevance pass (coadd.py:199-232, a few thousand points) is done with NumPy on the host; the binning, with the
reference's list order, and the layer extraction run on the device (csrc/partition.cu).

There is no CPU fallback for the binning: without the CUDA library this module cannot be imported.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .coadd import h2d
from .lakernel import ptr, stream_handle

PIXSCALE_NATIVE_ARCSEC = 0.11  # config.py:97
SCA_NSIDE = 4088  # config.py:98
CELL_DTYPE = np.dtype([("bottom", np.int32), ("left", np.int32), ("h", np.int32), ("w", np.int32), ("off", np.int64)])
assert CELL_DTYPE.itemsize == C.sizeof(_lib.PartCell)


def sparse_relevance(cfg, use_instamps, sp_arr, outpix):
    """The sparse-grid pass of InImage.partition_pixels (coadd.py:199-232) with NumPy on the host (a few thousand
    points): (relevant_matrix (sp_res, sp_res) bool, is_relevant)."""
    ns, R = cfg.n1P + 2, len(sp_arr) - 1
    pix_lower, pix_upper = -cfg.n2 - 0.5, cfg.NsideP + cfg.n2 - 0.5
    xs, ys = np.meshgrid(sp_arr, sp_arr)
    sp_out = np.asarray(outpix(np.stack([xs.ravel(), ys.ravel()], axis=1)), dtype=np.float64)
    sx, sy = sp_out[:, 0].reshape(R + 1, R + 1), sp_out[:, 1].reshape(R + 1, R + 1)
    inside = (pix_lower < sx) & (sx < pix_upper) & (pix_lower < sy) & (sy < pix_upper)
    inside[0, :] = inside[:, 0] = False  # the reference's loops run over 1 .. sp_res-1
    inside[R:, :] = inside[:, R:] = False
    relevant = np.zeros((R, R), dtype=bool)
    is_relevant = False
    for j, i in zip(*np.nonzero(inside)):
        i_st = int(np.floor_divide(sx[j, i] - pix_lower, cfg.n2))
        j_st = int(np.floor_divide(sy[j, i] - pix_lower, cfg.n2))
        if np.any(use_instamps[max(j_st - 2, 0):min(j_st + 3, ns), max(i_st - 2, 0):min(i_st + 3, ns)]):
            is_relevant = True
            relevant[max(j - 2, 0):min(j + 3, R), max(i - 2, 0):min(i + 3, R)] = True
    return relevant, is_relevant


class DevicePartition:
    """InImage.partition_pixels / extract_layers for the input images of one block.

    cfg needs n2, n1P, NsideP, dtheta (degrees); use_instamps is Block.use_instamps ((n1P+2)^2 bool)."""

    def __init__(self, cfg, use_instamps, sca_nside: int = SCA_NSIDE, sp_res: int = 90, relax_coef: float = 1.05):
        if not torch.cuda.is_available():
            raise RuntimeError("pyimcom_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.ns = cfg.n1P + 2
        self.use = np.ascontiguousarray(np.asarray(use_instamps, dtype=bool))
        assert self.use.shape == (self.ns, self.ns)
        self.sca, self.sp_res = int(sca_nside), int(sp_res)
        self.pix_lower = -cfg.n2 - 0.5  # coadd.py:206-207
        self.pix_upper = cfg.NsideP + cfg.n2 - 0.5
        # coadd.py:265-273
        self.npixmax = int(((cfg.n2 * cfg.dtheta * 3600.0) / PIXSCALE_NATIVE_ARCSEC + 1) ** 2 * relax_coef)
        self.sp_arr = np.linspace(0, self.sca, self.sp_res + 1, dtype=np.uint16)
        self.d_use = h2d(self.use.astype(np.uint8))

    def relevant_cells(self, outpix):
        """(relevant_matrix, is_relevant): which cells of the sparse grid can hold pixels of the block."""
        return sparse_relevance(self.cfg, self.use, self.sp_arr, outpix)

    # ---- one input image ----
    def partition(self, outpix, mask, indata=None):
        """Returns dict(is_relevant, pix_count, y_idx, x_idx, y_val, x_val, max_count[, data]) with the reference's
        shapes and dtypes (coadd.py:276-283, 389-392) as CUDA tensors (u16 index arrays sit in int16 storage: view
        them as uint16 on the host).  mask: (sca, sca) bool, the AND of the reference's masks; indata (n_inframe, sca,
        sca) float32, host or device."""
        relevant, is_relevant = self.relevant_cells(outpix)
        if not is_relevant:
            return dict(is_relevant=False)
        ns, st = self.ns, stream_handle()
        cells_ji = np.argwhere(relevant)  # raster order: j_sp outer, i_sp inner (coadd.py:333-337)
        cells = np.zeros(len(cells_ji), dtype=CELL_DTYPE)
        edges = self.sp_arr.astype(np.int64)
        cells["bottom"], cells["left"] = edges[cells_ji[:, 0]], edges[cells_ji[:, 1]]
        cells["h"] = edges[cells_ji[:, 0] + 1] - cells["bottom"]
        cells["w"] = edges[cells_ji[:, 1] + 1] - cells["left"]
        sizes = cells["h"].astype(np.int64) * cells["w"]
        cells["off"] = np.concatenate([[0], np.cumsum(sizes)[:-1]])
        ntot = int(sizes.sum())
        # detector coordinates of every pixel of every relevant cell, in the traversal order, through the WCS at once
        inxy = np.empty((ntot, 2), dtype=np.float64)
        for c in cells:
            o, h, w = int(c["off"]), int(c["h"]), int(c["w"])
            inxy[o:o + h * w, 0] = np.tile(np.arange(c["left"], c["left"] + w), h)
            inxy[o:o + h * w, 1] = np.repeat(np.arange(c["bottom"], c["bottom"] + h), w)
        out = np.asarray(outpix(inxy), dtype=np.float64)
        d_ox, d_oy = h2d(np.ascontiguousarray(out[:, 0])), h2d(np.ascontiguousarray(out[:, 1]))
        d_cells = h2d(cells.view(np.uint8).reshape(-1))
        d_mask = mask if torch.is_tensor(mask) else h2d(np.ascontiguousarray(mask, dtype=bool).astype(np.uint8))
        dev = "cuda"
        ncell = len(cells)
        i32, u32 = torch.int32, torch.int32  # (u32 counters live in int32 storage)
        sid_tmp = torch.empty(max(ntot, 1), dtype=i32, device=dev)
        rank_tmp = torch.empty(max(ntot, 1), dtype=u32, device=dev)
        cellmeta = torch.empty((max(ncell, 1), 4), dtype=i32, device=dev)
        cellcnt = torch.zeros((max(ncell, 1), _lib.PART_MAXSLOT), dtype=u32, device=dev)
        cellbase = torch.zeros((max(ncell, 1), _lib.PART_MAXSLOT), dtype=u32, device=dev)
        run = torch.empty(ns * ns, dtype=u32, device=dev)
        pix_count = torch.empty((ns, ns), dtype=u32, device=dev)
        y_idx = torch.zeros((ns, ns, self.npixmax), dtype=torch.int16, device=dev)
        x_idx = torch.zeros((ns, ns, self.npixmax), dtype=torch.int16, device=dev)
        y_val = torch.zeros((ns, ns, self.npixmax), dtype=torch.float64, device=dev)
        x_val = torch.zeros((ns, ns, self.npixmax), dtype=torch.float64, device=dev)
        err = torch.empty(1, dtype=i32, device=dev)
        _lib.dev_partition(ptr(d_cells), ncell, ptr(d_ox), ptr(d_oy), ptr(d_mask), self.sca, ptr(self.d_use), ns,
                           int(self.cfg.n2), float(self.pix_lower), float(self.pix_upper), self.npixmax, ptr(sid_tmp),
                           ptr(rank_tmp), ptr(cellmeta), ptr(cellcnt), ptr(cellbase), ptr(run), ptr(pix_count),
                           ptr(y_idx), ptr(x_idx), ptr(y_val), ptr(x_val), ptr(err), st)
        counts = pix_count.cpu().numpy().astype(np.uint32)  # one small read-back: max_count sizes the layer array
        code = int(err.item())
        if code == 1:
            raise _lib.B200Error(f"partition: a sparse-grid cell touches more than {_lib.PART_MAXSLOT} postage stamps")
        if code == 2:
            raise IndexError(f"partition: a postage stamp received more than npixmax = {self.npixmax} input pixels "
                             "(the reference's arrays overflow at the same point, coadd.py:354-358)")
        res = dict(is_relevant=True, pix_count=counts, y_idx=y_idx, x_idx=x_idx, y_val=y_val, x_val=x_val,
                   max_count=int(counts.max()), n_positions=ntot, n_cells=ncell, pix_count_dev=pix_count)
        if indata is not None:
            d_in = indata if torch.is_tensor(indata) else h2d(np.ascontiguousarray(indata, dtype=np.float32))
            nfr = d_in.shape[0]
            data = torch.empty((nfr, ns, ns, max(res["max_count"], 0)), dtype=torch.float32, device=dev)
            _lib.dev_extract_layers(ptr(d_in), nfr, self.sca, ptr(y_idx), ptr(x_idx), ptr(pix_count), ns * ns,
                                    self.npixmax, res["max_count"], ptr(data), st)
            res["data"] = data
        return res


def to_host(res):
    """The partition of one image as NumPy arrays with the reference's dtypes."""
    out = {}
    for k, v in res.items():
        if torch.is_tensor(v):
            a = v.cpu().numpy()
            out[k] = a.view(np.uint16) if a.dtype == np.int16 else a
        else:
            out[k] = v
    return out


class DeviceInStamp:
    """Host view of one InStamp (coadd.py:656-749) whose pixels were binned on the device: positions and counts for the
    host-side planning (make_selection, PSF-group bookkeeping); the layers stay in HBM (``data`` is None)."""

    def __init__(self, j_st, i_st, pix_count, x_val, y_val, n2):
        self.j_st, self.i_st = j_st, i_st
        self.pix_count = np.asarray(pix_count, dtype=np.uint32)
        self.pix_cumsum = np.cumsum([0] + list(self.pix_count), dtype=np.uint32)  # coadd.py:691
        self.x_val, self.y_val, self.data = x_val, y_val, None
        if j_st % 2 == 0 and i_st % 2 == 0:  # coadd.py:709-714
            self.psf_compute_point_pix = [i_st * n2 - 0.5, j_st * n2 - 0.5]

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


class PartitionedBlock:
    """Duck-typed coadd.Block for coadd.GpuBlock whose InStamps come straight from DevicePartition results.

    InStamp.__init__ (coadd.py:682-707) concatenates, per postage stamp, the lists of all input images; here each
    image's lists are copied on the device into the block's concatenated pixel arrays (``device_pixels``: exactly what
    GpuBlock would otherwise build on the host and upload), and only the positions come back for the planning.
    parts[k] is DevicePartition.partition(..., indata) of input image k (``is_relevant`` False: no pixels)."""

    def __init__(self, cfg, inimages, parts, outwcs=None, this_sub=0):
        self.cfg, self.inimages, self.n_inimage = cfg, list(inimages), len(inimages)
        self.outwcs, self.this_sub = outwcs, this_sub
        ns = cfg.n1P + 2
        nst, nimg = ns * ns, self.n_inimage
        counts = np.zeros((nimg, nst), dtype=np.int64)
        for k, part in enumerate(parts):
            if part.get("is_relevant", False):
                counts[k] = np.asarray(part["pix_count"], dtype=np.int64).ravel()
        per_stamp = counts.sum(axis=0)
        inst_off = np.concatenate([[0], np.cumsum(per_stamp)[:-1]]).astype(np.int64)
        within = np.cumsum(counts, axis=0) - counts  # exclusive over the images of one stamp
        npix = int(per_stamp.sum())
        dev, st = "cuda", stream_handle()
        gx = torch.empty(max(npix, 1), dtype=torch.float64, device=dev)
        gy = torch.empty(max(npix, 1), dtype=torch.float64, device=dev)
        gimg = torch.empty(max(npix, 1), dtype=torch.int32, device=dev)
        gdata = torch.empty((cfg.n_inframe, max(npix, 1)), dtype=torch.float32, device=dev)
        for k, part in enumerate(parts):
            if not part.get("is_relevant", False) or part["max_count"] == 0:
                continue
            data = part["data"]
            assert data.shape[0] == cfg.n_inframe and data.shape[-1] == part["max_count"]
            d_off = h2d(inst_off + within[k])
            _lib.dev_assemble_instamps(ptr(part["x_val"]), ptr(part["y_val"]), ptr(data), cfg.n_inframe, nst,
                                       part["x_val"].shape[-1], part["max_count"], ptr(part["pix_count_dev"]),
                                       ptr(d_off), k, npix, ptr(gx), ptr(gy), ptr(gimg), ptr(gdata), st)
        torch.cuda.current_stream().synchronize()
        h_x, h_y = gx[:npix].cpu().numpy(), gy[:npix].cpu().numpy()
        h_img = gimg[:npix].cpu().numpy()
        self.device_pixels = dict(d_x=gx[:npix], d_y=gy[:npix], d_img=gimg[:npix], d_data=gdata[:, :npix], h_x=h_x,
                                  h_y=h_y, h_img=h_img, inst_off=inst_off.reshape(ns, ns), npix_total=npix)
        self.instamps = [[None] * ns for _ in range(ns)]
        for j in range(ns):
            for i in range(ns):
                sid = j * ns + i
                a, b = int(inst_off[sid]), int(inst_off[sid] + per_stamp[sid])
                self.instamps[j][i] = DeviceInStamp(j, i, counts[:, sid], h_x[a:b], h_y[a:b], cfg.n2)

    def stamp_order(self):
        """OutStamp traversal of coadd_output_stamps: 2x2 groups (coadd.py:2056-2060)."""
        n1P = self.cfg.n1P
        for j in range(1, n1P + 1, 2):
            for i in range(1, n1P + 1, 2):
                for dj in range(2):
                    for di in range(2):
                        if j + dj <= n1P and i + di <= n1P:
                            yield (j + dj, i + di)
