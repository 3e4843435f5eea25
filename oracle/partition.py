"""TEST INFRASTRUCTURE -- CPU restatement of the input-pixel partitioning.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does.

Follows InImage.partition_pixels and InImage.extract_layers of the reference (coadd.py:174-380, 382-408) with the same
loops in the same order (pure Python: small cases only).  The WCS composition ``_inpix2world2outpix`` (astropy / gwcs,
coadd.py:133-150) is an input: ``outpix(inxys) -> (npix, 2)``.  Pinned by tests/golden/partition.npz, which the
reference's own InImage.partition_pixels / extract_layers produced with the same map and masks
(tests/golden/make_golden.py).
"""

import numpy as np

PIXSCALE_NATIVE_ARCSEC = 0.11  # config.py:97
SCA_NSIDE = 4088  # config.py:98


def idx_grid(xs, ys):
    """InImage.generate_idx_grid (coadd.py:112-131): all (x, y) combinations, x fastest."""
    return np.moveaxis(np.array(np.meshgrid(xs, ys)), 0, -1).reshape(-1, 2)


def npixmax_of(cfg, relax_coef=1.05):
    """coadd.py:265-273."""
    return int(((cfg.n2 * cfg.dtheta * 3600.0) / PIXSCALE_NATIVE_ARCSEC + 1) ** 2 * relax_coef)


def relevant_cells(outpix, cfg, use_instamps, sca_nside=SCA_NSIDE, sp_res=90):
    """The sparse-grid pass (coadd.py:199-232): (sp_arr, relevant_matrix, is_relevant)."""
    sp_arr = np.linspace(0, sca_nside, sp_res + 1, dtype=np.uint16)
    sp_outxys = outpix(idx_grid(sp_arr, sp_arr)).T.reshape(2, sp_res + 1, sp_res + 1)
    pix_lower = -cfg.n2 - 0.5
    pix_upper = cfg.NsideP + cfg.n2 - 0.5
    ns = cfg.n1P + 2
    is_relevant = False
    relevant = np.zeros((sp_res, sp_res), dtype=bool)
    for j in range(1, sp_res):
        for i in range(1, sp_res):
            if not (pix_lower < sp_outxys[0, j, i] < pix_upper and pix_lower < sp_outxys[1, j, i] < pix_upper):
                continue
            i_st = int((sp_outxys[0, j, i] - pix_lower) // cfg.n2)
            j_st = int((sp_outxys[1, j, i] - pix_lower) // cfg.n2)
            if np.any(use_instamps[max(j_st - 2, 0):min(j_st + 3, ns), max(i_st - 2, 0):min(i_st + 3, ns)]):
                is_relevant = True
                relevant[max(j - 2, 0):min(j + 3, sp_res), max(i - 2, 0):min(i + 3, sp_res)] = True
    return sp_arr, relevant, is_relevant


def partition_pixels(outpix, mask, cfg, use_instamps, sca_nside=SCA_NSIDE, sp_res=90, relax_coef=1.05):
    """InImage.partition_pixels (coadd.py:174-380) for one input image.

    mask: (sca_nside, sca_nside) bool, the AND of all the masks the reference combines (coadd.py:286-330).
    Returns dict(is_relevant, pix_count, y_idx, x_idx, y_val, x_val, max_count) with the reference's shapes/dtypes."""
    sp_arr, relevant, is_relevant = relevant_cells(outpix, cfg, use_instamps, sca_nside, sp_res)
    if not is_relevant:
        return dict(is_relevant=False)
    ns = cfg.n1P + 2
    npixmax = npixmax_of(cfg, relax_coef)
    y_idx = np.zeros((ns, ns, npixmax), dtype=np.uint16)
    x_idx = np.zeros((ns, ns, npixmax), dtype=np.uint16)
    y_val = np.zeros((ns, ns, npixmax), dtype=np.float64)
    x_val = np.zeros((ns, ns, npixmax), dtype=np.float64)
    pix_count = np.zeros((ns, ns), dtype=np.uint32)
    pix_lower = -cfg.n2 - 0.5
    pix_upper = cfg.NsideP + cfg.n2 - 0.5
    for j_sp in range(sp_res):
        for i_sp in range(sp_res):
            if not relevant[j_sp, i_sp]:
                continue
            left, right = (int(v) for v in sp_arr[i_sp:i_sp + 2])
            bottom, top = (int(v) for v in sp_arr[j_sp:j_sp + 2])
            inxys = idx_grid(np.arange(left, right), np.arange(bottom, top))
            outxys = outpix(inxys).T.reshape(2, top - bottom, right - left)
            for j in range(top - bottom):
                for i in range(right - left):
                    my_x, my_y = outxys[:, j, i]
                    if not (pix_lower < my_x < pix_upper and pix_lower < my_y < pix_upper):
                        continue
                    if not mask[bottom + j, left + i]:
                        continue
                    i_st = int((my_x - pix_lower) // cfg.n2)
                    j_st = int((my_y - pix_lower) // cfg.n2)
                    if not use_instamps[j_st, i_st]:
                        continue
                    k = pix_count[j_st, i_st]
                    y_idx[j_st, i_st, k] = bottom + j
                    x_idx[j_st, i_st, k] = left + i
                    y_val[j_st, i_st, k] = my_y
                    x_val[j_st, i_st, k] = my_x
                    pix_count[j_st, i_st] += 1
    return dict(is_relevant=True, pix_count=pix_count, y_idx=y_idx, x_idx=x_idx, y_val=y_val, x_val=x_val,
                max_count=int(np.max(pix_count)), sp_arr=sp_arr, relevant=relevant)


def extract_layers(indata, part, cfg):
    """InImage.extract_layers (coadd.py:382-408): data (n_inframe, ns, ns, max_count) float32."""
    ns = cfg.n1P + 2
    data = np.zeros((indata.shape[0], ns, ns, part["max_count"]), dtype=np.float32)
    for j_st in range(ns):
        for i_st in range(ns):
            n = part["pix_count"][j_st, i_st]
            data[:, j_st, i_st, :n] = indata[:, part["y_idx"][j_st, i_st, :n], part["x_idx"][j_st, i_st, :n]]
    return data
