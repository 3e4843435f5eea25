"""TEST INFRASTRUCTURE -- CPU restatement of the block output assembly.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does.

Follows Block.build_output_file / Block.compress_map / OutStamp.trapezoid(recover_mode=True) of the reference
(coadd.py:2139-2296, 2086-2137, 1222-1292) with plain NumPy, float32 maps in, the arrays of the output HDUs out (the FITS
container itself -- astropy -- is out of scope).  Pinned by tests/golden/output.npz, which the reference's own
``Block.compress_map`` and ``OutStamp.trapezoid`` produced on the same seeded maps (tests/golden/make_golden.py).
"""

import numpy as np

# EXTNAME, coefficient, dtype per quality-map letter of cfg.outmaps (coadd.py:2245-2303)
ENCODING = {
    "U": ("FIDELITY", -5000, np.uint16),
    "S": ("SIGMA", -10000, np.int16),
    "K": ("KAPPA", -5000, np.uint16),
    "T": ("INWTSUM", 200000, np.int16),
    "N": ("EFFCOVER", 50000, np.uint16),
}


def trapezoid_weights(fade_kernel: int) -> np.ndarray:
    """coadd.py:1269-1271 (truncated-sinc trapezoid)."""
    fk2 = 2 * fade_kernel
    s = np.arange(1, fk2 + 1, dtype=np.float64) / (fk2 + 1)
    s -= np.sin(2 * np.pi * s) / (2 * np.pi)
    return s


def recover(arr: np.ndarray, fade_kernel: int, pad_widths=(0, 0, 0, 0)) -> None:
    """In-place OutStamp.trapezoid(arr, fk, recover_mode=True, pad_widths) (coadd.py:1254-1292)."""
    fk2 = 2 * fade_kernel
    if not fk2 > 0:
        return
    ny, nx = arr.shape[-2:]
    pb, pt, pl, pr = pad_widths
    it, ir = ny - pt - 1, nx - pr - 1
    s = trapezoid_weights(fade_kernel)
    sT = s[None, :].T
    arr[..., pb:pb + fk2, :] /= sT
    arr[..., it:it - fk2:-1, :] /= sT
    arr[..., :, pl:pl + fk2] /= s
    arr[..., :, ir:ir - fk2:-1] /= s


def compress_map(map_: np.ndarray, coef: int, dtype) -> np.ndarray:
    """Block.compress_map without the HDU wrapper (coadd.py:2124-2133)."""
    a_min, a_max = (0, 65535) if dtype == np.uint16 else (-32768, 32767)
    return np.clip(np.floor(coef * np.log10(np.clip(map_, 1e-32, None)) + 0.5), a_min, a_max).astype(dtype)


def build_output(maps: dict, cfg, n_inimage: int, is_final: bool = True, pad_sides: str = "", keep_inputs=None) -> dict:
    """The arrays Block.build_output_file puts into its HDUs, keyed by EXTNAME (coadd.py:2156-2303).

    maps: out_map (n_out, n_inframe, side, side), T_weightmap (n_out, n_inimage, n1P, n1P) and the quality maps
    UC_map / Sigma_map / kappa_map / Tsum_map / Neff_map (n_out, side, side), all float32; they are not modified."""
    fk = cfg.fade_kernel
    side = cfg.NsideP + 2 * fk
    m = {k: np.array(v, dtype=np.float32, copy=True) for k, v in maps.items()}
    names = {"U": "UC_map", "S": "Sigma_map", "K": "kappa_map", "T": "Tsum_map", "N": "Neff_map"}
    if is_final:
        recover(m["out_map"], fk)
        width = cfg.postage_pad * cfg.n2
        pads = tuple(width * (sd not in pad_sides) for sd in "BTLR")
        for letter in cfg.outmaps:
            if letter in names:
                recover(m[names[letter]], fk, pads)
    sl = slice(fk, side - fk)
    out = {"PRIMARY": m["out_map"][:, :, sl, sl].copy(), "INWEIGHT": m["T_weightmap"].copy()}
    n_out = m["T_weightmap"].shape[0]
    out["INWTFLAT"] = np.transpose(m["T_weightmap"], axes=(0, 2, 1, 3)).reshape(n_out * cfg.n1P, n_inimage * cfg.n1P)
    for letter in "USKTN":
        if letter in cfg.outmaps:
            ext, coef, dt = ENCODING[letter]
            out[ext] = compress_map(m[names[letter]][:, sl, sl], coef, dt)
            if keep_inputs is not None:  # the float32 maps that were encoded (tests: tie analysis of differing codes)
                keep_inputs[ext] = (m[names[letter]][:, sl, sl].copy(), coef)
    return out
