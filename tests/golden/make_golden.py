"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (build container only).

    python tests/golden/make_golden.py

imports pyimcom.{routine,lakernel,psfutil,coadd} verbatim from /root/reference (oracle/refhost.py),
drives the reference's own InStamp -> OutStamp -> SysMatA/SysMatB -> PSFGrp/PSFOvl -> routine.py ->
{Chol single, Chol multi, Eigen single/multi, Iterative single/multi} path on the seeded synthetic
blocks of tests/cases.py, and stores what it produced.  /root/reference does not exist on the GPU
box, so the vectors are committed.  The inputs are NOT stored: they are regenerated from the seeds
(pyimcom_b200.synth is deterministic), only the reference's outputs are.
"""

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import refhost  # noqa: E402


def routine_vectors(ref):
    """Outputs of the reference's routine.py on the inputs of its own tests/pyimcom/test_routine.py."""
    r = ref.routine
    out = {}
    infunc, x_, y_, xs_, ys_, xpos, ypos = cases.interp_inputs()
    f1 = np.zeros((2, x_.size))
    r.iD5512C(infunc, x_, y_, f1)
    out["iD5512C"] = f1
    f2 = np.zeros((2, x_.size))
    r.iD5512C_sym(infunc, xs_, ys_, f2)
    out["iD5512C_sym"] = f2
    f3 = np.zeros((xpos.shape[0], xpos.shape[1] * ypos.shape[1]))
    r.gridD5512C(infunc[0], xpos, ypos, f3)
    out["gridD5512C"] = f3
    w = np.zeros(10)
    ws = []
    for fh in cases.GETW_FH:
        r.iD5512C_getw(w, fh)
        ws.append(w.copy())
    out["getw"] = np.array(ws)

    A, mBhalf, C = cases.kernel_toy()
    lam, Q = np.linalg.eigh(A)
    mPhalf = mBhalf @ Q
    m, n = mBhalf.shape
    kappa, Sigma, UC, T = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros((m, n))
    r.lakernel1(lam, Q, mPhalf, C, 1e-8, 1e-16, 1e16, 53, kappa, Sigma, UC, T, 0.5)
    out.update(lk1_kappa=kappa, lk1_Sigma=Sigma, lk1_UC=UC, lk1_T_sub=T[::25, ::33].copy(), lk1_Tabsmax=np.abs(T).max(),
               lk1_TQt_sub=(T @ Q.T)[::25, ::33].copy())
    A_ = A + np.identity(n)
    x = np.zeros(n)
    r.lsolve_sps(n, A_.copy(), x, mBhalf[0].copy())
    out["lsolve_x"] = x

    Nf, Df, Ef, kap, ucmin, smax = cases.reduced_inputs()
    mm = Df.size // kap.size
    ok, oS, oU, ow = np.zeros(mm), np.zeros(mm), np.zeros(mm), np.zeros(mm * kap.size)
    r.build_reduced_T_wrap(Nf, Df, Ef, kap, ucmin, smax, ok, oS, oU, ow)
    out.update(brt_kappa=ok, brt_Sigma=oS, brt_UC=oU, brt_w=ow)
    return out


def la_vectors(ref):
    """The reference's lakernel classes on the duck-typed inputs of tests/pyimcom/test_la.py."""
    out = {}
    for name, (kern, kappaC, extra) in cases.LA_CASES.items():
        outst = cases.la_outst(kappaC, **extra)
        K = getattr(ref.lakernel, kern)(outst)
        K()
        out[name + "_T"] = outst.T
        out[name + "_UC"] = outst.UC
        out[name + "_Sigma"] = outst.Sigma
        out[name + "_kappa"] = outst.kappa
    return out


def block_vectors():
    for name, spec in cases.BLOCK_CASES.items():
        blk = cases.make_block(spec)
        with contextlib.redirect_stdout(io.StringIO()):
            res = refhost.run_block(blk, spec["kernel"], spec["kappaC"], stamps=set(spec["stamps"]))
        out = {}
        for (j, i), d in res.items():
            tag = f"s{j}_{i}_"
            keep = ["outovlc", "T", "UC", "Sigma", "kappa", "outimage", "Tsum_stamp", "Tsum_inpix", "Neff", "inpix_cumsum"]
            if spec.get("store_ab", False):
                keep += ["sysmata", "mhalfb"]
            if "Ti64" in d and spec.get("store_ti64", False):
                keep += ["Ti64"]
            for k in keep:
                out[tag + k] = d[k]
        np.savez_compressed(os.path.join(HERE, f"block_{name}.npz"), **out)
        print("wrote", name, {k: v.shape for k, v in out.items() if k.endswith("_T")})


def output_vectors(ref):
    """Block.build_output_file's array work (coadd.py:2156-2303) by the reference's own OutStamp.trapezoid and
    Block.compress_map on the seeded block maps of cases.output_case."""
    trap, comp = ref.coadd.OutStamp.trapezoid, ref.coadd.Block.compress_map
    enc = {"U": ("UC_map", "FIDELITY", -5000, np.uint16), "S": ("Sigma_map", "SIGMA", -10000, np.int16),
           "K": ("kappa_map", "KAPPA", -5000, np.uint16), "T": ("Tsum_map", "INWTSUM", 200000, np.int16),
           "N": ("Neff_map", "EFFCOVER", 50000, np.uint16)}
    out = {}
    for name in cases.OUTPUT_CASES:
        cfg, maps, n_inimage, pad_sides, is_final = cases.output_case(name)
        fk = cfg.fade_kernel
        NsidePf = cfg.NsideP + fk * 2
        if is_final:  # coadd.py:2161-2176
            trap(maps["out_map"], fk, recover_mode=True)
            width = cfg.postage_pad * cfg.n2
            pads = tuple(width * (sd not in pad_sides) for sd in "BTLR")
            for letter in cfg.outmaps:
                trap(maps[enc[letter][0]], fk, True, pads)
        out[name + "_PRIMARY"] = maps["out_map"][:, :, fk:NsidePf - fk, fk:NsidePf - fk]
        out[name + "_INWTFLAT"] = np.transpose(maps["T_weightmap"], axes=(0, 2, 1, 3)).reshape(
            (cfg.n_out * cfg.n1P, n_inimage * cfg.n1P))  # coadd.py:2238-2241
        for letter in cfg.outmaps:
            mname, ext, coef, dt = enc[letter]
            out[name + "_" + ext] = comp(maps[mname][:, fk:NsidePf - fk, fk:NsidePf - fk], coef, dt)
    return out


def partition_vectors(ref):
    """InImage.partition_pixels + extract_layers (coadd.py:174-408) run by the reference itself on a bare InImage whose
    WCS composition, masks and layer reader are the seeded stand-ins of cases.partition_case."""
    import types

    co = ref.coadd
    out = {}
    saved = (co.Stn.sca_nside, co.get_all_data, co.Mask)
    try:
        for name in cases.PARTITION_CASES:
            cfg, outpix, mask_a, mask_b, use, indata, sca, sp_res = cases.partition_case(name)
            cfg.inlayercache = None
            co.Stn.sca_nside = sca
            co.get_all_data = lambda self, indata=indata: setattr(self, "indata", indata)
            co.Mask = types.SimpleNamespace(load_cr_mask=lambda self, m=mask_a: m.copy(),
                                            load_mask_from_maskfile=lambda cfg_, obs, idsca, m=mask_b: m.copy())
            im = object.__new__(co.InImage)
            im.blk = types.SimpleNamespace(cfg=cfg, use_instamps=use, pmask=None, timer=ref.config.Timer(), obsdata=None)
            im.idsca, im.exists_ = (1, 1), True
            im._inpix2world2outpix = outpix
            with contextlib.redirect_stdout(io.StringIO()):
                im.partition_pixels(sp_res=sp_res)
            out[name + "_is_relevant"] = np.array(im.is_relevant)
            if not im.is_relevant:
                continue
            for k in ("pix_count", "y_idx", "x_idx", "y_val", "x_val"):
                out[name + "_" + k] = getattr(im, k).copy()
            with contextlib.redirect_stdout(io.StringIO()):
                im.extract_layers()
            out[name + "_data"] = im.data
    finally:
        co.Stn.sca_nside, co.get_all_data, co.Mask = saved
    return out


if __name__ == "__main__":
    assert refhost.available(), "needs /root/reference (build container)"
    ref = refhost.load()
    np.savez_compressed(os.path.join(HERE, "routine.npz"), **routine_vectors(ref))
    np.savez_compressed(os.path.join(HERE, "la.npz"), **la_vectors(ref))
    np.savez_compressed(os.path.join(HERE, "output.npz"), **output_vectors(ref))
    np.savez_compressed(os.path.join(HERE, "partition.npz"), **partition_vectors(ref))
    block_vectors()
