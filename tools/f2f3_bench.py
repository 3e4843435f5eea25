"""Rows f2 / f3 at production-like sizes: device time beside the CPU oracle (Python loops / NumPy) on the same inputs.

    python tools/f2f3_bench.py > profiles/f2f3_r01.txt
"""
import sys
import time
import warnings

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import output as OO  # noqa: E402
from oracle import partition as OP  # noqa: E402
from pyimcom_b200.coadd import assemble_output  # noqa: E402
from pyimcom_b200.partition import DevicePartition, to_host  # noqa: E402
from pyimcom_b200.synth import StampConfig  # noqa: E402


def dev_time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


# ---- f2: one 4088^2 input image against a block of 24 x 24 postage stamps of 32 output pixels (0.0390625") ----
cfg = StampConfig(n1=24, n2=32, postage_pad=0, dtheta_arcsec=0.0390625, fade_kernel=3)
rng = np.random.default_rng(11)
s, rot, ctr = 0.11 / 0.0390625, 0.4, np.array([2100.0, 1900.0])
cs, sn = s * np.cos(rot), s * np.sin(rot)
mid = cfg.NsideP / 2 - 0.5


def outpix(inxys):
    d = np.asarray(inxys, dtype=np.float64) - ctr
    return np.stack([cs * d[:, 0] - sn * d[:, 1] + 2e-6 * d[:, 0] * d[:, 1] + mid,
                     sn * d[:, 0] + cs * d[:, 1] - 1e-6 * d[:, 0] ** 2 + mid], axis=1)


ns = cfg.n1P + 2
use = np.ones((ns, ns), dtype=bool)
mask = rng.random((4088, 4088)) < 0.95
indata = rng.standard_normal((6, 4088, 4088)).astype(np.float32)
dp = DevicePartition(cfg, use)
d_mask = torch.from_numpy(mask.astype(np.uint8)).cuda()
d_in = torch.from_numpy(indata).cuda()
res = dp.partition(outpix, d_mask, d_in)
t_dev = dev_time(lambda: dp.partition(outpix, d_mask, d_in))
got = to_host(res)
t0 = time.perf_counter()
want = OP.partition_pixels(outpix, mask, cfg, use)
data = OP.extract_layers(indata, want, cfg)
t_cpu = time.perf_counter() - t0
same = all(np.array_equal(got[k], want[k]) for k in ("pix_count", "y_idx", "x_idx", "y_val", "x_val")) and \
    np.array_equal(got["data"], data)
print(f"f2 partition: {ns}x{ns} stamps, {res['n_cells']} relevant cells, {res['n_positions']} positions mapped, "
      f"{int(want['pix_count'].sum())} pixels kept, max {want['max_count']} per stamp (npixmax {dp.npixmax})")
print(f"   device (WCS stand-in on the host + H2D of positions + binning + extract_layers, mask and layers resident): "
      f"{1e3 * t_dev:.1f} ms;  CPU oracle (the reference's Python loops): {t_cpu:.2f} s;  identical: {same}")

# ---- f3: output assembly of a full paper-4 block (NsideP = 2688 incl. padding, 6 layers) ----
cfg3 = StampConfig(n1=80, n2=32, postage_pad=2, dtheta_arcsec=0.0390625, fade_kernel=3, n_inframe=6)
side = cfg3.NsideP + 2 * cfg3.fade_kernel
q = (1, side, side)
maps = {"out_map": rng.standard_normal((1, 6, side, side)), "T_weightmap": rng.uniform(0, 0.4, (1, 6, cfg3.n1P, cfg3.n1P)),
        "UC_map": 10.0 ** rng.uniform(-9, 0, q), "Sigma_map": 10.0 ** rng.uniform(-2, 1, q),
        "kappa_map": 10.0 ** rng.uniform(-9, -2, q), "Tsum_map": 1 + 0.05 * rng.standard_normal(q),
        "Neff_map": rng.uniform(0.5, 6, q)}
maps = {k: v.astype(np.float32) for k, v in maps.items()}
dmaps = {k: torch.from_numpy(v).cuda() for k, v in maps.items()}
t_dev3 = dev_time(lambda: assemble_output(dmaps, cfg3, 6, True, "", download=False))
t_dev3h = dev_time(lambda: assemble_output(dmaps, cfg3, 6, True, "", download=True))
got3 = assemble_output(dmaps, cfg3, 6, True, "")
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    t0 = time.perf_counter()
    want3 = OO.build_output(maps, cfg3, 6, True, "")
    t_cpu3 = time.perf_counter() - t0
exact = np.array_equal(got3["PRIMARY"], want3["PRIMARY"])
diff = {e: int((got3[e].astype(np.int64) != want3[e].astype(np.int64)).sum()) for e in
        ("FIDELITY", "SIGMA", "KAPPA", "INWTSUM", "EFFCOVER")}
dmax = max(int(np.abs(got3[e].astype(np.int64) - want3[e].astype(np.int64)).max()) for e in diff)
nbytes = 4 * 6 * side * side * 2 + 5 * side * side * (4 + 4 + 2)
print(f"f3 output assembly: side {side}, 6 layers + 5 quality maps: device {1e3 * t_dev3:.2f} ms on the device "
      f"({nbytes / t_dev3 / 1e9:.0f} GB/s of the algorithmic bytes), {1e3 * t_dev3h:.1f} ms with the download;  "
      f"NumPy: {1e3 * t_cpu3:.0f} ms;  PRIMARY identical: {exact};  codes differing (of {side * side} per map): {diff}, "
      f"max |difference| {dmax}")
