import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import numpy as np, torch
import bench_configs as BC
from pyimcom_b200 import pyimcom_croutines as G
from pyimcom_b200 import lakernel as GL
from pyimcom_b200.coadd import GpuBlock
from pyimcom_b200.psfovl_host import PSFTables
from pyimcom_b200.synth import StampConfig, SynthBlock
spec = BC.CONFIGS["2_eigen_tests_block"]
cfg = StampConfig(**spec["cfg"])
blk = SynthBlock(cfg, n_image=spec["n_image"], seed=spec["seed"], psf_sigmas=spec["sig"], star=True)
tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
gb = GpuBlock(blk, tab).prepare()
ds, indata = gb.build_system(5)
A = ds.A[:ds.n,:ds.n].cpu().numpy()
w = np.linalg.eigvalsh(A)
print("n", ds.n, "eig", w[0], w[-1], "fro", np.linalg.norm(A), "rank>1e-10:", (w>1e-10).sum(), "rank>1e-13", (w>1e-13).sum())
t0=time.perf_counter(); lam, Vt, sweeps = GL.eigh_device(ds.A.clone(), ds.n); torch.cuda.synchronize(); t=time.perf_counter()-t0
lam = lam[:ds.n].cpu().numpy(); V = Vt[:ds.n,:ds.n].cpu().numpy().T
print("mult", os.environ.get("B200_EIGH_FLOOR_MULT"), "sweeps", sweeps, "time", t, "eigval err", np.abs(np.sort(lam)-w).max(), "resid", np.abs(A@V-V*lam).max(), "orth", np.abs(V.T@V-np.eye(ds.n)).max())
