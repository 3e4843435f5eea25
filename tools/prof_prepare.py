import cProfile, pstats, sys, os, io, time
sys.path.insert(0, '/root/repo'); os.chdir('/root/repo')
import torch, bench
from pyimcom_b200 import pyimcom_croutines as G
from pyimcom_b200.coadd import GpuBlock
from pyimcom_b200.psfovl_host import PSFTables
blk = bench.make_block(0, n1=6)
tab = PSFTables(blk, G.iD5512C, G.gridD5512C, dedup=True)
for _ in range(3):
    g = GpuBlock(blk, tab); g.prepare(); g.run(); g.download()
torch.cuda.synchronize()
pr = cProfile.Profile()
ts=[]
for _ in range(3):
    t0=time.perf_counter(); pr.enable(); g = GpuBlock(blk, tab); g.prepare(); pr.disable(); torch.cuda.synchronize(); t1=time.perf_counter()
    g.run(); g.download(); ts.append(t1-t0)
print("prepare ms", [round(1e3*t,2) for t in ts])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(22); print(s.getvalue()[:4500])
