import sys, time, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench, util
from qtip_b200 import synth
wl = bench.workload(1, 0.1)
fa = synth.make_fasta(wl["records"], seed=wl["genome_seed"], lead_n=wl["lead_n"])
case = dict(fasta=fa, tu=wl["tu"], tb=[], tc=[], td=[])
hs = util.HostSide(case)
kw = dict(wl["kw"]); kw["chunksz"]=131072; kw["cg"]=(34,2619)
olap = int(hs.L.qthost_olap(hs.h))
t=time.time(); n=hs.L.qthost_windows(hs.h, 131072, olap); tw=time.time()-t
print('windows+ncount(host loop)', tw)
for i in range(2):
    t=time.time(); rc, cg, draws, n = hs.plan(kw); dt=time.time()-t-tw
    print('plan only approx s', dt, 'ns/read', dt/n[0]*1e9, 'ns/draw', dt/draws*1e9)
