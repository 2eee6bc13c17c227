import sys; sys.path.insert(0,'.')
import numpy as np, torch
import bayhunter_b200 as bh
from bayhunter_b200 import synthetic
from tools.quick_bench import make_engine
eng, rows, nlay, noise = make_engine("joint5", 8192)
dev=torch.device("cuda:0"); tr,tn,tz=(torch.from_numpy(a).to(dev) for a in (rows,nlay,noise))
def t(n=6, sync_each=True):
    best=1e9
    for r in range(n):
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); out=eng.eval(tr,tn,tz); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1))
    return round(best,3)
for prof in (1,0,1,0):
    for ov in (0,16):
        eng.set(profile=prof, swd_spw_rp=ov, swd_spw_lp=ov)
        print("profile",prof,"phase S", ov or "rule(22)", "ms", t())
# back-to-back without sync (like the sampler): 20 evals
eng.set(profile=0, swd_spw_rp=0, swd_spw_lp=0)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(20): out=eng.eval(tr,tn,tz)
e1.record(); torch.cuda.synchronize(); print("back-to-back profile 0 rule:", round(e0.elapsed_time(e1)/20,3))
