import sys; sys.path.insert(0,'.')
import numpy as np, torch
from tools.quick_bench import make_engine
eng, rows, nlay, noise = make_engine("joint5", 8192)
dev=torch.device("cuda:0"); tr,tn,tz=(torch.from_numpy(a).to(dev) for a in (rows,nlay,noise))
def times(n=8):
    out=[]
    for r in range(n):
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); eng.eval(tr,tn,tz); e1.record(); torch.cuda.synchronize()
        out.append(round(e0.elapsed_time(e1),2))
    return out
eng.set(profile=0)
for name, kw in (("rule(22,22) gate25", dict(swd_spw_rp=0, swd_spw_lp=0, rf_gate_pct=25, concurrent=1)),
                 ("rule(22,22) gate0", dict(swd_spw_rp=0, swd_spw_lp=0, rf_gate_pct=0, concurrent=1)),
                 ("rule(22,22) serial", dict(swd_spw_rp=0, swd_spw_lp=0, rf_gate_pct=25, concurrent=0)),
                 ("(20,24) gate25", dict(swd_spw_rp=20, swd_spw_lp=24, rf_gate_pct=25, concurrent=1)),
                 ("(23,23) gate25", dict(swd_spw_rp=23, swd_spw_lp=23, rf_gate_pct=25, concurrent=1)),
                 ("(24,24) gate25", dict(swd_spw_rp=24, swd_spw_lp=24, rf_gate_pct=25, concurrent=1)),
                 ("(24,28) gate25", dict(swd_spw_rp=24, swd_spw_lp=28, rf_gate_pct=25, concurrent=1)),
                 ("(16,16) gate25", dict(swd_spw_rp=16, swd_spw_lp=16, rf_gate_pct=25, concurrent=1))):
    eng.set(**kw)
    print("%-22s" % name, times())
