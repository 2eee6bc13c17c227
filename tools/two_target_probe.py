import sys; sys.path.insert(0,'.')
import numpy as np, torch, json
import bayhunter_b200 as bh
from bayhunter_b200 import synthetic
rng=np.random.default_rng(0)
x=np.linspace(1,41,21); t=-5+0.2*np.arange(201)
specs=[bh.TargetSpec("rdispph",x,3.5+rng.normal(0,.1,21),cov="white"), bh.TargetSpec("prf",t,rng.normal(0,.02,201),cov="exp")]
B=8192
for lmax in (9, 21):
    rows,nlay=synthetic.draw_batch(B,(3,9),seed=1,lmax=lmax); noise=synthetic.draw_noise(B,["rdispph","prf"],seed=2)
    eng=bh.Engine(specs,B,lmax); eng.set(profile=1)
    dev=torch.device("cuda:0"); tr,tn,tz=(torch.from_numpy(a).to(dev) for a in (rows,nlay,noise))
    for auto in (0,1):
        eng.set(swd_autotune=auto)
        for r in range(14):
            out=eng.eval(tr,tn,tz); torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); out=eng.eval(tr,tn,tz); e1.record(); torch.cuda.synchronize()
        print("lmax",lmax,"autotune",auto,"total %.3f"%e0.elapsed_time(e1), {k:round(v,3) for k,v in eng.last_kernel_ms().items()}, eng.last_counters()[2:6])
