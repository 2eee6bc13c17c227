#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch
from tools.quick_bench import make_engine
eng, rows, nlay, noise = make_engine("joint5", 8192)
dev=torch.device("cuda:0"); tr,tn,tz=(torch.from_numpy(a).to(dev) for a in (rows,nlay,noise))
def times(n=10):
    out=[]
    for r in range(n):
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); eng.eval(tr,tn,tz); e1.record(); torch.cuda.synchronize()
        out.append(round(e0.elapsed_time(e1),2))
    return out
eng.set(profile=0); print("default profile0", times())
eng.set(profile=1); print("default profile1", times())
eng.set(profile=0)
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(30): eng.eval(tr,tn,tz)
e1.record(); torch.cuda.synchronize(); print("back-to-back x30:", round(e0.elapsed_time(e1)/30,3))
import numpy as np
lg=np.empty(8192); mf=np.empty((8192,6)); st=np.empty(8192,dtype=np.int32)
import time
for k in range(3):
    t0=time.perf_counter()
    for r in range(20): eng.eval_host(rows,nlay,noise,out=(lg,mf,st,None))
    print("eval_host x20 (pageable numpy):", round((time.perf_counter()-t0)/20*1e3,3))
PY
python bench.py --steps 50 --warmup 3 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3), round(d['sampler']['ms_per_lockstep_iteration'],3))"
