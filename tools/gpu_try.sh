#!/bin/bash
python tools/quick_bench.py joint5 8192 concurrent=1 swd_f32_walk=1,concurrent=1 swd_f32_walk=1,concurrent=0 swd_f32_walk=1,swd_max_spec=16,concurrent=1 swd_f32_walk=1,swd_max_spec=4,concurrent=1 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if not l.startswith('{'): print(l.strip()[:300]); continue
    d=json.loads(l); print({k:v for k,v in d.items() if k.startswith('swd_')}, 'conc', d['concurrent'], 'total', d['total_ms'], {k:round(v,2) for k,v in d['kernels'].items()}, 'consumed', d['consumed'], 'eval64', d['evaluated'], 'rounds', d['rounds'], 'same', d['same_as_first'])
"
python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, bayhunter_b200 as bh
from bayhunter_b200 import synthetic
c = synthetic.CONFIGS["joint5"]; rng=np.random.default_rng(0)
specs=[bh.TargetSpec(r, c["periods"], 3.5+rng.normal(0,.1,30), cov="exp") for r in c["refs"][:4]]
B=2048
rows,nlay=synthetic.draw_batch(B,(2,12),seed=3); noise=synthetic.draw_noise(B,c["refs"][:4],seed=4)
eng=bh.Engine(specs,B,rows.shape[1])
a=eng.eval_host(rows,nlay,noise,want_synth=True); ca=eng.last_counters()
eng.set(swd_f32_walk=1)
b=eng.eval_host(rows,nlay,noise,want_synth=True); cb=eng.last_counters()
print("identical synth:", np.array_equal(a[3],b[3],equal_nan=True), "status:", np.array_equal(a[2],b[2]), "logL:", np.array_equal(a[0],b[0]))
print("counters off:", ca[:2], "on:", cb[:2], "fp32 evals, violations:", cb[18], cb[19])
PY
