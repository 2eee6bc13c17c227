"""Developer probe: cost of a large layer CAPACITY (lmax) when the models themselves are shallow --
what a transdimensional run with BayHunter's default prior (layers 1..20 -> 21 rows) looks like."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayhunter_b200 as bh
from bayhunter_b200 import synthetic
c = synthetic.CONFIGS["joint5"]
B = 8192
rng = np.random.default_rng(0)
specs = []
for ref in c["refs"]:
    if ref == "prf":
        x = synthetic.rf_time_axis(c["rf"]); y = rng.normal(0, 0.02, x.size)
    else:
        x = c["periods"]; y = 3.5 + rng.normal(0, 0.1, x.size)
    specs.append(bh.TargetSpec(ref, x, y, cov="exp"))
noise = synthetic.draw_noise(B, c["refs"], seed=8)
dev = torch.device("cuda:0")
for nrows, lmax in ((6, 6), (6, 21), ((3, 9), 9), ((3, 9), 21)):
    rows, nlay = synthetic.draw_batch(B, nrows, seed=7, lmax=lmax)
    eng = bh.Engine(specs, B, lmax)
    eng.set(profile=1)
    tr, tn, tz = (torch.from_numpy(a).to(dev) for a in (rows, nlay, noise))
    best = None
    for r in range(3):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = eng.eval(tr, tn, tz); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    print(json.dumps(dict(nrows=nrows, lmax=lmax, total_ms=round(best, 3), kernels={k: round(v, 3) for k, v in eng.last_kernel_ms().items()},
                          rounds=eng.last_counters()[2:10], logL_sum=float(out[0].sum()))), flush=True)
    eng.close()
