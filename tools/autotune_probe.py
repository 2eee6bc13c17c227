"""Developer probe: models-per-warp rule vs autotuned choice across batch sizes (joint5 targets)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayhunter_b200 as bh
from bayhunter_b200 import synthetic
cfg = sys.argv[1] if len(sys.argv) > 1 else "joint5"
c = synthetic.CONFIGS[cfg]
rng = np.random.default_rng(0)
specs = []
for ref in c["refs"]:
    if ref == "prf":
        x = synthetic.rf_time_axis(c["rf"]); y = rng.normal(0, 0.02, x.size)
    else:
        x = c["periods"]; y = 3.5 + rng.normal(0, 0.1, x.size)
    specs.append(bh.TargetSpec(ref, x, y, cov="exp"))
dev = torch.device("cuda:0")
for B in [int(a) for a in sys.argv[2:]] or [1024, 2048, 4096, 6144, 7936, 8192, 12288, 16384]:
    rows, nlay = synthetic.draw_batch(B, c["nrows"], seed=7)
    noise = synthetic.draw_noise(B, c["refs"], seed=8)
    tr, tn, tz = (torch.from_numpy(a).to(dev) for a in (rows, nlay, noise))
    res = {}
    for auto in (0, 1):
        eng = bh.Engine(specs, B, rows.shape[1])
        eng.set(swd_autotune=auto)
        for r in range(14 if auto else 3):
            out = eng.eval(tr, tn, tz); torch.cuda.synchronize()
        best = 1e9
        for r in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); out = eng.eval(tr, tn, tz); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[auto] = (round(best, 3), eng.last_counters()[3:10:2], float(out[0].sum()))
        eng.close()
    print(json.dumps(dict(B=B, rule_ms=res[0][0], tuned_ms=res[1][0], rule_maxrounds=res[0][1], tuned_maxrounds=res[1][1],
                          same=res[0][2] == res[1][2])), flush=True)
