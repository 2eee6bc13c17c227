"""Developer probe: per-kernel timings for a config under several tunables (not the bench contract)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bayhunter_b200 as bh
from bayhunter_b200 import synthetic


def make_engine(cfg, B):
    c = synthetic.CONFIGS[cfg]
    rng = np.random.default_rng(0)
    specs = []
    for ref in c["refs"]:
        if ref in ("prf", "srf"):
            x = synthetic.rf_time_axis(c["rf"]); y = rng.normal(0, 0.02, x.size)
        else:
            x = c["periods"]; y = 3.5 + rng.normal(0, 0.1, x.size)
        specs.append(bh.TargetSpec(ref, x, y, cov="exp"))
    rows, nlay = synthetic.draw_batch(B, c["nrows"], seed=7)
    noise = synthetic.draw_noise(B, c["refs"], seed=8)
    eng = bh.Engine(specs, B, rows.shape[1])
    return eng, rows, nlay, noise


def run(cfg, B, settings, reps=3):
    eng, rows, nlay, noise = make_engine(cfg, B)
    dev = torch.device("cuda:0")
    tr, tn, tz = (torch.from_numpy(a).to(dev) for a in (rows, nlay, noise))
    for st in settings:
        eng.set(profile=1, **st)
        best = None
        for r in range(reps):
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            out = eng.eval(tr, tn, tz)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            k = eng.last_kernel_ms()
            if best is None or ms < best[0]:
                best = (ms, k)
        cons, ev = eng.last_counts()
        print(json.dumps(dict(cfg=cfg, B=B, **st, total_ms=round(best[0], 3),
                              evals_per_s=round(B / best[0] * 1e3), kernels={a: round(b, 3) for a, b in best[1].items()},
                              consumed=cons, evaluated=ev, valid=float(out[2].float().mean()))), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "one":      # one setting, for ncu: tools/quick_bench.py one <cfg> <B> <spw> <spec> <conc>
        cfg, B, spw, spec, conc = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
        run(cfg, B, [dict(swd_searches_per_warp=spw, swd_max_spec=spec, concurrent=conc)], reps=2)
    if which in ("all", "joint5"):
        run("joint5", 8192, [dict(swd_searches_per_warp=s, swd_max_spec=m, concurrent=c)
                             for (s, m, c) in ((32, 1, 0), (32, 8, 0), (16, 8, 0), (8, 8, 0), (4, 8, 0),
                                               (32, 8, 1), (16, 8, 1), (8, 8, 1), (16, 4, 1), (16, 16, 1))])
    if which in ("all", "swd2"):
        run("swd2", 4096, [dict(swd_searches_per_warp=s, swd_max_spec=m, concurrent=0)
                           for (s, m) in ((32, 1), (32, 8), (16, 8), (8, 8), (4, 8), (2, 8), (4, 4), (4, 16), (2, 16), (1, 32))])
    if which in ("all", "transd3"):
        run("transd3", 4096, [dict(swd_searches_per_warp=s, swd_max_spec=m, concurrent=c)
                              for (s, m, c) in ((32, 8, 0), (8, 8, 0), (8, 8, 1), (4, 8, 1))])
