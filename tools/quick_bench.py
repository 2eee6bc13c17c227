"""Developer probe: per-kernel timings for a config under several tunables (not the bench contract).

  python tools/quick_bench.py <cfg> <B> [key=value,key=value ...]...

Each further argument is one setting (comma separated engine tunables); with none, a default sweep
runs.  BH_B200_LIB selects the library build (see bayhunter_b200/_lib.py)."""
import sys, os, json
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
    ref = None
    for st in settings:
        try:
            eng.set(profile=1, **st)
        except Exception as ex:      # older library builds do not know every tunable
            print(json.dumps(dict(cfg=cfg, skipped=st, why=str(ex))), flush=True)
            continue
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
        ctr = eng.last_counters() if hasattr(eng, 'last_counters') else []
        logL = out[0].cpu().numpy()
        if ref is None:
            ref = logL
        same = bool(np.array_equal(ref, logL, equal_nan=True))
        print(json.dumps(dict(cfg=cfg, B=B, **st, total_ms=round(best[0], 3),
                              evals_per_s=round(B / best[0] * 1e3), kernels={a: round(b, 3) for a, b in best[1].items()},
                              consumed=cons, evaluated=ev, rounds=ctr[2:10], valid=float(out[2].float().mean()),
                              logL_sum=float(np.nansum(logL)), same_as_first=same)), flush=True)


def parse(arg):
    d = {}
    for kv in arg.split(","):
        k, v = kv.split("=")
        d[k] = int(v)
    return d


if __name__ == "__main__":
    cfg = sys.argv[1] if len(sys.argv) > 1 else "joint5"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else synthetic.CONFIGS[cfg]["B"]
    settings = [parse(a) for a in sys.argv[3:]]
    if not settings:
        settings = [dict(swd_searches_per_warp=s, swd_group_searches_per_warp=g, swd_max_spec=m, concurrent=c)
                    for (s, g, m, c) in ((32, 32, 8, 0), (32, 16, 8, 0), (16, 16, 8, 0), (16, 8, 8, 0), (8, 8, 8, 0),
                                         (32, 16, 8, 1), (16, 8, 8, 1), (16, 8, 4, 1), (16, 8, 16, 1))]
    print("# lib:", os.environ.get("BH_B200_LIB", "default"), flush=True)
    run(cfg, B, settings)
