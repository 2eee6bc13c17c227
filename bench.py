#!/usr/bin/env python
"""bench.py -- joint SWD+RF forward + log-likelihood evaluations per second.

Contract (driver): `python bench.py --gpus N --steps K --warmup W`; for N > 1 it is
launched under torchrun (one rank per GPU).  Rank 0 prints ONE JSON line.

A step = one pass of the hot path over one batch of B chains per GPU: every chain's
proposed layered model goes through Rayleigh+Love phase+group dispersion (30
periods), the 512-sample P receiver function and the correlated-noise Gaussian
log-likelihood (BASELINE.json configs[2], the joint configuration the metric is
quoted on; weak scaling: B per GPU fixed).  Inputs are synthetic (seeded model
draws); each step evaluates a different, perturbed batch.

  value     device-resident inputs, CUDA-event time per step, L2 flushed between
            steps (not timed), max over ranks
  e2e       the same metric through the host-buffer C ABI (bh_engine_eval_host_async /
            bh_engine_wait, two calls in flight): host inputs -> H2D -> kernels -> D2H
            of logL / misfits / status every step, wall clock; sub-keys give the
            blocking call and pageable (plain numpy) caller memory
  roofline  dominant kernel, timed live with CUDA events inside the engine, against
            the resource that binds it: the fp64 pipe (algorithmic flops of SURVEY 8d
            over the DFMA rate measured on this device at start-up); `hbm` beside it
            with SURVEY 8d's algorithmic bytes (328 B per evaluation in config 3)
  parity    SURVEY 8d's error metrics: the GPU outputs of the cpu_baseline sample
            against the oracle outputs of the same models
  configs   the other BASELINE configurations (swd2, transd3, joint5 with the Gauss
            law on the RF target), 20 steps each, N = 1 only
  cpu_baseline  the oracle (reference rfmini C++ when compiled, SURF96 C restatement,
            numpy likelihood) on the host cores, bounded sample, pool started before
            the clock

`--impl reference` times that CPU path alone (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "joint_swd_rf_loglik_evals_per_sec"
UNIT = "evals/s"


# --------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------
def workload(cfg_name, B, seed, gauss_rf=False):
    """Targets (ref, x, y, cov, extra) and the seeded model batch of one rank."""
    from bayhunter_b200 import synthetic, gauss_corr_inverse
    c = synthetic.CONFIGS[cfg_name]
    rng = np.random.default_rng(seed + 1000)
    targets = []
    for ref in c["refs"]:
        extra = {}
        cov = "exp"
        if ref in ("prf", "srf"):
            x = synthetic.rf_time_axis(c["rf"])
            y = rng.normal(0.0, 0.02, x.size)         # observed trace: synthetic noise-like data
            if gauss_rf:
                # SURVEY 8d second run: r_rf = 0.9 fixed, Gauss law, host-computed R^-1 with rcond = 1e-5
                ci, ld = gauss_corr_inverse(0.9, x.size, rcond=1e-5)
                cov, extra = "gauss", dict(corr_inv=ci, logcorr_det=ld)
        else:
            x = np.asarray(c["periods"], dtype=np.float64)
            y = 3.0 + 0.02 * x + rng.normal(0.0, 0.02, x.size)
        targets.append((ref, x, y, cov, extra))
    rows, nlay = synthetic.draw_batch(B, c["nrows"], seed=seed)
    noise = synthetic.draw_noise(B, c["refs"], seed=seed + 1)
    if gauss_rf:
        for t, ref in enumerate(c["refs"]):
            if ref in ("prf", "srf"):
                noise[:, 2 * t] = 0.9
    return c, targets, rows, nlay, noise


def algorithmic_model(c, nlay, counts_consumed):
    """algorithmic bytes and flops of one step (B evaluations)."""
    T = len(c["refs"])
    B = nlay.size
    L = nlay.astype(np.float64)
    nbytes = float(np.sum(8 * (4 * L + 2 * T) + 8 * (T + 2)))
    flops = 0.0
    nsw = sum(1 for r in c["refs"] if r not in ("prf", "srf"))
    if nsw:
        # counted secular evaluations; split between Rayleigh/Love by curve count
        nr = sum(1 for r in c["refs"] if r.startswith("r") and r not in ("prf", "srf"))
        nl = nsw - nr
        per_eval = (nr * (175.0 * (L.mean() - 1) + 30.0) + nl * (28.0 * (L.mean() - 1) + 10.0)) / nsw
        flops += counts_consumed * per_eval
    if c["rf"] is not None:
        n = c["rf"]["n"]
        N = 2 ** int(np.ceil(np.log2(2 * n)))
        flops += float(np.sum((N / 2 + 1) * (470.0 * (L - 1) + 100.0) + 2.5 * N * np.log2(N)))
        flops += 6.0 * n * B
    flops += 6.0 * sum(len(c["periods"]) for r in c["refs"] if r not in ("prf", "srf")) * B
    return nbytes, flops


# --------------------------------------------------------------------------------------
# CPU arm (oracle) -- the only place bench.py executes oracle/
# --------------------------------------------------------------------------------------
def _cpu_init():
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"   # tutorialhunt.py:12-14
    from oracle import joint_oracle as jo
    jo.lib()


def _cpu_worker(args):
    targets, rows, nlay, noise, want_out = args
    from oracle import joint_oracle as jo
    ot = [jo.OracleTarget(ref, x, y, cov=cov, **extra) for ref, x, y, cov, extra in targets]
    t0 = time.perf_counter()
    out = jo.evaluate_batch(ot, rows, nlay, noise)
    dt = time.perf_counter() - t0
    return (dt, out) if want_out else (dt, None)


class CpuArm(object):
    """The CPU path with the reference's own parallel model: one single-threaded process per core
    (src/mcmcOptimizer.py:219-252).  The worker pool is started -- and every worker has loaded the
    oracle libraries -- before anything is timed."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        from oracle import joint_oracle as jo
        jo.lib()
        self.jo = jo
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init)
        self.pool.map(abs, range(4 * self.cores))            # every worker has started

    def close(self):
        self.pool.close()
        self.pool.join()

    def run(self, targets, rows, nlay, noise, per_core, want_out=False):
        cores = self.cores
        n = min(rows.shape[0], per_core * cores)
        per = max(1, n // cores)
        jobs = [(targets, rows[i * per:(i + 1) * per], nlay[i * per:(i + 1) * per], noise[i * per:(i + 1) * per], want_out)
                for i in range(cores)]
        jobs = [j for j in jobs if j[1].shape[0] > 0]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        nev = sum(j[1].shape[0] for j in jobs)
        jo = self.jo
        # SURF96: no compiled reference exists (no Fortran compiler in the build image nor on the GPU box, and the
        # reference sources do not travel); RF uses oracle/_ref (the reference's own C++) when it was built
        out = None
        if want_out:
            out = tuple(np.concatenate([r[1][k] for r in res]) for k in range(4))
        return dict(value=nev / wall, unit=UNIT, cores=len(jobs), kind="port",
                    sample="%d models (%d per core) of the same batch; RF via %s, SWD via C restatement of "
                           "surfdisp96.f, likelihood dense numpy as Targets.py; worker pool started before the clock" %
                           (nev, per, "oracle/_ref (reference rfmini C++)" if jo.ref_rfmini() is not None
                            else "C restatement of rfmini"),
                    seconds=wall, n=nev), out


# --------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []          # (time, line)
        self.proc = None

    def start(self, wait_s=4.0):
        """Start `nvidia-smi -lms` and return once its first sample has arrived (it takes ~1 s to come up)."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < wait_s:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, windows):
        """windows: [(t0, t1)] of the timed regions; samples inside them (else the nearest ones) are summarised."""
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"], samples=0)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        parsed = []
        for ts, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                parsed.append((ts, float(f[1]), float(f[2]), f[3], f[5:9]))
            except ValueError:
                continue
        inside = [p for p in parsed if any(a <= p[0] <= b for a, b in windows)]
        where = "inside the timed regions"
        if not inside and parsed and windows:
            mid = 0.5 * (windows[0][0] + windows[-1][1])
            inside = sorted(parsed, key=lambda p: abs(p[0] - mid))[:3]
            where = "nearest to the timed regions (they are shorter than nvidia-smi's sampling period)"
        sm = [p[1] for p in inside]
        mx = [p[2] for p in inside]
        reasons = set()
        for p in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = []
        for p in inside:
            try:
                pw.append(float(p[3]))
            except ValueError:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), samples_total=len(parsed), window=where,
                    power_w=float(np.median(pw)) if pw else None)


# --------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    from bayhunter_b200 import synthetic
    arm = CpuArm()
    per_core = args.ref_per_core
    c, targets, rows, nlay, noise = workload(cfg, per_core * arm.cores, seed=20260101)
    times = []
    last = None
    for s in range(args.warmup + args.steps):
        rng = np.random.default_rng(5000 + s)
        r = synthetic.perturb_batch(rows, nlay, rng)
        last, _ = arm.run(targets, r, nlay, noise, per_core)
        if s >= args.warmup:
            times.append(last["seconds"])
    arm.close()
    nev = last["n"]
    value = nev * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same workload (and the same `config` object) as the b200 arm; each step evaluates a bounded sample of it
        "config": config_object(cfg, args.chains_per_gpu or synthetic.CONFIGS[cfg]["B"], c, args.gpus),
        "sample_models_per_step": nev,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_object(cfg, B, c, world):
    """`config` of the JSON line, identical in both arms."""
    return {"workload": cfg_description(cfg, B, c), "global_chains": world * B,
            "l2": "flushed between timed steps (256 MiB write, outside the event pairs)",
            "parallelism": "chains sharded, dp%d" % world}


def cfg_description(cfg, B, c, gauss_rf=False):
    rf = "" if c["rf"] is None else " + P-RF %d samples (nsamp %d%s)" % (
        c["rf"]["n"], 2 ** int(np.ceil(np.log2(2 * c["rf"]["n"]))), ", Gauss law r = 0.9" if gauss_rf else "")
    nr = c["nrows"]
    rows = "%d rows (%d layers + half-space)" % (nr, nr - 1) if np.isscalar(nr) else "%d-%d rows" % nr
    return "%s: %s, %d periods%s, %s, %d chains per GPU" % (cfg, "+".join(r for r in c["refs"] if r not in ("prf", "srf")),
                                                       len(c["periods"]), rf, rows, B)


# --------------------------------------------------------------------------------------
# SURVEY 8d error metrics
# --------------------------------------------------------------------------------------
def parity_metrics(targets, gpu, ora):
    """gpu / ora: (logL, misfits, status, synth) of the same models."""
    g_logL, _, g_stat, g_syn = gpu
    o_logL, _, o_stat, o_syn = ora
    ok = (o_stat == 1) & (g_stat == 1)
    res = {"n": int(o_stat.size), "n_valid": int(ok.sum()), "flags_equal": bool(np.array_equal(g_stat, o_stat))}
    o = 0
    ph, gr, rf = [], [], []
    for ref, x, _, _, _ in targets:
        n = x.size
        a, b = g_syn[ok, o:o + n], o_syn[ok, o:o + n]
        o += n
        if not ok.any():
            continue
        if ref in ("prf", "srf"):
            rf.append(np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1))
        elif ref.endswith("ph"):
            ph.append((np.abs(a - b) / np.abs(b)).ravel())
        else:
            gr.append((np.abs(a - b) / np.abs(b)).ravel())
    if ph:
        res["swd_phase_max"] = float(np.concatenate(ph).max())
    if gr:
        g = np.concatenate(gr)
        res["swd_group_max"] = float(g.max())
        res["swd_group_frac_le_1e-6"] = float((g <= 1e-6).mean())
        res["swd_group_frac_le_2e-5"] = float((g <= 2e-5).mean())
    if rf:
        res["rf_max"] = float(np.concatenate(rf).max())
    if ok.any():
        e = np.abs(g_logL[ok] - o_logL[ok]) / np.maximum(1.0, np.abs(o_logL[ok]))
        res["logL_max"] = float(e.max())
        res["logL_median"] = float(np.median(e))
        res["logL_frac_le_1e-6"] = float((e <= 1e-6).mean())
    res["what"] = ("GPU vs CPU oracle on the cpu_baseline sample; SURVEY 8d metrics: SWD max |c_gpu - c_ref| / c_ref per "
                   "sample, RF max |y_gpu - y_ref| / max |y_ref| per trace, logL |d logL| / max(1, |logL_ref|); group-velocity "
                   "samples inherit SURF96's REAL*4 finite difference, which amplifies 1e-9 root differences (SURVEY D.1-7)")
    return res


# --------------------------------------------------------------------------------------
# main arm
# --------------------------------------------------------------------------------------
def make_specs(bh, targets):
    return [bh.TargetSpec(ref, x, y, cov=cov, **extra) for ref, x, y, cov, extra in targets]


def time_config(bh, torch, dev, cfg, B, seed, steps, warmup, gauss_rf=False):
    """ms per step of another BASELINE configuration, device resident, L2 flushed between steps."""
    from bayhunter_b200 import synthetic
    c, targets, rows0, nlay, noise = workload(cfg, B, seed, gauss_rf=gauss_rf)
    eng = bh.Engine(make_specs(bh, targets), B, rows0.shape[1])
    eng.set(profile=1)
    d_nlay = torch.from_numpy(nlay).to(dev)
    d_noise = torch.from_numpy(noise).to(dev)
    d_b = [torch.from_numpy(synthetic.perturb_batch(rows0, nlay, np.random.default_rng(seed * 31 + s))).to(dev)
           for s in range(warmup + steps)]
    T = len(targets)
    out = (torch.empty(B, dtype=torch.float64, device=dev), torch.empty((B, T + 1), dtype=torch.float64, device=dev),
           torch.empty(B, dtype=torch.int32, device=dev), None)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    ms, km = [], {}
    for s in range(warmup + steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        eng.eval(d_b[s], d_nlay, d_noise, out=out)
        e1.record(st)
        k = eng.last_kernel_ms()
        flush.fill_(s & 0xFF)
        torch.cuda.synchronize(dev)
        if s >= warmup:
            ms.append(e0.elapsed_time(e1))
            for a, v in k.items():
                km.setdefault(a, []).append(v)
    valid = float(out[2].float().mean().item())
    eng.close()
    m = float(np.mean(ms))
    return {"workload": cfg_description(cfg, B, c, gauss_rf), "steps": steps, "ms_per_step": m,
            "evals_per_s": B / (m * 1e-3), "kernel_ms": {a: float(np.mean(v)) for a, v in km.items()},
            "valid_fraction": valid, "mean_rows": float(nlay.mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="joint5", choices=["swd2", "joint5", "transd3"])
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--ref-per-core", type=int, default=256)
    ap.add_argument("--cpu-per-core", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the 'configs' object (other BASELINE configurations)")
    ap.add_argument("--pool-samples", type=int, default=100)
    ap.add_argument("--sampler-iters", type=int, default=40,
                    help="lock-step MCMC iterations timed for the auxiliary 'sampler' object (0: skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the CPU pool forks first, before this process holds a CUDA context, threads or pinned pages; it idles until the end
    arm = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm()

    import torch
    import torch.distributed as dist
    import bayhunter_b200 as bh
    from bayhunter_b200 import synthetic, chains, _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bh._lib.require_device()
    bh._lib.set_device(local_rank)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # fp64 DFMA peak of this device, measured now (independent DFMA chains on every SM)
    tf, mhz = ctypes.c_double(0), ctypes.c_double(0)
    _lib.check(_lib.load().bh_measure_fp64_peak(ctypes.byref(tf), ctypes.byref(mhz)))
    fp64_peak = tf.value

    cfg = args.config
    B = args.chains_per_gpu or synthetic.CONFIGS[cfg]["B"]
    seed = chains.chain_seed(20260101, rank)
    c, targets, rows0, nlay, noise = workload(cfg, B, seed)
    specs = make_specs(bh, targets)
    L = rows0.shape[1]
    T = len(specs)
    eng = bh.Engine(specs, B, L)

    nsteps = args.warmup + args.steps
    # one perturbed batch per step, resident in HBM before timing starts
    batches = []
    for s in range(nsteps):
        rng = np.random.default_rng(seed * 7919 + s)
        batches.append(synthetic.perturb_batch(rows0, nlay, rng))
    d_batches = [torch.from_numpy(b).to(dev) for b in batches]
    d_nlay = torch.from_numpy(nlay).to(dev)
    d_noise = torch.from_numpy(noise).to(dev)
    out = (torch.empty(B, dtype=torch.float64, device=dev),
           torch.empty((B, T + 1), dtype=torch.float64, device=dev),
           torch.empty(B, dtype=torch.int32, device=dev), None)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    stream = torch.cuda.current_stream(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    eng.set(profile=1)
    if os.environ.get("BH_GATE"):        # developer sweep of the RF gate (tools/gpu_try.sh)
        eng.set(rf_gate_pct=int(os.environ["BH_GATE"]))
    kernel_ms = {}
    counts = [0, 0]
    windows = []
    for s in range(args.warmup):
        eng.eval(d_batches[s], d_nlay, d_noise, out=out)
        flush.fill_(s & 0xFF)
    barrier()
    t_wall0 = time.perf_counter()
    for s in range(args.warmup, nsteps):
        ev[s][0].record(stream)
        eng.eval(d_batches[s], d_nlay, d_noise, out=out)
        ev[s][1].record(stream)
        km = eng.last_kernel_ms()           # waits for this step's kernels (event sync)
        for k, v in km.items():
            kernel_ms.setdefault(k, []).append(v)
        cc = eng.last_counts()
        counts[0] += cc[0]; counts[1] += cc[1]
        flush.fill_(s & 0xFF)               # L2 flush between timed steps (outside the events)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    windows.append((t_wall0, t_wall0 + t_wall))
    step_ms = [ev[s][0].elapsed_time(ev[s][1]) for s in range(args.warmup, nsteps)]
    total_ms = float(sum(step_ms))
    valid_frac = float(out[2].float().mean().item())
    dev_logL = out[0].cpu().numpy()

    # ---- a longer run of the same step when the driver's K is short: sustained clocks ----
    sustained = None
    if args.steps < 100 and rank == 0 and world == 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ns = 200
        ts0 = time.perf_counter()
        e0.record(stream)
        for s in range(ns):
            eng.eval(d_batches[args.warmup + s % args.steps], d_nlay, d_noise, out=out)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        windows.append((ts0, time.perf_counter()))
        sustained = {"steps": ns, "ms_per_step": e0.elapsed_time(e1) / ns,
                     "value": B * ns / (e0.elapsed_time(e1) * 1e-3),
                     "what": "200 back-to-back steps (no L2 flush, one event pair): the rate under sustained clocks"}
        eng.eval(d_batches[nsteps - 1], d_nlay, d_noise, out=out)      # leave the last batch's results in `out`
        torch.cuda.synchronize(dev)

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    eng.set(profile=0)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_batches = [pin(b) for b in batches]
    h_nlay, h_noise = pin(nlay), pin(noise)

    def host_outs(pinned):
        if pinned:
            return (torch.empty(B, dtype=torch.float64).pin_memory().numpy(),
                    torch.empty((B, T + 1), dtype=torch.float64).pin_memory().numpy(),
                    torch.empty(B, dtype=torch.int32).pin_memory().numpy(), None)
        return (np.empty(B), np.empty((B, T + 1)), np.empty(B, dtype=np.int32), None)

    def run_host(pinned, pipelined):
        ins = [b.numpy() for b in h_batches] if pinned else batches
        nl, nz = (h_nlay.numpy(), h_noise.numpy()) if pinned else (nlay, noise)
        outs = [host_outs(pinned), host_outs(pinned)]
        for s in range(args.warmup):
            eng.eval_host(ins[s], nl, nz, out=outs[0])
        barrier()
        t0 = time.perf_counter()
        if pipelined:
            prev = None
            for s in range(args.warmup, nsteps):
                tk = eng.submit_host(ins[s], nl, nz, outs[s & 1])
                if prev is not None:
                    eng.wait(prev)
                prev = tk
            eng.wait(prev)
        else:
            for s in range(args.warmup, nsteps):
                eng.eval_host(ins[s], nl, nz, out=outs[s & 1])
        dt = time.perf_counter() - t0
        windows.append((t0, t0 + dt))
        return dt, outs[(nsteps - 1) & 1][0].copy()

    e2e_s, last_logL = run_host(pinned=True, pipelined=True)
    e2e_block_s, _ = run_host(pinned=True, pipelined=False)
    e2e_page_s, page_logL = run_host(pinned=False, pipelined=True)
    e2e_page_block_s, _ = run_host(pinned=False, pipelined=False)
    h2d = batches[0].nbytes + nlay.nbytes + noise.nbytes
    d2h = B * 8 + B * (T + 1) * 8 + B * 4
    # device and host paths must agree bit for bit on the last batch
    same = bool(np.array_equal(last_logL, dev_logL, equal_nan=True) and np.array_equal(page_logL, dev_logL, equal_nan=True))

    # ---- posterior pooling: the single collective of the path (not part of a step) ----
    pool_ms = None
    if world > 1:
        blk = chains.PosteriorBlock(B, args.pool_samples, L, T, device=dev)
        blk.record(0, d_batches[-1], d_nlay, out[0], out[1], d_noise)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pooled = chains.pool_posterior(blk)
        e1.record()
        torch.cuda.synchronize(dev)
        pool_ms = e0.elapsed_time(e1)
        assert pooled["likes"].shape[0] == world * B
        # values, not only shapes: every rank's block must arrive where its chains belong
        mine = pooled["likes"][rank * B:(rank + 1) * B]
        assert torch.equal(torch.nan_to_num(mine), torch.nan_to_num(blk.likes)), "pooled posterior differs from the local block"

    clocks = sampler.stop(windows) if rank == 0 else None

    # ---- auxiliary: the device sampler around the same hot path ----
    # B chains advance in lock step (propose kernel -> engine -> accept kernel, no host round trip);
    # reported beside the headline metric, not part of it.
    smp = None
    if args.sampler_iters > 0:
        from bayhunter_b200 import Targets as T_, SingleChain as sc
        cls = {"rdispph": T_.RayleighDispersionPhase, "rdispgr": T_.RayleighDispersionGroup,
               "ldispph": T_.LoveDispersionPhase, "ldispgr": T_.LoveDispersionGroup,
               "prf": T_.PReceiverFunction, "srf": T_.SReceiverFunction}
        jt = T_.JointTarget([cls[ref](x, y) for ref, x, y, _, _ in targets])
        nr = c["nrows"]
        lay = (nr - 1, nr - 1 + 3) if np.isscalar(nr) else (nr[0] - 1, nr[1] - 1)
        priors = dict(vs=(2, 5), z=(0, 60), layers=lay, vpvs=(1.4, 2.1), swdnoise_corr=0.,
                      swdnoise_sigma=(1e-5, 0.1), rfnoise_corr=(0.35, 0.75), rfnoise_sigma=(1e-5, 0.05))
        ip = dict(iter_burnin=100000, iter_main=50000, thickmin=0.1, acceptance=(40, 45))
        lo = rank * B
        ens = sc.ChainEnsemble(jt, priors, ip, nchains=B, first_chain=lo, seed=20260101, max_accepted=64,
                               chain_seeds=np.arange(lo, lo + B) + 7)
        ens.init()
        ens.run(12)          # includes the engine's models-per-warp autotuning (9 timed evaluations)
        barrier()
        t0 = time.perf_counter()
        ens.run(args.sampler_iters)
        torch.cuda.synchronize(dev)
        smp_s = time.perf_counter() - t0
        st = ens.state()
        smp = dict(seconds=smp_s, proposed=float(st["proposed"].sum()), accepted=float(st["accepted"].sum()),
                   iters=args.sampler_iters + 12, mean_rows=float(st["k"].mean()), layers_prior=list(lay))
        ens.close()

    # ---- max over ranks ----
    if world > 1:
        if smp is not None:
            tt = torch.tensor([smp["seconds"]], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            smp["seconds"] = float(tt[0])
        t = torch.tensor([total_ms, e2e_s, e2e_block_s, e2e_page_s, e2e_page_block_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, e2e_block_s, e2e_page_s, e2e_page_block_s = (float(v) for v in t)
        cnt = torch.tensor(counts, dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        K = args.steps
        value = world * B * K / (total_ms * 1e-3)
        rate = lambda sec: world * B * K / sec
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        kmean = {k: float(np.mean(v)) for k, v in kernel_ms.items()}
        dom = max(kmean, key=kmean.get)
        nbytes, flops = algorithmic_model(c, nlay, counts[0] / K)
        # flops of the dominant kernel alone (175 flop per Rayleigh layer evaluation + 30, 28 + 10 Love,
        # times the COUNTED secular evaluations)
        dom_name, dom_parts = dom + "_kernel", None
        if dom in ("swd", "swd_love", "swd_pool", "swd_pool_love"):
            nsw = sum(1 for r in c["refs"] if r not in ("prf", "srf"))
            nr = sum(1 for r in c["refs"] if r.startswith("r") and r not in ("prf", "srf"))
            Lm = float(nlay.mean())
            fl_r = counts[0] / K * nr * (175.0 * (Lm - 1) + 30.0) / nsw
            fl_l = counts[0] / K * (nsw - nr) * (28.0 * (Lm - 1) + 10.0) / nsw
            dom_flops = fl_r + fl_l
        else:
            dom_flops = flops
        dom_s = kmean[dom] * 1e-3
        if dom in ("swd_pool", "swd_pool_love") and "swd_pool" in kmean and "swd_pool_love" in kmean:
            # full batches: the dispersion search runs as two launches of swd_pool_kernel (Rayleigh, Love) that start
            # together on two streams and share every SM; the pair is the dominant "kernel", its duration the longer one
            dom = "swd_pool"
            dom_name = "swd_pool_kernel<Rayleigh> + swd_pool_kernel<Love> (concurrent launches sharing the SMs; duration = the longer)"
            dom_s = max(kmean["swd_pool"], kmean["swd_pool_love"]) * 1e-3
            dom_parts = {"swd_pool_kernel<2> (Rayleigh)": {"ms": kmean["swd_pool"], "algorithmic_flops": fl_r},
                         "swd_pool_kernel<1> (Love)": {"ms": kmean["swd_pool_love"], "algorithmic_flops": fl_l}}
        traffic = None
        pipe = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(cfg, {}).get(dom + "_kernel")
            pp = tr.get("fp64_pipe", {}).get(cfg)
            if pp and dom == pp.get("kernel", "swd"):
                # the fp64 pipe of a sub-partition takes one warp instruction per 2 cycles: instructions of the
                # profiled launch (same workload) against the pipe cycles of the LIVE launch duration and clock
                mhz_live = float((clocks or {}).get("sm_mhz") or mhz.value or 1965.0)
                pipe = {"busy_ncu": pp["busy_pct_ncu"] / 100.0,
                        "busy_live": pp["fp64_warp_instructions"] * 2.0 / (148 * 4 * mhz_live * 1e6 * dom_s),
                        "fp64_warp_instructions_per_launch": pp["fp64_warp_instructions"],
                        "source": "profiles/traffic.json (ncu --set full capture of the same workload); busy_live = "
                                  "fp64 warp instructions x 2 cycles / (592 sub-partitions x SM clock x live kernel time)"}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_object(cfg, B, c, world),
            "valid_fraction": valid_frac,
            "e2e": {"value": rate(e2e_s), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "bh_engine_eval_host_async + bh_engine_wait, two calls in flight, pinned host buffers "
                           "(every step: H2D of its inputs, kernels, D2H of logL / misfits / status)",
                    "blocking_pinned": rate(e2e_block_s), "pageable_pipelined": rate(e2e_page_s),
                    "pageable_blocking": rate(e2e_page_block_s),
                    "pageable_over_device_resident": rate(e2e_page_s) / value,
                    "matches_device_path": same},
            # per step: prepare(SWD rows), layer_order, swd_kernel (or the two swd_pool_kernel launches of a full batch),
            # loglik (+ swd_gate, prepare(RF tables), rf_spectrum, rf_synth with an RF target)
            "gpu_launches": ((8 if c["rf"] is not None else 4) + (1 if "swd_pool_love" in kmean else 0)) * K,
            "roofline": {"bound": "fp64", "kernel": dom_name, "achieved": dom_flops / dom_s / 1e12,
                         "peak": fp64_peak, "unit": "TFLOP/s", "frac": dom_flops / dom_s / 1e12 / fp64_peak,
                         "traffic": traffic,
                         "peak_source": "measured on this device at start-up (bh_measure_fp64_peak: independent DFMA chains "
                                        "on every SM, best of 3)",
                         "algorithmic_flops_per_launch": dom_flops,
                         "flop_model": "SURVEY 8d: 175 flop per Rayleigh layer evaluation + 30, 28 + 10 Love, x counted "
                                       "secular evaluations",
                         "secular_evals_per_step": counts[0] / K, "secular_evaluated_per_step": counts[1] / K,
                         "step_flops": flops, "step_tflops": flops / (total_ms / K * 1e-3) / 1e12,
                         "pipe": pipe, "launches_of_the_pair": dom_parts,
                         "hbm": {"achieved": nbytes / (total_ms / K * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": nbytes / (total_ms / K * 1e-3) / 1e9 / hbm_peak,
                                 "algorithmic_bytes_per_step": nbytes, "bytes_per_eval": nbytes / B,
                                 "peak_source": peak_src,
                                 "note": "SURVEY 8d bytes over the whole step: HBM is not what bounds this path"}},
            "kernel_ms": kmean,
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
        if sustained is not None:
            line["sustained"] = sustained
        if pool_ms is not None:
            line["pool_allgather_ms"] = pool_ms
        if smp is not None:
            line["sampler"] = {
                "chain_iterations_per_s": world * B * args.sampler_iters / smp["seconds"],
                "ms_per_lockstep_iteration": 1e3 * smp["seconds"] / args.sampler_iters,
                "evaluated_fraction": smp["proposed"] / (B * smp["iters"]),
                "accept_rate": smp["accepted"] / max(1.0, smp["proposed"]),
                "mean_rows": smp["mean_rows"], "layers_prior": smp["layers_prior"],
                "what": "bh_sampler_run: B chains per GPU in lock step on the device (proposal, prior check, "
                        "forward models, likelihood, Metropolis-Hastings, proposal-width control); rank-0 "
                        "statistics, wall clock max over ranks"}
        if world == 1 and not args.no_configs:
            others = {}
            for name, (ocfg, oB, g) in {"swd2": ("swd2", 4096, False), "transd3": ("transd3", 4096, False),
                                        "joint5_gauss": ("joint5", 8192, True)}.items():
                try:
                    others[name] = time_config(bh, torch, dev, ocfg, oB, seed + 17, 20, 3, gauss_rf=g)
                except Exception as ex:         # a failing side measurement must not lose the headline
                    others[name] = {"error": repr(ex)}
            if "joint5_gauss" in others and "ms_per_step" in others["joint5_gauss"] and cfg == "joint5":
                others["joint5_gauss"]["over_exp_law_step"] = others["joint5_gauss"]["ms_per_step"] / (total_ms / K)
            line["configs"] = others
        if arm is not None:
            n_cpu = min(B, args.cpu_per_core * arm.cores)
            cb, ora = arm.run(targets, batches[-1][:n_cpu], nlay[:n_cpu], noise[:n_cpu], args.cpu_per_core, want_out=True)
            arm.close()
            cb.pop("seconds", None)
            n_cpu = cb.pop("n")
            line["cpu_baseline"] = cb
            g = eng.eval(d_batches[-1][:n_cpu].contiguous(), d_nlay[:n_cpu].contiguous(), d_noise[:n_cpu].contiguous(),
                         want_synth=True)
            torch.cuda.synchronize(dev)
            line["parity"] = parity_metrics(targets, tuple(a.cpu().numpy() for a in g), ora)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
