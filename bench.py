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

  value   device-resident inputs, CUDA-event time per step, L2 flushed between
          steps (not timed), max over ranks
  e2e     the same metric through the host-buffer C-ABI call
          (bh_engine_eval_host): pinned host inputs -> H2D -> kernels -> D2H of
          logL / misfits / status, wall clock, every step
  roofline / fp64   dominant kernel, timed live with CUDA events inside the engine:
                    algorithmic bytes against the HBM peak, algorithmic flops against
                    the fp64 DFMA peak (fp64.frac), and the fp64 pipe's utilisation
                    (fp64.pipe: ncu's figure and executed fp64 warp instructions x 2
                    cycles against the live launch duration, profiles/traffic.json)
  cpu_baseline      the oracle (reference rfmini C++ when compiled, SURF96 C
                    restatement, numpy likelihood) on the host cores, bounded sample

`--impl reference` times that CPU path alone (rank 0 only).
"""
import argparse
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
def workload(cfg_name, B, seed):
    """Targets (observed = st3 truth forward-modelled by the oracle-free closed data below)
    and the seeded model batch of one rank."""
    from bayhunter_b200 import synthetic
    c = synthetic.CONFIGS[cfg_name]
    rng = np.random.default_rng(seed + 1000)
    targets = []
    for ref in c["refs"]:
        if ref in ("prf", "srf"):
            x = synthetic.rf_time_axis(c["rf"])
            y = rng.normal(0.0, 0.02, x.size)         # observed trace: synthetic noise-like data
        else:
            x = np.asarray(c["periods"], dtype=np.float64)
            y = 3.0 + 0.02 * x + rng.normal(0.0, 0.02, x.size)
        targets.append((ref, x, y))
    rows, nlay = synthetic.draw_batch(B, c["nrows"], seed=seed)
    noise = synthetic.draw_noise(B, c["refs"], seed=seed + 1)
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
# CPU baseline (oracle) -- the only place bench.py executes oracle/
# --------------------------------------------------------------------------------------
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"   # tutorialhunt.py:12-14
    targets, rows, nlay, noise = args
    from oracle import joint_oracle as jo
    ot = [jo.OracleTarget(ref, x, y, cov="exp") for ref, x, y in targets]
    t0 = time.perf_counter()
    jo.evaluate_batch(ot, rows, nlay, noise)
    return time.perf_counter() - t0


def cpu_baseline(targets, rows, nlay, noise, per_core, cores=None):
    """evals/s of the CPU path with the reference's own parallel model: one
    single-threaded process per core (src/mcmcOptimizer.py:219-252)."""
    import multiprocessing as mp
    from oracle import joint_oracle as jo
    jo.lib()
    cores = cores or os.cpu_count() or 1
    n = min(rows.shape[0], per_core * cores)
    per = max(1, n // cores)
    jobs = [(targets, rows[i * per:(i + 1) * per], nlay[i * per:(i + 1) * per], noise[i * per:(i + 1) * per])
            for i in range(cores)]
    jobs = [j for j in jobs if j[1].shape[0] > 0]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    nev = sum(j[1].shape[0] for j in jobs)
    kind = "port"   # SURF96 has no compiled reference here (no Fortran compiler); RF uses oracle/_ref when present
    return dict(value=nev / wall, unit=UNIT, cores=len(jobs), kind=kind,
                sample="%d models (%d per core) of the same batch; RF via %s, SWD via C restatement of "
                       "surfdisp96.f, likelihood dense numpy as Targets.py" %
                       (nev, per, "oracle/_ref (reference rfmini C++)" if jo.ref_rfmini() is not None
                        else "C restatement of rfmini"),
                seconds=wall)


# --------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    from bayhunter_b200 import synthetic
    cores = os.cpu_count() or 1
    per_core = args.ref_per_core
    c, targets, rows, nlay, noise = workload(cfg, per_core * cores, seed=20260101)
    times = []
    last = None
    for s in range(args.warmup + args.steps):
        rng = np.random.default_rng(5000 + s)
        r = synthetic.perturb_batch(rows, nlay, rng)
        last = cpu_baseline(targets, r, nlay, noise, per_core, cores)
        if s >= args.warmup:
            times.append(last["seconds"])
    nev = per_core * last["cores"]
    value = nev * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same workload as the b200 arm; each step evaluates a bounded sample of it
        "config": {"workload": cfg_description(cfg, args.chains_per_gpu or synthetic.CONFIGS[cfg]["B"], c),
                   "sample_models_per_step": nev},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cfg_description(cfg, B, c):
    rf = "" if c["rf"] is None else " + P-RF %d samples (nsamp %d)" % (c["rf"]["n"], 2 ** int(np.ceil(np.log2(2 * c["rf"]["n"]))))
    nr = c["nrows"]
    rows = "%d rows (%d layers + half-space)" % (nr, nr - 1) if np.isscalar(nr) else "%d-%d rows" % nr
    return "%s: %s, %d periods%s, %s, %d chains per GPU" % (cfg, "+".join(r for r in c["refs"] if r not in ("prf", "srf")),
                                                       len(c["periods"]), rf, rows, B)


# --------------------------------------------------------------------------------------
# main arm
# --------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="joint5", choices=["swd2", "joint5", "transd3"])
    ap.add_argument("--chains-per-gpu", type=int, default=0)
    ap.add_argument("--ref-per-core", type=int, default=96)
    ap.add_argument("--cpu-per-core", type=int, default=192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pool-samples", type=int, default=100)
    ap.add_argument("--sampler-iters", type=int, default=40,
                    help="lock-step MCMC iterations timed for the auxiliary 'sampler' object (0: skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import bayhunter_b200 as bh
    from bayhunter_b200 import synthetic, chains

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bh._lib.require_device()
    bh._lib.set_device(local_rank)

    cfg = args.config
    B = args.chains_per_gpu or synthetic.CONFIGS[cfg]["B"]
    seed = chains.chain_seed(20260101, rank)
    c, targets, rows0, nlay, noise = workload(cfg, B, seed)
    specs = [bh.TargetSpec(ref, x, y, cov="exp") for ref, x, y in targets]
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
    sampler = ClockSampler(local_rank)
    kernel_ms = {}
    counts = [0, 0]
    for s in range(args.warmup):
        eng.eval(d_batches[s], d_nlay, d_noise, out=out)
        flush.fill_(s & 0xFF)
    barrier()
    if rank == 0:
        sampler.start()
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
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [ev[s][0].elapsed_time(ev[s][1]) for s in range(args.warmup, nsteps)]
    total_ms = float(sum(step_ms))
    valid_frac = float(out[2].float().mean().item())

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_batches = [pin(b) for b in batches]
    h_nlay, h_noise = pin(nlay), pin(noise)
    h_logL = torch.empty(B, dtype=torch.float64).pin_memory()
    h_mis = torch.empty((B, T + 1), dtype=torch.float64).pin_memory()
    h_stat = torch.empty(B, dtype=torch.int32).pin_memory()
    eng.set(profile=0)
    for s in range(args.warmup):
        eng.eval_host_ptr(h_batches[s].data_ptr(), h_nlay.data_ptr(), h_noise.data_ptr(), B, L,
                          h_logL.data_ptr(), h_mis.data_ptr(), h_stat.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for s in range(args.warmup, nsteps):
        eng.eval_host_ptr(h_batches[s].data_ptr(), h_nlay.data_ptr(), h_noise.data_ptr(), B, L,
                          h_logL.data_ptr(), h_mis.data_ptr(), h_stat.data_ptr())
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    h2d = batches[0].nbytes + nlay.nbytes + noise.nbytes
    d2h = B * 8 + B * (T + 1) * 8 + B * 4
    # device and host paths must agree bit for bit on the last batch
    same = bool(np.array_equal(h_logL.numpy(), out[0].cpu().numpy(), equal_nan=True))

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

    # ---- auxiliary: the device sampler around the same hot path ----
    # B chains advance in lock step (propose kernel -> engine -> accept kernel, no host round trip);
    # reported beside the headline metric, not part of it.
    smp = None
    if args.sampler_iters > 0:
        from bayhunter_b200 import Targets as T_, SingleChain as sc
        cls = {"rdispph": T_.RayleighDispersionPhase, "rdispgr": T_.RayleighDispersionGroup,
               "ldispph": T_.LoveDispersionPhase, "ldispgr": T_.LoveDispersionGroup,
               "prf": T_.PReceiverFunction, "srf": T_.SReceiverFunction}
        jt = T_.JointTarget([cls[ref](x, y) for ref, x, y in targets])
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
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])
        cnt = torch.tensor(counts, dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    if rank == 0:
        K = args.steps
        value = world * B * K / (total_ms * 1e-3)
        e2e_value = world * B * K / e2e_s
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
        # bytes/flops of the dominant kernel alone: the dispersion kernel reads the REAL*4 rows
        # (16 B per layer row) and writes the two root tables and the curves
        if dom in ("swd", "swd_love"):
            nsw = sum(1 for r in c["refs"] if r not in ("prf", "srf"))
            dom_bytes = float(np.sum(16.0 * nlay)) + 3 * 8.0 * nsw * len(c["periods"]) * B
            nr = sum(1 for r in c["refs"] if r.startswith("r") and r not in ("prf", "srf"))
            Lm = float(nlay.mean())
            dom_flops = counts[0] / K * (nr * (175.0 * (Lm - 1) + 30.0) + (nsw - nr) * (28.0 * (Lm - 1) + 10.0)) / nsw
        else:
            dom_bytes, dom_flops = nbytes, flops
        dom_s = kmean[dom] * 1e-3
        fp64_peak = 148 * 64 * 2 * float(peaks.get("sm_max_mhz", 1965.0)) * 1e6 / 1e12   # DFMA/clk/SM * 2 flop
        traffic = None
        pipe = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(cfg, {}).get(dom + "_kernel")
            pp = tr.get("fp64_pipe", {}).get(cfg)
            if pp and dom == "swd":
                # the fp64 pipe of a sub-partition takes one warp instruction per 2 cycles: instructions of the
                # profiled launch (same workload) against the pipe cycles of the LIVE launch duration and clock
                mhz = float(clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
                pipe = {"busy_ncu": pp["busy_pct_ncu"] / 100.0,
                        "busy_live": pp["fp64_warp_instructions"] * 2.0 / (148 * 4 * mhz * 1e6 * dom_s),
                        "fp64_warp_instructions_per_launch": pp["fp64_warp_instructions"],
                        "source": "profiles/traffic.json (ncu --set full capture of the same workload); busy_live = "
                                  "fp64 warp instructions x 2 cycles / (592 sub-partitions x SM clock x live kernel time)"}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg_description(cfg, B, c), "global_chains": world * B,
                       "l2": "flushed between timed steps (256 MiB write, outside the event pairs)",
                       "valid_fraction": valid_frac, "parallelism": "chains sharded, dp%d" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "bh_engine_eval_host (pinned host buffers)", "matches_device_path": same},
            # per step: prepare(SWD rows), layer_order, swd_kernel, loglik (+ swd_gate, prepare(RF tables),
            # rf_spectrum, rf_synth with an RF target)
            "gpu_launches": (8 if c["rf"] is not None else 4) * K,
            "roofline": {"bound": "hbm", "kernel": dom + "_kernel", "achieved": dom_bytes / dom_s / 1e9, "peak": hbm_peak,
                         "unit": "GB/s", "frac": dom_bytes / dom_s / 1e9 / hbm_peak, "traffic": traffic,
                         "algorithmic_bytes_per_launch": dom_bytes,
                         "peak_source": peak_src,
                         "note": "fp64-compute-bound path: see 'fp64' for the binding roofline"},
            "fp64": {"kernel": dom + "_kernel", "achieved_tflops": dom_flops / dom_s / 1e12,
                     "peak_tflops": fp64_peak, "frac": dom_flops / dom_s / 1e12 / fp64_peak,
                     "peak_source": "148 SM x 64 DFMA/clk x 2 x sm_max_mhz; DFMA issue rate measured (tools/micro/fp64_latency.cu: "
                                    "1 warp-DFMA per 2 cycles per sub-partition), clock nominal",
                     "step_flops": flops, "step_tflops": flops / (total_ms / K * 1e-3) / 1e12,
                     "secular_evals_per_step": counts[0] / K, "secular_evaluated_per_step": counts[1] / K,
                     "pipe": pipe},
            "kernel_ms": kmean,
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
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
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline(targets, batches[-1], nlay, noise, args.cpu_per_core)
            cb.pop("seconds", None)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
