"""tools/make_profiles.py <tag> [round]: copy the judged summaries of a tools/gpu_record.sh run from
gpurun_out/ (scratch) into profiles/ (tracked): bench lines, ncu launch list + per-kernel shares,
ncu --set full summaries of the dispersion and RF spectrum kernels, DRAM traffic per launch."""
import collections, csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]; rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
shutil.copy(os.path.join(G, tag + "_bench.json"), os.path.join(P, rnd + "_bench.json"))
shutil.copy(os.path.join(G, tag + "_bench_ref.json"), os.path.join(P, rnd + "_bench_reference_arm.json"))
shutil.copy(os.path.join(G, tag + "_launches.csv"), os.path.join(P, rnd + "_launches.csv"))

rows = [r for r in csv.reader(open(os.path.join(G, tag + "_launches.csv"))) if len(r) > 5]
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if r[0] == "ID": hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    try: v = float(d["Metric Value"].replace(",", ""))
    except ValueError: continue
    u = d["Metric Unit"]
    ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    name = d["Kernel Name"].split("(")[0].split("::")[-1]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, rnd + "_launch_shares.txt"), "w") as f:
    f.write("# per-kernel totals of profiles/%s_launches.csv (ncu --metrics gpu__time_duration.sum --clock-control none;\n"
            "# python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --sampler-iters 2: 5 evaluations + the sustained run + the e2e passes + a short sampler run).\n"
            "# Times under ncu are cold-cache and serialised: the SHARES are what carries over to the live step.\n" % rnd)
    f.write("%-44s %8s %10s %7s\n" % ("kernel", "launches", "total ms", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-44s %8d %10.3f %6.1f%%\n" % (k[:44], n, t, 100 * t / tot))

traffic = {}
pipes = {}
for kern, short in (("swd", "swd_pool_kernel_rayleigh"), ("swdl", "swd_pool_kernel_love"), ("rf", "rf_spectrum")):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, kern))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_kernels.py"), rep], capture_output=True, text=True).stdout
    hist = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    hist = hist[hist.index("warp instructions executed"):] if "warp instructions executed" in hist else ""
    with open(os.path.join(P, "%s_%s_ncu_summary.txt" % (rnd, short)), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:%s %s python bench.py --steps 1 --warmup 3 "
                "--no-cpu-baseline --no-configs --sampler-iters 0  (joint5, B = 8192, one launch after 3 warm-up evaluations%s)\n"
                % ("rf_spectrum" if kern == "rf" else "swd_pool_kernel", {"swd": "-s 6 -c 1", "swdl": "-s 7 -c 1", "rf": "-s 3 -c 1"}[kern],
                   "" if kern == "rf" else "; under ncu the kernel has the device to itself, in the live step the Rayleigh and Love launches share the SMs"))
        f.write(out)
        f.write("\n## SASS opcode histogram (executed warp instructions)\n" + hist)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw))); h = rr[0]; units = rr[1]; vals = rr[2]
    tb = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(key); v = float(vals[i].replace(",", "")); u = units[i].lower()
        tb += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    traffic[short] = int(tb)
    if kern != "rf":
        # fp64 pipe: ncu's own utilisation figure and the executed fp64 warp instructions of the launch
        i = h.index("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
        pipe = {"busy_pct_ncu": float(vals[i].replace(",", ""))}
        i = h.index("gpu__time_duration.sum"); v = float(vals[i].replace(",", "")); u = units[i].lower()
        pipe["ms_ncu"] = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
        tot = 0; share = 0.0
        for line in hist.splitlines():
            t = line.split()
            if line.startswith("warp instructions executed:"): tot = int(t[-1])
            elif len(t) >= 2 and t[0] in ("DFMA", "DMUL", "DADD", "DSETP"): share += float(t[1].rstrip("%"))
        pipe["fp64_warp_instructions"] = int(tot * share / 100.0)
        pipe["warp_instructions"] = tot
        pipes[short] = pipe
tsum = sum(p["ms_ncu"] for p in pipes.values())
pair = {"kernel": "swd_pool",
        "busy_pct_ncu": sum(p["busy_pct_ncu"] * p["ms_ncu"] for p in pipes.values()) / tsum,
        "fp64_warp_instructions": sum(p["fp64_warp_instructions"] for p in pipes.values()),
        "warp_instructions": sum(p["warp_instructions"] for p in pipes.values()),
        "launches": pipes}
tj = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures summarised in this directory (bytes)",
      "joint5": {"swd_pool_kernel": traffic["swd_pool_kernel_rayleigh"] + traffic["swd_pool_kernel_love"],
                 "swd_pool_kernel_rayleigh": traffic["swd_pool_kernel_rayleigh"], "swd_pool_kernel_love": traffic["swd_pool_kernel_love"],
                 "rf_spectrum_kernel": traffic["rf_spectrum"]},
      "fp64_pipe": {"_comment": "the two swd_pool_kernel launches of one evaluation (Rayleigh, Love), each captured alone: "
                                "sm__pipe_fp64_cycles_active (% of peak; pair = time-weighted) and executed DFMA+DMUL+DADD+DSETP warp "
                                "instructions (each occupies the pipe of its sub-partition for 2 cycles)",
                    "joint5": pair}}
json.dump(tj, open(os.path.join(P, "traffic.json"), "w"), indent=2)
print(open(os.path.join(P, rnd + "_launch_shares.txt")).read()); print(tj)
