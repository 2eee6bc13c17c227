"""Per-CUDA-source-line executed-instruction shares from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = {}
sass_tot = 0
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 2 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0] == "" and r[2].startswith("0x"):
        try: sass_tot += int(r[7])
        except ValueError: pass
        continue
    if hdr and len(r) > 8 and r[0] != "":
        try: ln = int(r[0]); inst = int(r[7]); samp = int(r[6]); thr = int(r[8])
        except ValueError: continue
        k = (cur, ln, r[1].strip())
        a = agg.setdefault(k, [0, 0, 0]); a[0] += inst; a[1] += samp; a[2] += thr
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("line-attributed inst", tot, "samples", ts)
for (f, ln, src), (inst, samp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% inst %5.2f%% samp  thr %4.1f  %s:%d  %s" % (100 * inst / tot, 100 * samp / max(ts, 1), thr / max(inst, 1), f, ln, src[:100]))
