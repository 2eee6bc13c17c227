"""Per-SASS-region stall breakdown from an .ncu-rep: groups consecutive instructions by the CUDA
source line they map to is not available in the sass-only view, so this prints (a) total samples
per stall reason and (b) the top-N instructions by stall_wait / stall_math / stall_short_sb with
their opcode and the preceding instruction (the usual producer)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); ins = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    d = {k: int(r[ix[k]] or 0) for k in reasons}
    for k, v in d.items(): tot[k] += v
    ins.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), d, int(r[ix["# Samples"]] or 0)))
S = sum(tot.values())
print("samples", S)
for k, v in tot.most_common(): print("  %-24s %6.2f%%" % (k, 100.0 * v / S))
opw = collections.Counter(); opn = collections.Counter()
for i, (src, n, d, s) in enumerate(ins):
    t = src.split(); o = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    o = o.split(".")[0]
    opw[o] += d["stall_wait"]; opn[o] += s
print("stall_wait by opcode of the stalled instruction:")
for o, v in opw.most_common(12): print("  %-8s wait %6.2f%%  all-samples %6.2f%%" % (o, 100.0 * v / max(1, tot["stall_wait"]), 100.0 * opn[o] / S))
