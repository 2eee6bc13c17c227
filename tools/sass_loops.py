"""Static loop report from `cuobjdump -sass`: every backward branch = one loop; prints its size,
opcode histogram and how many MUFU.RSQ64H it holds (to spot the secular-function layer loops)."""
import re, sys, subprocess, collections
so, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = None; ins = []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); continue
    if fn and pat in fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
print("function instructions:", len(ins))
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*$", t)
    m = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
    if not m: continue
    tgt = int(m.group(1), 16)
    if tgt < a and tgt in addr_idx:
        body = ins[addr_idx[tgt]:i + 1]
        ops = collections.Counter()
        for _, x in body:
            toks = x.split()
            o = toks[1] if toks[0].startswith("@") else toks[0]
            ops[o.split(".")[0]] += 1
        rsq = sum(1 for _, x in body if "RSQ64H" in x)
        f64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
        print("loop %#x..%#x  n=%d  fp64=%d rsq=%d  %s" % (tgt, a, len(body), f64, rsq, dict(ops.most_common(12))))
