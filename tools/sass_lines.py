"""Static SASS instruction counts per source line for one kernel of the in-tree library
(nvdisasm -g line info): where the code size is.  usage: sass_lines.py <kernel-substr> [top]"""
import re, subprocess, sys, os, tempfile, collections, glob
pat = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bayhunter_b200", "libbayhunter_b200.so")
if len(sys.argv) > 3: lib = sys.argv[3]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
cnt = collections.Counter(); files = collections.Counter(); total = 0
for cub in glob.glob(os.path.join(d, "swd_kernel*.cubin")) + glob.glob(os.path.join(d, "*.cubin")):
    out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    infn = False; cur = None
    for line in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", line)
        if m: infn = pat in m.group(1); continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/\s+\S", line):
            cnt[cur] += 1; total += 1
            if cur: files[cur[0]] += 1
    if total: break
print("total instructions", total, dict(files))
src = {}
for (f, ln), c in cnt.most_common(top):
    path = None
    for root in ("bayhunter_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, f)
        if os.path.exists(p): path = p
    text = ""
    if path:
        if path not in src: src[path] = open(path).read().splitlines()
        text = src[path][ln - 1].strip()[:90]
    print("%5d  %s:%d  %s" % (c, f, ln, text))
