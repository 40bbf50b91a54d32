#!/usr/bin/env python
"""ncu_lines.py rep lib.so kernel_mangled_substr [top] -- warp-stall samples and executed instructions per CUDA source line.
Joins the per-SASS-instruction table of an ncu capture (--page source) with nvdisasm's line info of the same kernel
(instruction order is the same in both)."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "blosc_b200." not in f.replace("backend_cuda-", "")] or os.listdir(tmp)
dis = ""
for f in os.listdir(tmp):
    d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kern in d:
        dis = d; break
lines = dis.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
ins = []          # (file:line) per instruction
cur = "?"
for l in lines[start + 1:]:
    if l.startswith("//-----") or (l.startswith(".text.") ):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"; continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
        ins.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ci = h.index("Warp Stall Sampling (All Samples)"); ii = h.index("Instructions Executed")
if len(data) != len(ins):
    print(f"warning: {len(data)} profiled instructions vs {len(ins)} disassembled", file=sys.stderr)
samp = collections.Counter(); ex = collections.Counter()
for r, loc in zip(data, ins):
    samp[loc] += int(r[ci] or 0); ex[loc] += int(r[ii] or 0)
ts, te = sum(samp.values()) or 1, sum(ex.values()) or 1
srcs = {}
def text(loc):
    f, n = loc.split(":") if ":" in loc else (loc, "0")
    for root in ("c-blosc_b200/csrc",):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            k = int(n) - 1
            return srcs[p][k].strip()[:100] if 0 <= k < len(srcs[p]) else ""
    return ""
print(f"total samples {ts}, warp instructions {te}")
for loc, s in samp.most_common(top):
    print(f"{100*s/ts:5.1f}% smp {100*ex[loc]/te:5.1f}% ins  {loc:24s} {text(loc)}")
