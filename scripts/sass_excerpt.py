#!/usr/bin/env python
"""sass_excerpt.py lib.so -> profiles/r2_sass_<kernel>.txt: instruction mix of every kernel and the SASS around its memory
instructions (cuobjdump -sass; no GPU needed).  What to look for: LDG.E.128 / STG.E batches in the filter, LDS/STS/ATOMS
in the codecs, BAR.SYNC / BAR.ARV (named barriers) in the team encoder -- and the absence of UTMALDG / UBLKCP: these
kernels move bytes with ordinary vector loads, the working sets are tables and rings in shared memory, not tiles."""
import collections, os, re, subprocess, sys
lib = sys.argv[1]; outdir = sys.argv[2] if len(sys.argv) > 2 else "profiles"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
summary = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+([^;]+);", f)
    ops = collections.Counter(re.sub(r"@!?U?P\d\s+", "", i).split()[0] for i in ins)
    short = re.sub(r"^_Z\d+", "", name)
    short = re.sub(r"(ILi(\d)E)?Ev?\d*.*$", lambda m: ("_" + m.group(2)) if m.group(2) else "", short)
    mem = [(k, v) for k, v in ops.most_common() if re.match(r"(LDG|STG|LDS|STS|ATOMS|ATOMG|RED|BAR|SHFL|MATCH|VOTE|PRMT|SHF|LDC|UTMA|UBLKCP|LDGSTS|CCTL|MEMBAR|ERRBAR)", k)]
    summary.append((short, len(ins), mem))
    with open(os.path.join(outdir, f"r2_sass_{short}.txt"), "w") as o:
        o.write(f"{name}: {len(ins)} SASS instructions (cuobjdump -sass, sm_100a)\n")
        o.write("memory / warp-level instruction mix: " + ", ".join(f"{k} {v}" for k, v in mem) + "\n\n")
        # excerpt: the first 3 windows of 24 instructions around 128-bit global loads or atomics / barriers
        lines = [l for l in f.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        keys = [i for i, l in enumerate(lines) if re.search(r"LDG\.E\.128|ATOMS|BAR\.(SYNC|ARV)|MATCH|LDS\.64", l)]
        shown, last = 0, -100
        for i in keys:
            if i - last < 40: continue
            o.write(f"--- around instruction {i}\n")
            for l in lines[max(0, i - 6):i + 18]:
                o.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() + "\n")
            o.write("\n"); last = i; shown += 1
            if shown >= 3: break
with open(os.path.join(outdir, "r2_sass_summary.txt"), "w") as o:
    o.write("kernel, SASS instructions, memory / warp-level instruction mix (cuobjdump -sass of libblosc_b200.so, sm_100a)\n")
    for short, n, mem in summary:
        o.write(f"{short:28s} {n:6d}  " + ", ".join(f"{k} {v}" for k, v in mem) + "\n")
print(open(os.path.join(outdir, "r2_sass_summary.txt")).read())
