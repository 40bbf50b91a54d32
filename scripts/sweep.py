#!/usr/bin/env python
"""Device-resident compress/decompress GB/s and ratio over the BASELINE.json config sweep
(typesize {1,2,4,8,16} x codec x filter) on one GPU -- supplementary to bench.py."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
from bench import bench_words
pkg = g.load_package()
n = 256 << 20
src = torch.from_numpy(bench_words(n, np).copy()).cuda()
chunk = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
out = torch.zeros(n, dtype=torch.uint8, device="cuda")
rows = []
cfgs = [(c, s, ts) for c in ("lz4", "blosclz") for s in (1, 2) for ts in (1, 2, 4, 8, 16)] + [("lz4", 0, 4), ("blosclz", 0, 4)]
if len(sys.argv) > 1:
    cfgs = [c for c in cfgs if c[0] == sys.argv[1] and (len(sys.argv) < 3 or c[1] == int(sys.argv[2]))]
for comp, shuf, ts in cfgs:
    def once():
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cb = pkg.compress_ctx(5, shuf, ts, n, src, chunk, n + 16, comp)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        nb = pkg.decompress_ctx(chunk, out, n)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        return cb, nb, t1 - t0, t2 - t1
    once(); once()
    r = [once() for _ in range(5)]
    cb, nb = r[0][0], r[0][1]
    tc = sorted(x[2] for x in r)[2]; td = sorted(x[3] for x in r)[2]
    ok = nb == n and bool(torch.equal(out, src))
    rows.append({"codec": comp, "filter": ["none", "shuffle", "bitshuffle"][shuf], "typesize": ts, "cbytes": cb, "ratio": round(n / cb, 2),
                 "compress_gbs": round(n / tc / 1e9, 1), "decompress_gbs": round(n / td / 1e9, 1), "roundtrip_ok": ok})
    print(rows[-1], flush=True)
json.dump(rows, open("gpurun_out/sweep.json", "w"), indent=1)
