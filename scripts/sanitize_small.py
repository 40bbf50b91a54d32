#!/usr/bin/env python
"""Tiny workload for compute-sanitizer --tool racecheck (shared-memory hazards in the codec kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import __graft_entry__ as g
from datagen import gen
pkg = g.load_package()
for kind in ("bench", "text"):
    n = 600000
    src = gen(kind, n)
    d_src = torch.from_numpy(src).cuda()
    for comp, ts, shuf, clevel in (("lz4", 4, 1, 5), ("blosclz", 8, 2, 5), ("blosclz", 4, 1, 9), ("lz4", 2, 2, 1)):
        d_chunk = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
        cb = pkg.compress_ctx(clevel, shuf, ts, n, d_src, d_chunk, n + 16, comp)
        d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
        assert cb > 0 and pkg.decompress_ctx(d_chunk, d_out, n) == n and torch.equal(d_out, d_src)
print("racecheck workload ok")
