#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck): ragged sizes, both codecs, all
filters, getitem, corrupted chunks.  Buffers are exact-size device allocations so any overrun shows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import __graft_entry__ as g
from datagen import gen
pkg = g.load_package()
rng = np.random.default_rng(0)
for kind in ("bench", "text", "mixed"):
    for n in (1000, 70001, 300000, (1 << 20) + 77):
        src = gen(kind, n)
        d_src = torch.from_numpy(src).cuda()
        for comp, ts, shuf, clevel in (("lz4", 4, 1, 5), ("blosclz", 8, 2, 5), ("lz4", 1, 0, 9), ("blosclz", 3, 1, 1), ("lz4", 16, 2, 5), ("blosclz", 4, 1, 9)):
            d_chunk = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
            cb = pkg.compress_ctx(clevel, shuf, ts, n, d_src, d_chunk, n + 16, comp)
            assert cb > 0
            d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
            assert pkg.decompress_ctx(d_chunk, d_out, n) == n and torch.equal(d_out, d_src)
            exact = d_chunk[:cb].clone()                      # exact-size chunk allocation
            assert pkg.decompress_ctx(exact, d_out, n) == n
            cnt = min(64, n // ts - 5)
            item = torch.empty(cnt * ts, dtype=torch.uint8, device="cuda")
            assert pkg.getitem(exact, 5, cnt, item) == cnt * ts
            bad = exact.clone()
            pos = torch.from_numpy(rng.integers(16, cb, 6)).cuda()
            bad[pos] = torch.from_numpy(rng.integers(0, 256, 6, dtype=np.uint8)).cuda()
            r = pkg.decompress_ctx(bad, d_out, n)
            assert r in (n, -1)
print("sanitize workload ok")
