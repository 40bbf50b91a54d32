#!/usr/bin/env python
"""Small workload for compute-sanitizer on the segment-parallel LZ4 parse (BLOSC_B200_PARSE=fast): ragged sizes, unsplit
streams longer than a parse window, exact-size device buffers so that any overrun shows."""
import os, sys
os.environ["BLOSC_B200_PARSE"] = "fast"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import __graft_entry__ as g
from datagen import gen
pkg = g.load_package()
for kind in ("bench", "text", "mixed", "zeros", "rand"):
    for n in (1000, 70001, 300001, (1 << 20) + 77):
        src = gen(kind, n)
        d_src = torch.from_numpy(src).cuda()
        for ts, shuf, clevel, bs in ((4, 1, 5, 0), (1, 0, 9, 0), (16, 2, 5, 0), (32, 1, 5, 1 << 19), (3, 1, 1, 0)):
            d_chunk = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
            cb = pkg.compress_ctx(clevel, shuf, ts, n, d_src, d_chunk, n + 16, "lz4", bs)
            assert cb > 0
            exact = d_chunk[:cb].clone()
            d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
            assert pkg.decompress_ctx(exact, d_out, n) == n and torch.equal(d_out, d_src)
    # unaligned device source
    base = torch.from_numpy(gen("bench", 400000 + 8)).cuda()
    for off in (1, 2, 3):
        d_chunk = torch.empty(400000 + 16, dtype=torch.uint8, device="cuda")
        cb = pkg.compress_ctx(5, 0, 1, 400000, base[off:off + 400000], d_chunk, 400000 + 16, "lz4")
        d_out = torch.empty(400000, dtype=torch.uint8, device="cuda")
        assert cb > 0 and pkg.decompress_ctx(d_chunk, d_out, 400000) == 400000 and torch.equal(d_out, base[off:off + 400000])
print("fast-parse sanitize workload ok")
