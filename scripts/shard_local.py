#!/usr/bin/env python
"""shard_local.py -- the per-rank compute of the pipelined sharding path on ONE GPU (no collective): K chunks through
compress_sharded_pipelined / decompress_sharded_pipelined (dist=None) next to the frames API on the same buffer."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g
pkg = g.load_package()
from cblosc_b200 import sharding
K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chunk = 256 << 20
i = np.arange(chunk // 4, dtype=np.uint32)
one = torch.from_numpy((((i << 26) ^ (i << 18) ^ (i << 11) ^ (i << 3) ^ i) & ((1 << 19) - 1)).view(np.uint8).copy()).cuda()
full = one.repeat(K); total = K * chunk; dev = full.device
bound = pkg.frame_bound(total, 1, chunk)
d_frame = torch.empty(bound, dtype=torch.uint8, device=dev); d_out = torch.empty(total, dtype=torch.uint8, device=dev)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fb = pkg.frame_compress(5, 1, 4, total, full, d_frame, bound, "lz4", 0, chunk)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    nb = pkg.frame_decompress(d_frame, fb, d_out, total)
    torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"frames API      : compress {1e3*(t1-t0):7.1f} ms  decompress {1e3*(t2-t1):7.1f} ms  ({total/(t1-t0)/1e9:.0f} / {total/(t2-t1)/1e9:.0f} GB/s)")
for w in (4, 8):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        frames, sizes = sharding.compress_sharded_pipelined(pkg, None, full, total, chunk, 0, 1, dev, clevel=5, doshuffle=1, typesize=4, compressor="lz4", workers=w)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        back = sharding.decompress_sharded_pipelined(pkg, None, frames, sizes, total, chunk, 0, 1, dev, workers=w)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    assert torch.equal(back, full)
    print(f"pipelined, {w} thr: compress {1e3*(t1-t0):7.1f} ms  decompress {1e3*(t2-t1):7.1f} ms  ({total/(t1-t0)/1e9:.0f} / {total/(t2-t1)/1e9:.0f} GB/s)")
