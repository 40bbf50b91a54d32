#!/usr/bin/env python
"""kbench.py -- quick device-resident timing of compress / decompress with per-kernel CUDA-event
times (library's own profiling hooks), for A/B runs of library builds:
    BLOSC_B200_LIB=path/to/variant.so python scripts/kbench.py [tag] [codec:shuffle:typesize ...]
Prints one line per workload."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import __graft_entry__ as g

pkg = g.load_package()
tag = sys.argv[1] if len(sys.argv) > 1 else "lib"
specs = sys.argv[2:] or ["lz4:1:4", "blosclz:2:8"]
nbytes = int(os.environ.get("KBENCH_BYTES", 256 << 20))
steps = int(os.environ.get("KBENCH_STEPS", 10))
i = np.arange(nbytes // 4, dtype=np.uint32)
src = (((i << np.uint32(26)) ^ (i << np.uint32(18)) ^ (i << np.uint32(11)) ^ (i << np.uint32(3)) ^ i) & np.uint32((1 << 19) - 1)).view(np.uint8)
d_src = torch.from_numpy(src.copy()).cuda()
d_chunk = torch.zeros(nbytes + 16, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
for spec in specs:
    comp, shuf, ts = spec.split(":")
    shuf, ts = int(shuf), int(ts)
    for _ in range(3):
        cb = pkg.compress_ctx(5, shuf, ts, nbytes, d_src, d_chunk, nbytes + 16, comp)
        nb = pkg.decompress_ctx(d_chunk, d_out, nbytes)
    assert cb > 0 and nb == nbytes and torch.equal(d_out, d_src), (cb, nb)
    pkg.set_profiling(True); pkg.prof_reset()
    torch.cuda.synchronize()
    tc = td = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        pkg.compress_ctx(5, shuf, ts, nbytes, d_src, d_chunk, nbytes + 16, comp)
        t1 = time.perf_counter()
        pkg.decompress_ctx(d_chunk, d_out, nbytes)
        td += time.perf_counter() - t1; tc += t1 - t0
    prof = pkg.prof_get(); pkg.set_profiling(False)
    k = {n: round(v[0] / max(v[1], 1), 4) for n, v in prof.items() if v[1]}
    print(f"[{tag}] {spec:16s} cbytes {cb:10d}  comp {nbytes / (tc / steps) / 1e9:7.1f} GB/s  dec {nbytes / (td / steps) / 1e9:7.1f} GB/s  "
          f"c+d {2 * nbytes / ((tc + td) / steps) / 1e9:7.1f}  kernels(ms) {k}", flush=True)
