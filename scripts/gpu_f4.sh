#!/bin/bash
# zlib / zstd decode on the GPU + the whole GPU suite + a short bench for sanity
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_after_f4.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_after_f4.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f"%(d["value"], d["e2e"]["value"]), {k:round(v["ms_avg"],3) for k,v in d["kernels"].items()}, d["clocks"])
PY
python - <<'PY'
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import __graft_entry__ as g
from datagen import bench_words, compress
pkg = g.load_package()
ref = C.CDLL("oracle/_ref/libblosc_ref.so"); ref.blosc_compress_ctx.restype = C.c_int
n = 64 << 20
src = bench_words(n)
for codec in ("zlib", "zstd"):
    cb, chunk = compress(ref, "blosc_compress_ctx", 5, 1, 4, src, n + 16, codec, 0, 16)
    d_chunk = torch.from_numpy(chunk[:cb].copy()).cuda(); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2): assert pkg.decompress_ctx(d_chunk, d_out, n) == n
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): pkg.decompress_ctx(d_chunk, d_out, n)
    dt = (time.perf_counter() - t0) / 3
    assert (d_out.cpu().numpy() == src).all()
    print("%s chunk 64 MiB (ratio %.1f): GPU decode %.2f GB/s (%.1f ms)" % (codec, n / cb, n / dt / 1e9, dt * 1e3))
PY
