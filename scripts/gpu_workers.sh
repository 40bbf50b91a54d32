#!/bin/bash
# frames: chunks in flight (BLOSC_B200_FRAME_WORKERS) vs throughput on the 8 GiB workload; cfg 3 ncu capture
mkdir -p gpurun_out
for W in 2 4 6 8; do
  BLOSC_B200_FRAME_WORKERS=$W timeout 600 python bench.py --workload lz4-shuffle-cl5-8GiB-sharded --steps 3 --warmup 1 > gpurun_out/bench_workers_$W.json 2> gpurun_out/bench_workers_$W.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_workers_$W.json").read().strip().splitlines()[-1])
print("workers $W", "value %.1f comp %.1f dec %.1f e2e %.1f" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"]), {k:round(v["value"],1) for k,v in d["typesize_sweep"].items()})
PY
done
bash scripts/gpu_ncu.sh blosclz-bitshuffle-ts8-cl5-256MiB cfg3
