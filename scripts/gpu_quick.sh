#!/bin/bash
# quick GPU iteration: parity tests + both bench workloads (+ optional ncu of a workload's codec kernels)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for W in lz4-shuffle-ts4-cl5-256MiB blosclz-bitshuffle-ts8-cl5-256MiB; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$W.json").read().strip().splitlines()[-1])
print("$W", "value %.1f comp %.1f dec %.1f | e2e %.1f (c %.1f d %.1f) | cpu %s" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"], d["e2e"]["compress_gbs"], d["e2e"]["decompress_gbs"], d.get("cpu_baseline",{}).get("value")))
print({k:round(v["ms_avg"],3) for k,v in d["kernels"].items()})
PY
done
if [ -n "$1" ]; then bash scripts/gpu_ncu.sh $1 $2; fi
