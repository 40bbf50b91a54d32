#!/bin/bash
# fastest GPU iteration: LZ4 workload bench only (no tests), optional ncu capture tag
mkdir -p gpurun_out
W=lz4-shuffle-ts4-cl5-256MiB
timeout 600 python bench.py --workload $W --steps 10 --warmup 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$W.json").read().strip().splitlines()[-1])
print("$W", "value %.1f comp %.1f dec %.1f | e2e %.1f (c %.1f d %.1f) | cpu %s" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"], d["e2e"]["compress_gbs"], d["e2e"]["decompress_gbs"], d.get("cpu_baseline",{}).get("value")))
print({k:round(v["ms_avg"],3) for k,v in d["kernels"].items()})
PY
if [ -n "$1" ]; then bash scripts/gpu_ncu.sh $W $1; fi
