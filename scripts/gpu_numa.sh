#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -8
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_numa_$i.json 2>gpurun_out/bench_numa_$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_numa_$i.json").read().strip().splitlines()[-1])
print("run $i value %.1f e2e %.1f (c %.1f d %.1f)"%(d["value"], d["e2e"]["value"], d["e2e"]["compress_gbs"], d["e2e"]["decompress_gbs"]), d["config"].get("host_buffers"), d["clocks"], "cpu", round(d["cpu_baseline"]["value"],1))
PY
done
