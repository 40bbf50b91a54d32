#!/bin/bash
# packed LZ4 tables for chunks in flight: frames parity on the GPU, headline unchanged, 8 GiB workload with / without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frames.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pack_default.json 2>/dev/null
for P in 0 auto; do
  if [ $P = auto ]; then unset BLOSC_B200_LZ4_PACK; else export BLOSC_B200_LZ4_PACK=$P; fi
  timeout 600 python bench.py --workload lz4-shuffle-cl5-8GiB-sharded --steps 2 --warmup 1 > gpurun_out/bench_pack_$P.json 2>/dev/null
done
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_pack_default.json").read().strip().splitlines()[-1])
print("default value %.1f e2e %.1f"%(d["value"], d["e2e"]["value"]), {k:round(v["ms_avg"],3) for k,v in d["kernels"].items()})
for P in ("0","auto"):
    d=json.loads(open("gpurun_out/bench_pack_%s.json"%P).read().strip().splitlines()[-1])
    print("pack", P, "value %.1f comp %.1f dec %.1f e2e %.1f" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"]), {k:round(v["value"],1) for k,v in d["typesize_sweep"].items()}, round(d["kernels"]["encode"]["ms_avg"],2))
PY
