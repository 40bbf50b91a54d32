#!/bin/bash
# N=2: stream-ordering test + sharded workload with the NCCL scatter / gather-v legs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_frames.py -m gpu -x -q -k "ordered or frame" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 2 --workload lz4-shuffle-cl5-8GiB-sharded > gpurun_out/bench_sharded_n2.json 2> gpurun_out/bench_sharded_n2.err; echo "sharded rc=$?"
grep -i "error\|Traceback" gpurun_out/bench_sharded_n2.err | head -5
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_sharded_n2.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value %.1f comp %.1f dec %.1f e2e %.1f ms/step %.2f" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"], d["ms_per_step"]), d.get("with_scatter_gather"), {k:round(v["value"],1) for k,v in d.get("typesize_sweep",{}).items()})
PY
