#!/bin/bash
# multi-GPU check: bench.py under torchrun exactly as the driver launches it (default workload, both arms)
# and the 8 GiB sharded workload with the NCCL scatter / gather-v legs.   usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "default rc=$?"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "reference rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 2 --workload lz4-shuffle-cl5-8GiB-sharded > gpurun_out/bench_sharded_n$N.json 2> gpurun_out/bench_sharded_n$N.err; echo "sharded rc=$?"
tail -3 gpurun_out/bench_n$N.err gpurun_out/bench_sharded_n$N.err
python - <<PY
import json
for f in ("bench_n$N","bench_ref_n$N","bench_sharded_n$N"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "value %.1f comp %.1f dec %.1f e2e %.1f ms/step %.2f" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"], d["ms_per_step"]), d.get("with_scatter_gather"), {k:round(v["value"],1) for k,v in d.get("typesize_sweep",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
