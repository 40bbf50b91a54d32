#!/bin/bash
# frames + sharded workload on one GPU: parity tests, then the 8 GiB bench (both arms)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/frames_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/frames_pytest.log
tail -5 gpurun_out/frames_pytest.log
timeout 600 python bench.py --workload lz4-shuffle-cl5-8GiB-sharded --steps 5 --warmup 2 > gpurun_out/bench_sharded_n1.json 2> gpurun_out/bench_sharded_n1.err; echo "rc=$?"
tail -c 3000 gpurun_out/bench_sharded_n1.json; tail -5 gpurun_out/bench_sharded_n1.err
timeout 600 python bench.py --impl reference --workload lz4-shuffle-cl5-8GiB-sharded --steps 1 --warmup 1 > gpurun_out/bench_sharded_ref.json 2>&1
cat gpurun_out/bench_sharded_ref.json
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"
cat gpurun_out/bench_default.json
