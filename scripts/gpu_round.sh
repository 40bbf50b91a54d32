#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list.
# Usage (from the dev container):  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh'
mkdir -p gpurun_out
{ nvidia-smi -L; nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; free -g | head -2; } > gpurun_out/host.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --workload blosclz-bitshuffle-ts8-cl5-256MiB --steps 10 --warmup 3 > gpurun_out/bench_ours_cfg3.json 2> gpurun_out/bench_ours_cfg3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/pytest_gpu.txt; cat gpurun_out/smoke.txt | tail -5; cat gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err; cat gpurun_out/bench_ref.json; cat gpurun_out/bench_ours_cfg3.json; cat gpurun_out/host.txt
