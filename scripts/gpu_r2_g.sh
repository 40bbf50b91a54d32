#!/bin/bash
# round 2, call G: evidence snapshot -- whole GPU suite, smoke, bench.py both arms, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -3 gpurun_out/r2g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -2 gpurun_out/r2g_smoke.log
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err; head -c 400 gpurun_out/r2g_bench_ref.json; echo
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 600 gpurun_out/r2g_bench.err; head -c 1500 gpurun_out/r2g_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --workload lz4-shuffle-ts4-cl5-256MiB --no-cpu --no-traffic > gpurun_out/r2g_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2g_launches.csv
