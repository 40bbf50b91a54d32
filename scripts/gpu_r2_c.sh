#!/bin/bash
# round 2, call C: team-mode LZ4 encoder (walker + 3 preparer warps) -- parity on hardware, A/B, ncu; new bench.py
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_hostile_input.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
SPECS="lz4:1:4 lz4:1:2 lz4:1:8 lz4:1:16 lz4:2:4 lz4:0:4"
BLOSC_B200_LZ4_TEAM=0 timeout 600 python scripts/kbench.py warp $SPECS 2>&1 | tee gpurun_out/r2c_kbench_warp.log
timeout 600 python scripts/kbench.py team $SPECS blosclz:2:8 2>&1 | tee gpurun_out/r2c_kbench_team.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_team_kernel -s 3 -c 1 -f -o gpurun_out/enc_r2c python scripts/kbench.py ncu lz4:1:4 > gpurun_out/ncu_enc_r2c.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 1500 gpurun_out/r2c_bench.err; head -c 3000 gpurun_out/r2c_bench.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err; head -c 600 gpurun_out/r2c_bench_ref.json
ls -la gpurun_out/*.ncu-rep
