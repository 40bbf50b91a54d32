#!/bin/bash
# round 2, call B: tile-speculative LZ4 chain encoder -- parity on hardware, A/B against the previous loop, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_hostile_input.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
SPECS="lz4:1:4 lz4:1:2 lz4:1:8 lz4:1:16 lz4:2:4 lz4:0:4 blosclz:2:8"
BLOSC_B200_LIB=c-blosc_b200/lib/libblosc_b200_old.so timeout 600 python scripts/kbench.py old $SPECS 2>&1 | tee gpurun_out/r2b_kbench_old.log
timeout 600 python scripts/kbench.py new $SPECS 2>&1 | tee gpurun_out/r2b_kbench_new.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -f -o gpurun_out/enc_r2b python scripts/kbench.py ncu lz4:1:4 > gpurun_out/ncu_enc_r2b.log 2>&1
ls -la gpurun_out/*.ncu-rep
