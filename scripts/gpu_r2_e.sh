#!/bin/bash
# round 2, call E: fast parse iteration -- speed + ncu of parse_kernel (cfg 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fast_parse.py -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -2 gpurun_out/r2e_pytest.log
SPECS="${SPECS:-lz4:1:4 lz4:1:2 lz4:1:8 lz4:1:16 lz4:2:4 lz4:0:4}"
BLOSC_B200_PARSE=fast timeout 600 python scripts/kbench.py fast $SPECS 2>&1 | tee gpurun_out/r2e_kbench_fast.log
BLOSC_B200_PARSE=fast timeout 900 ncu --set full --clock-control none --import-source on -k regex:parse_kernel -s 3 -c 1 -f -o gpurun_out/parse_r2e python scripts/kbench.py ncu lz4:1:4 > gpurun_out/ncu_parse_r2e.log 2>&1
BLOSC_B200_PARSE=fast timeout 900 ncu --set full --clock-control none --import-source on -k regex:index_kernel -s 3 -c 1 -f -o gpurun_out/index_r2e python scripts/kbench.py ncu lz4:1:4 > gpurun_out/ncu_index_r2e.log 2>&1
ls -la gpurun_out/*r2e*.ncu-rep
