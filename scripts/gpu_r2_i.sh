#!/bin/bash
# round 2, call I: filter grouping + host getitem on hardware; ncu of decode_kernel and filter_kernel (cfg 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_frames.py -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; tail -2 gpurun_out/r2i_pytest.log
timeout 300 python scripts/kbench.py grp lz4:1:4 lz4:1:8 lz4:1:2 blosclz:2:8 2>&1 | tee gpurun_out/r2i_kbench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 3 -c 1 -f -o gpurun_out/decode_r2i python scripts/kbench.py ncu lz4:1:4 > gpurun_out/ncu_decode_r2i.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 6 -c 2 -f -o gpurun_out/filter_r2i python scripts/kbench.py ncu lz4:1:4 > gpurun_out/ncu_filter_r2i.log 2>&1
ls -la gpurun_out/*r2i*.ncu-rep
