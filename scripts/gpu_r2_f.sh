#!/bin/bash
# round 2, call F: shared-memory carve-out for the codec kernels -- exact encoders (warp / team), decoders, fast parse
mkdir -p gpurun_out
SPECS="lz4:1:4 lz4:1:2 lz4:1:8 lz4:1:16 blosclz:2:8 blosclz:1:4"
BLOSC_B200_LZ4_TEAM=0 timeout 600 python scripts/kbench.py warp $SPECS 2>&1 | tee gpurun_out/r2f_kbench_warp.log
timeout 600 python scripts/kbench.py team lz4:1:4 lz4:1:2 lz4:1:8 lz4:1:16 2>&1 | tee gpurun_out/r2f_kbench_team.log
BLOSC_B200_PARSE=fast timeout 600 python scripts/kbench.py fast lz4:1:4 lz4:1:2 lz4:1:8 lz4:1:16 2>&1 | tee gpurun_out/r2f_kbench_fast.log
