#!/bin/bash
# round 2, call A: reference CPU sweep on the GPU box's host + source-level ncu capture of the encoder
mkdir -p gpurun_out
lscpu > gpurun_out/r2_lscpu.txt 2>&1
timeout 900 python scripts/cpu_ref.py cfg2 cfg3 > gpurun_out/r2_cpu_sweep.json 2> gpurun_out/r2_cpu_sweep.log
grep best gpurun_out/r2_cpu_sweep.log
bash scripts/gpu_ncu.sh lz4-shuffle-ts4-cl5-256MiB r2a
