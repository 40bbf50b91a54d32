#!/bin/bash
# ncu --set full capture of the codec kernels (one launch each) on the default bench workload
mkdir -p gpurun_out
W=${1:-lz4-shuffle-ts4-cl5-256MiB}
TAG=${2:-r1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -f -o gpurun_out/enc_$TAG python bench.py --workload $W --steps 1 --warmup 3 > gpurun_out/ncu_enc_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 3 -c 1 -f -o gpurun_out/dec_$TAG python bench.py --workload $W --steps 1 --warmup 3 > gpurun_out/ncu_dec_$TAG.log 2>&1
tail -3 gpurun_out/ncu_enc_$TAG.log; ls -la gpurun_out/*.ncu-rep
