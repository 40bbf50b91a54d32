#!/bin/bash
# round 2, last call: bench.py both arms on the final code (traffic measured in the run) + ncu of the pair decoder
mkdir -p gpurun_out
T=r2f
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; head -c 200 gpurun_out/${T}_bench_ref.json; echo
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err; head -c 300 gpurun_out/${T}_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --workload lz4-shuffle-ts4-cl5-256MiB --no-cpu --no-traffic > gpurun_out/${T}_bench_under_ncu.log 2>&1
cap() {
  env $4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $5 -c 1 -f -o gpurun_out/${T}_$1 python scripts/kbench.py ncu $3 > gpurun_out/${T}_ncu_$1.log 2>&1
  python scripts/ncu_brief.py gpurun_out/${T}_$1.ncu-rep 12 > gpurun_out/${T}_ncu_$1.txt 2>&1
  python scripts/ncu_lines.py gpurun_out/${T}_$1.ncu-rep c-blosc_b200/lib/libblosc_b200.so $6 14 >> gpurun_out/${T}_ncu_$1.txt 2>&1
  rm -f gpurun_out/${T}_$1.ncu-rep
}
cap decode_pair decode_pair_kernel lz4:1:4 X=1 3 decode_pair_kernel
cap decode_lz4 decode_kernel lz4:1:4 BLOSC_B200_LZ4D_PAIR=0 3 decode_kernelILi1E
BLOSC_B200_LZ4_TEAM=0 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/${T}_racecheck_noteam.log 2>&1; echo "racecheck (team encoder off, pair decoder on) rc=$?"; tail -2 gpurun_out/${T}_racecheck_noteam.log
BLOSC_B200_PARSE=fast timeout 300 python scripts/kbench.py fast lz4:1:4 lz4:1:8 2>&1 | tee gpurun_out/${T}_kbench_fast.log | cut -c1-300
timeout 300 python scripts/kbench.py exact lz4:1:4 lz4:1:2 lz4:1:8 lz4:1:16 blosclz:2:8 lz4hc:1:4 2>&1 | tee gpurun_out/${T}_kbench_exact.log | cut -c1-300
