#!/bin/bash
# round 2 evidence run: whole GPU suite, smoke, both bench arms (cfg 2 + cfg 3 + fast parse + cfg 5 in one line), launch
# list, `ncu --set full` captures of every kernel of the path with text summaries, compute-sanitizer.
mkdir -p gpurun_out
T=r2f
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; head -c 300 gpurun_out/${T}_bench_ref.json; echo
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 400 gpurun_out/${T}_bench.err; head -c 700 gpurun_out/${T}_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --workload lz4-shuffle-ts4-cl5-256MiB --no-cpu --no-traffic > gpurun_out/${T}_bench_under_ncu.log 2>&1
cap() {  # name kernel-regex spec env skip mangled-name-substring
  env $4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $5 -c 1 -f -o gpurun_out/${T}_$1 python scripts/kbench.py ncu $3 > gpurun_out/${T}_ncu_$1.log 2>&1
  python scripts/ncu_brief.py gpurun_out/${T}_$1.ncu-rep 12 > gpurun_out/${T}_ncu_$1.txt 2>&1
  python scripts/ncu_lines.py gpurun_out/${T}_$1.ncu-rep c-blosc_b200/lib/libblosc_b200.so $6 14 >> gpurun_out/${T}_ncu_$1.txt 2>&1
  case "$1" in encode_team|parse) ;; *) rm -f gpurun_out/${T}_$1.ncu-rep ;; esac     # gpurun_out/ travels back only below 64 MiB
}
cap encode_team encode_team_kernel lz4:1:4 X=1 3 encode_team_kernel
cap decode_pair decode_pair_kernel lz4:1:4 X=1 3 decode_pair_kernel
cap decode_lz4 decode_kernel lz4:1:4 BLOSC_B200_LZ4D_PAIR=0 3 decode_kernelILi1E
cap filter_shuffle filter_kernel lz4:1:4 X=1 6 filter_kernelILi4E
cap filter_unshuffle filter_kernel lz4:1:4 X=1 7 filter_kernelILi4E
cap compact compact_kernel lz4:1:4 X=1 3 compact_kernel
cap index index_kernel lz4:1:4 BLOSC_B200_PARSE=fast 3 index_kernel
cap parse parse_kernel lz4:1:4 BLOSC_B200_PARSE=fast 3 parse_kernel
cap fscan fscan_kernel lz4:1:4 BLOSC_B200_PARSE=fast 3 fscan_kernel
cap compact_fast compact_kernel lz4:1:4 BLOSC_B200_PARSE=fast 3 compact_kernel
cap encode_blosclz encode_kernel blosclz:2:8 X=1 3 _Z13encode_kernel
cap decode_blosclz decode_kernel blosclz:2:8 X=1 3 decode_kernelILi0E
cap filter_bitshuffle filter_kernel blosclz:2:8 X=1 6 filter_kernelILi8E
cap filter_bitunshuffle filter_kernel blosclz:2:8 X=1 7 filter_kernelILi8E
cap encode_warp_ts2 encode_kernel lz4:1:2 X=1 3 _Z13encode_kernel
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_case.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/${T}_memcheck.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_fast.py > gpurun_out/${T}_memcheck_fast.log 2>&1; echo "memcheck fast rc=$?"; tail -3 gpurun_out/${T}_memcheck_fast.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/${T}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/${T}_racecheck.log
BLOSC_B200_PARSE=fast timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/${T}_racecheck_fast.log 2>&1; echo "racecheck fast rc=$?"; tail -3 gpurun_out/${T}_racecheck_fast.log
ls gpurun_out | grep ${T}_ | wc -l
