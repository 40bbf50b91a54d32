#!/bin/bash
# compute-sanitizer memcheck + racecheck on the small workloads
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_case.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
grep -c "hazard" gpurun_out/racecheck.log; tail -4 gpurun_out/racecheck.log
