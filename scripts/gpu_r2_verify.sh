#!/bin/bash
# round 2, verification of the final state: whole GPU suite, racecheck with the team encoder off, kernel sweeps
mkdir -p gpurun_out
T=r2v
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
BLOSC_B200_LZ4_TEAM=0 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_small.py > gpurun_out/${T}_racecheck_noteam.log 2>&1; echo "racecheck (team encoder off, pair decoder on) rc=$?"; tail -2 gpurun_out/${T}_racecheck_noteam.log
BLOSC_B200_PARSE=fast timeout 300 python scripts/kbench.py fast lz4:1:4 lz4:1:2 lz4:1:8 lz4:1:16 2>&1 | tee gpurun_out/${T}_kbench_fast.log | cut -c1-300
timeout 300 python scripts/kbench.py exact lz4:1:4 lz4:1:2 lz4:1:8 lz4:1:16 blosclz:2:8 lz4hc:1:4 2>&1 | tee gpurun_out/${T}_kbench_exact.log | cut -c1-300
