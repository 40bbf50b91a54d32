#!/bin/bash
# round 2, verification of the final state (after the BloscLZ dense decode path): GPU suite, bench both arms, sweep
mkdir -p gpurun_out
T=r2w
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; head -c 200 gpurun_out/${T}_bench_ref.json; echo
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.err; head -c 300 gpurun_out/${T}_bench.json; echo
timeout 600 python scripts/sweep.py > gpurun_out/${T}_sweep.log 2>&1; cp gpurun_out/sweep.json gpurun_out/${T}_sweep.json; tail -3 gpurun_out/${T}_sweep.log
