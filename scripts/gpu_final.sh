#!/bin/bash
# Round-end evidence run: tests, smoke, both bench arms, both workloads, ncu launch list, sweep.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 600 python bench.py --workload blosclz-bitshuffle-ts8-cl5-256MiB --steps 10 --warmup 3 > gpurun_out/bench_ours_cfg3.json 2> gpurun_out/bench_ours_cfg3.err
timeout 600 python bench.py --impl reference --workload blosclz-bitshuffle-ts8-cl5-256MiB --steps 10 --warmup 3 > gpurun_out/bench_ref_cfg3.json 2> gpurun_out/bench_ref_cfg3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
timeout 900 python scripts/sweep.py > gpurun_out/sweep.txt 2>&1
timeout 600 python bench.py --workload lz4-shuffle-cl5-8GiB-sharded --steps 5 --warmup 2 > gpurun_out/bench_sharded_n1.json 2> gpurun_out/bench_sharded_n1.err
timeout 600 python bench.py --impl reference --workload lz4-shuffle-cl5-8GiB-sharded --steps 1 --warmup 1 > gpurun_out/bench_sharded_ref.json 2>&1
bash scripts/gpu_ncu.sh lz4-shuffle-ts4-cl5-256MiB final > gpurun_out/ncu_final.log 2>&1
cat gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/smoke.txt
python - <<'PY'
import json
for f in ("bench_ref","bench_ours","bench_ref_cfg3","bench_ours_cfg3","bench_sharded_n1","bench_sharded_ref"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.1f comp %.1f dec %.1f e2e %.1f ms/step %.2f" % (d["value"], d["compress_gbs"], d["decompress_gbs"], d["e2e"]["value"], d["ms_per_step"]), d.get("roofline",{}).get("frac"), d.get("clocks"), d.get("concurrent",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
