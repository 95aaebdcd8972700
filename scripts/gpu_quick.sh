#!/bin/bash
# short GPU session: parity tests, sweep, bench, launch list of the timed steps
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "=== ntt sweep"; timeout 600 python scripts/ntt_bench.py 2>&1 | tee gpurun_out/ntt_bench.log | grep -E "^(16|17) (36|60)"
echo "=== bench ours"; timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_ours.log
echo "=== ncu launch list (timed steps only)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --profile-range > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
