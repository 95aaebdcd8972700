#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -1 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ours.log; cut -c1-160 gpurun_out/bench_ours.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --graph off --profile-range > gpurun_out/ncu_launch.log 2>&1
