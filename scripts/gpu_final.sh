#!/bin/bash
# final evidence of the round (kept under 64 MiB): parity tests, smoke, bench, op bench, launch list, roofline traffic, one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_ours.log; cut -c1-200 gpurun_out/bench_ours.log
timeout 600 python scripts/op_bench.py 2>&1 | grep '^{' | tee gpurun_out/op_bench.log | cut -c1-300
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --graph off --profile-range > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --profile-from-start off --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/roofline_traffic.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --graph off --profile-roofline > gpurun_out/ncu_roofline.log 2>&1
for k in fast_fwd_blockpass_h; do
  timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:${k}" -s 1 -c 1 -f -o gpurun_out/prof9_${k} python bench.py --steps 1 --warmup 2 --no-cpu-baseline --graph off --profile-range > gpurun_out/ncu9_${k}.log 2>&1
  tail -1 gpurun_out/ncu9_${k}.log | cut -c1-120
done
du -sh gpurun_out
