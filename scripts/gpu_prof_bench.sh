#!/bin/bash
# full ncu captures of the key switch's forward NTT launches inside the real mult (bench.py --profile-range)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_e2e.log; cat gpurun_out/bench_e2e.log | cut -c1-900
for v in 0 1; do
  k=fast_fwd_blockpass; [ $v == 1 ] && k=fast_fwd_blockpass_persist
  CKKS_B200_OPTIONS="1=$v" timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:^${k}\$" -s 1 -c 1 -f -o gpurun_out/prof4_ks_${k} python bench.py --steps 1 --warmup 2 --no-cpu-baseline --profile-range > gpurun_out/ncu4_${k}.log 2>&1
  tail -2 gpurun_out/ncu4_${k}.log | cut -c1-300
done
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:^fast_fwd_colpass\$" -s 1 -c 1 -f -o gpurun_out/prof4_ks_fast_fwd_colpass python bench.py --steps 1 --warmup 2 --no-cpu-baseline --profile-range > gpurun_out/ncu4_col.log 2>&1
