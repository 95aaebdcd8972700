#!/bin/bash
# full ncu captures of the fast forward NTT kernels (current build), sources imported
mkdir -p gpurun_out
for k in fast_fwd_blockpass fast_fwd_colpass fast_fwd_blockpass_persist; do
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:^${k}\$" -s 3 -c 1 -f -o gpurun_out/prof3_${k} python scripts/ntt_bench.py --quick > gpurun_out/ncu_${k}.log 2>&1
  tail -2 gpurun_out/ncu_${k}.log
done
