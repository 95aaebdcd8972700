#!/bin/bash
mkdir -p gpurun_out
for k in fast_fwd_blockpass_pp fast_colpass_pp; do
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:${k}" -s 4 -c 1 -f -o gpurun_out/prof6_${k} python scripts/ntt_lab.py --opts "3=2,4=1" --iters 2 > gpurun_out/ncu6_${k}.log 2>&1
tail -2 gpurun_out/ncu6_${k}.log | cut -c1-200
done
