#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "fast_transforms" 2>&1 | tail -3
echo "== key-switch mix"; timeout 300 python scripts/ntt_lab.py --opts "8=1;8=1,3=1;8=1,2=0" 2>&1 | tee -a gpurun_out/lab2.txt
echo "== slab (130 rows, period 13, 2 big)"; timeout 300 python scripts/ntt_lab.py --rows 130 --period 13 --big 2 --opts "8=1;8=1,3=1" 2>&1 | tee -a gpurun_out/lab2.txt
echo "== e2e diag"; timeout 300 python scripts/e2e_diag.py 2>&1 | grep -v Warn | tail -6 | tee gpurun_out/e2e_diag.txt
