#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/fp64_lat.bin 2>&1 | tee gpurun_out/fp64_lat.txt
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "fast_transforms" 2>&1 | tail -5
for l in libckks_b200.so libckks_b200_c2.so; do
  CKKS_B200_LIB=liberate-fhe_b200/csrc/$l timeout 300 python scripts/ntt_lab.py --opts "3=1;3=1,2=0;2=0" 2>&1 | tee -a gpurun_out/lab1.txt
done
CKKS_B200_LIB=liberate-fhe_b200/csrc/libckks_b200.so timeout 300 python scripts/ntt_lab.py --logN 17 --rows 240 --period 60 --big 6 --opts "3=1" 2>&1 | tee -a gpurun_out/lab1.txt
