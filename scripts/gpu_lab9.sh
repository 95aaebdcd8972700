#!/bin/bash
mkdir -p gpurun_out
for l in libckks_b200.so libckks_b200_col4.so; do
  CKKS_B200_LIB=liberate-fhe_b200/csrc/$l timeout 200 python scripts/ntt_lab.py --opts "10=1" --iters 20 2>&1 | cut -c1-260 | tee -a gpurun_out/lab9.txt
  echo -n "$l mult: "; CKKS_B200_LIB=liberate-fhe_b200/csrc/$l timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step']*1e3,1), 'ntt', round(d['roofline']['achieved'],1))" | tee -a gpurun_out/lab9.txt
done
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "fast_transforms and default" 2>&1 | tail -2
