#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for o in "" "9=40" "9=32" "9=100" "10=1,9=400" "10=3,9=40" "10=3,9=64" "11=400" "11=24" "3=1,9=40"; do
  echo -n "opts[$o] "; CKKS_B200_OPTIONS="$o" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step']*1e3,1), d['gpu_launches'], round(d['e2e']['value'],1), 'ntt', round(d['roofline']['achieved'],1), round(d['roofline']['ms_per_launch']*1e3,1))"
done | tee gpurun_out/lab5.txt
