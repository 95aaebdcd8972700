#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for o in "" "10=1,9=400"; do
  echo -n "opts[$o] "; CKKS_B200_OPTIONS="$o" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step']*1e3,1), d['gpu_launches'], round(d['e2e']['value'],1), round(d['e2e']['serial_value'],1), 'ntt', round(d['roofline']['achieved'],1), round(d['roofline']['ms_per_launch']*1e3,1), d['clocks'])"
done | tee gpurun_out/lab6.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-cpu-baseline --profile-range > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-100
