#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for o in "" "14=0"; do
  echo -n "opts[$o] "; CKKS_B200_OPTIONS="$o" timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step']*1e3,1), d['gpu_launches'], round(d['e2e']['value'],1), round(d['e2e']['serial_value'],1), 'ntt', round(d['roofline']['achieved'],1), d.get('cpu_baseline',{}).get('gpu_result_bit_exact'))"
done | tee gpurun_out/lab6.txt
