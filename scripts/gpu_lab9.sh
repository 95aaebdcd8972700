#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_vs_reference_engine.py -q -x 2>&1 | tail -2
for o in "" "16=0"; do
echo -n "opts[$o] mult: "; CKKS_B200_OPTIONS="$o" timeout 200 python bench.py --steps 30 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step']*1e3,1), d['gpu_launches'], d.get('cpu_baseline',{}).get('gpu_result_bit_exact'))" | tee -a gpurun_out/lab9.txt
done
