#!/bin/bash
mkdir -p gpurun_out
for o in "15=0" "15=1" "15=2" "15=3" "15=7"; do
echo -n "opts[$o] mult: "; CKKS_B200_OPTIONS="$o" timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step']*1e3,1), 'ntt', round(d['roofline']['achieved'],1), round(d['roofline']['ms_per_launch']*1e3,1))" | tee -a gpurun_out/lab9.txt
done
timeout 300 python -m pytest tests/test_gpu_ops.py -q -x -k "fast_transforms" 2>&1 | tail -2
