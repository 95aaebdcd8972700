#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_csprng.py -q -x 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step']*1e3,1), d['gpu_launches'], round(d['e2e']['value'],1), d.get('cpu_baseline',{}).get('gpu_result_bit_exact'))"
