#!/bin/bash
mkdir -p gpurun_out
show() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['n_gpus'], d['config'].get('cuda_graph'), round(d['value'],1), round(d['ms_per_step']*1e3,1), d['gpu_launches'], 'e2e', round(d['e2e']['value'],1))"; }
for g in on; do
  echo "N=2 graph=$g: "; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --graph $g > gpurun_out/g2_$g.log 2>&1; echo "exit $?"; grep -E '^\{' gpurun_out/g2_$g.log | tail -1 | show; grep -iE "error|Traceback" gpurun_out/g2_$g.log | head -5
done
echo "N=1 graph=on: "; timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --graph on > gpurun_out/g1_on.log 2>&1; echo "exit $?"; tail -1 gpurun_out/g1_on.log | show; grep -iE "error|Traceback" gpurun_out/g1_on.log | head -5
