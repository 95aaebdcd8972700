#!/bin/bash
# multi-GPU session: distributed golden flow + scaling bench at N = 1 and N = $1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
echo "=== dist flow check (world $N)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_flow_check.py > gpurun_out/dist_flow_$N.log 2>&1; grep -v '^\s*$' gpurun_out/dist_flow_$N.log | grep -E 'DIST_FLOW|rank [0-9]+:|Error|error|File|line' | tail -30
echo "=== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_N$N.log
echo "=== bench N=1"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_N1.log
