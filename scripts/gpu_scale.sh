#!/bin/bash
# scaling bench on one box: N = 8, 4, 2, 1 (strong scaling of one gold mult)
mkdir -p gpurun_out
for N in 8 4 2; do
  echo "=== bench N=$N"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 3 2>&1 | grep -E '^\{|Error|error' | tail -2 | tee gpurun_out/bench_N$N.log | cut -c1-250
done
echo "=== bench N=1"; timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_N1.log | cut -c1-250
