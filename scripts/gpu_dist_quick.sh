#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_flow_check.py > gpurun_out/dist_flow_2.log 2>&1; grep -E 'DIST_FLOW|Error|error' gpurun_out/dist_flow_2.log | tail -5
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | grep -E '^\{|rror' | tail -2 | cut -c1-260
