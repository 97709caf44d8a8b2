#!/bin/bash
# 2-GPU torchrun check of bench.py (both arms), bounded
mkdir -p gpurun_out
echo "=== bench g2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_g2.log 2>&1
echo "exit $?"; tail -c 1500 gpurun_out/bench_g2.log
echo "=== reference g2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/reference_g2.log 2>&1
echo "exit $?"; tail -c 800 gpurun_out/reference_g2.log
