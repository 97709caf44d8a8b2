#!/bin/bash
mkdir -p gpurun_out
echo "=== train tests"
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q --tb=short -x > gpurun_out/train.log 2>&1
echo "exit $?"; tail -40 gpurun_out/train.log
