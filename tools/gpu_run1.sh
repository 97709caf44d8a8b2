#!/bin/bash
# First GPU bring-up: each group in its own process so that a trapped kernel cannot poison the rest.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -5 gpurun_out/$name.log; }
run ops      python -m pytest tests/test_gpu_ops.py -m gpu -q -x --tb=short
run simt     python -m pytest tests/test_gpu_dense.py -m gpu -q --tb=short -k "simt"
run tc_lin   python -m pytest tests/test_gpu_dense.py -m gpu -q --tb=short -k "tc_linear"
run tc_conv  python -m pytest tests/test_gpu_dense.py -m gpu -q --tb=short -k "tc_conv"
run tc_diag  python tools/tc_diag.py
run pipe_simt python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short -k "simt"
run pipe_tc  python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short -k "tc and not full_size"
run full     python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short -k "full_size or batched"
