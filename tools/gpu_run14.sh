#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log | cut -c1-400; }
run 240 dense python -m pytest tests/test_gpu_dense.py -m gpu -q --tb=short -x
run 240 pipe python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_ops.py tests/test_gpu_train.py -m gpu -q --tb=short -x
run 200 bench_b1 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline
grep -o '"phases_ms": {[^}]*}' gpurun_out/bench_b1.log
DC_TC_HALO=0 timeout 200 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/bench_b1_nohalo.log 2>&1
grep -o '"phases_ms": {[^}]*}' gpurun_out/bench_b1_nohalo.log
run 200 bench_b4 python bench.py --steps 20 --warmup 3 --batch 4 --no-cpu-baseline
grep -o '"phases_ms": {[^}]*}' gpurun_out/bench_b4.log
