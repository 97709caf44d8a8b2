#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-250; }
run 200 tests python -m pytest tests/test_gpu_dense.py tests/test_gpu_pipeline.py -m gpu -q --tb=short -x
run 120 bench_b1 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline
grep -o '"phases_ms": {[^}]*}' gpurun_out/bench_b1.log
run 120 bench_b4 python bench.py --steps 20 --warmup 3 --batch 4 --no-cpu-baseline
grep -o '"phases_ms": {[^}]*}' gpurun_out/bench_b4.log
