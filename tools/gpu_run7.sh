#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log | cut -c1-300; }
run 150 ops python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x
run 200 pipe python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short
run 120 smoke python __graft_entry__.py smoke
run 240 bench python bench.py --steps 20 --warmup 3
run 200 bench_b1 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline
run 200 bench_b4 python bench.py --steps 10 --warmup 3 --batch 4 --no-cpu-baseline
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 260 --csv --log-file gpurun_out/launches_r05.csv python bench.py --steps 2 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
ls -la gpurun_out
